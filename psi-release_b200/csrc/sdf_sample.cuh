// One trilinear SDF sample with analytic gradient -- the arithmetic of
// F.grid_sample(sdf, normalised verts, padding_mode='border', align_corners=True) as the reference
// calls it (source/fitting_habitat.py:145-152; torch 1.2 semantics), in the
// reference's operation order.  Shared by sdf_fwd_kernel (sdf.cu) and the fused skinning epilogue
// of the fitting loop (lbs.cu).
#pragma once
namespace psi {

struct SdfGrid {
    const float *grid;      // [D][D][D], x slowest, z contiguous
    int D;
    float gmin[3], inv_extent[3], gscale[3];   // min, 1/(max-min), (D-1)/(max-min)
};

// A sample in two halves, so that a thread can put the eight gathers of SEVERAL samples in flight before it consumes
// any of them (the lookup is latency bound): sdf_prepare -> corner offsets + weights, the caller loads the eight
// corners (sdf_corner_offset), sdf_finish -> value + gradient.  Same arithmetic, same order as before the split.
struct SdfCell {
    size_t r00, r01, r10, r11;       // row offsets of the four (x, y) corner columns
    int z0, z1;
    float wx0, wx1, wy0, wy1, wz0, wz1;
    float gm[3];                     // d f / d v per axis (0 outside the grid: border clamp)
    bool xin, yin, zin;
};

__device__ __forceinline__ SdfCell sdf_prepare(const SdfGrid &sg, float px, float py, float pz) {
    const int D = sg.D;
    const float dm1 = (float)(D - 1);
    const float xyz[3] = {px, py, pz};
    float f[3];
    SdfCell c;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // reference operation order: (v-min)/(max-min)*2-1, then ((u+1)/2)*(D-1)
        const float u = (xyz[a] - sg.gmin[a]) * sg.inv_extent[a] * 2.0f - 1.0f;
        float cc = ((u + 1.0f) * 0.5f) * dm1;
        float mult = sg.gscale[a];
        if (!(cc > 0.0f)) { cc = 0.0f; mult = 0.0f; }          // border clamp, zero gradient
        else if (cc >= dm1) { cc = dm1; mult = 0.0f; }
        f[a] = cc;
        c.gm[a] = mult;
    }
    int i0[3];
    float w0[3], w1[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float fl = floorf(f[a]);
        i0[a] = (int)fl;
        w1[a] = f[a] - fl;
        w0[a] = (fl + 1.0f) - f[a];
    }
    const int x1 = min(i0[0] + 1, D - 1), y1 = min(i0[1] + 1, D - 1);
    // corners past the border carry weight 0 (f is clamped to D-1 => w1 == 0): clamping the
    // index keeps the load in bounds without changing the value.
    c.z0 = i0[2];
    c.z1 = min(i0[2] + 1, D - 1);
    c.r00 = ((size_t)i0[0] * D + i0[1]) * D; c.r01 = ((size_t)i0[0] * D + y1) * D;
    c.r10 = ((size_t)x1 * D + i0[1]) * D;    c.r11 = ((size_t)x1 * D + y1) * D;
    c.xin = i0[0] + 1 <= D - 1; c.yin = i0[1] + 1 <= D - 1; c.zin = i0[2] + 1 <= D - 1;
    c.wx0 = w0[0]; c.wx1 = c.xin ? w1[0] : 0.f;
    c.wy0 = w0[1]; c.wy1 = c.yin ? w1[1] : 0.f;
    c.wz0 = w0[2]; c.wz1 = c.zin ? w1[2] : 0.f;
    return c;
}

// the eight corner values in the order s000, s001, s010, s011, s100, s101, s110, s111
__device__ __forceinline__ void sdf_gather(const SdfGrid &sg, const SdfCell &c, float (&s)[8]) {
    const float *__restrict__ grid = sg.grid;
    s[0] = __ldg(grid + c.r00 + c.z0); s[1] = __ldg(grid + c.r00 + c.z1);
    s[2] = __ldg(grid + c.r01 + c.z0); s[3] = __ldg(grid + c.r01 + c.z1);
    s[4] = __ldg(grid + c.r10 + c.z0); s[5] = __ldg(grid + c.r10 + c.z1);
    s[6] = __ldg(grid + c.r11 + c.z0); s[7] = __ldg(grid + c.r11 + c.z1);
}

__device__ __forceinline__ float sdf_finish(const SdfCell &c, const float (&s)[8], float *g3) {
    // interpolate along z, then y, then x
    const float c00 = s[0] * c.wz0 + s[1] * c.wz1, c01 = s[2] * c.wz0 + s[3] * c.wz1;
    const float c10 = s[4] * c.wz0 + s[5] * c.wz1, c11 = s[6] * c.wz0 + s[7] * c.wz1;
    const float c0 = c00 * c.wy0 + c01 * c.wy1, c1 = c10 * c.wy0 + c11 * c.wy1;
    const float val = c0 * c.wx0 + c1 * c.wx1;
    if (g3) {
        const float gx = c.xin ? (c1 - c0) : 0.f;
        const float gy = c.yin ? ((c01 - c00) * c.wx0 + (c11 - c10) * c.wx1) : 0.f;
        const float d00 = c.zin ? (s[1] - s[0]) : 0.f, d01 = c.zin ? (s[3] - s[2]) : 0.f;
        const float d10 = c.zin ? (s[5] - s[4]) : 0.f, d11 = c.zin ? (s[7] - s[6]) : 0.f;
        const float gz = (d00 * c.wy0 + d01 * c.wy1) * c.wx0 + (d10 * c.wy0 + d11 * c.wy1) * c.wx1;
        g3[0] = gx * c.gm[0];
        g3[1] = gy * c.gm[1];
        g3[2] = gz * c.gm[2];
    }
    return val;
}

// returns the interpolated value; g3 = d value / d (x,y,z) (0 outside the grid: border clamp)
__device__ __forceinline__ float sdf_sample(const SdfGrid &sg, float px, float py, float pz, float *g3) {
    const SdfCell c = sdf_prepare(sg, px, py, pz);
    float s[8];
    sdf_gather(sg, c, s);
    return sdf_finish(c, s, g3);
}

}  // namespace psi
