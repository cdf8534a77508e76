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

// returns the interpolated value; g3 = d value / d (x,y,z) (0 outside the grid: border clamp)
__device__ __forceinline__ float sdf_sample(const SdfGrid &sg, float px, float py, float pz, float *g3) {
    const float *__restrict__ grid = sg.grid;
    const int D = sg.D;
    const float dm1 = (float)(D - 1);
    const float xyz[3] = {px, py, pz};
    float f[3], gm[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // reference operation order: (v-min)/(max-min)*2-1, then ((u+1)/2)*(D-1)
        const float u = (xyz[a] - sg.gmin[a]) * sg.inv_extent[a] * 2.0f - 1.0f;
        float c = ((u + 1.0f) * 0.5f) * dm1;
        float mult = sg.gscale[a];
        if (!(c > 0.0f)) { c = 0.0f; mult = 0.0f; }          // border clamp, zero gradient
        else if (c >= dm1) { c = dm1; mult = 0.0f; }
        f[a] = c;
        gm[a] = mult;
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
    const int x1 = min(i0[0] + 1, D - 1), y1 = min(i0[1] + 1, D - 1), z1 = min(i0[2] + 1, D - 1);
    // corners past the border carry weight 0 (f is clamped to D-1 => w1 == 0): clamping the
    // index keeps the load in bounds without changing the value.
    const size_t r00 = ((size_t)i0[0] * D + i0[1]) * D, r01 = ((size_t)i0[0] * D + y1) * D;
    const size_t r10 = ((size_t)x1 * D + i0[1]) * D, r11 = ((size_t)x1 * D + y1) * D;
    const float s000 = __ldg(grid + r00 + i0[2]), s001 = __ldg(grid + r00 + z1);
    const float s010 = __ldg(grid + r01 + i0[2]), s011 = __ldg(grid + r01 + z1);
    const float s100 = __ldg(grid + r10 + i0[2]), s101 = __ldg(grid + r10 + z1);
    const float s110 = __ldg(grid + r11 + i0[2]), s111 = __ldg(grid + r11 + z1);
    const bool xin = i0[0] + 1 <= D - 1, yin = i0[1] + 1 <= D - 1, zin = i0[2] + 1 <= D - 1;
    const float wx0 = w0[0], wx1 = xin ? w1[0] : 0.f;
    const float wy0 = w0[1], wy1 = yin ? w1[1] : 0.f;
    const float wz0 = w0[2], wz1 = zin ? w1[2] : 0.f;
    // interpolate along z, then y, then x
    const float c00 = s000 * wz0 + s001 * wz1, c01 = s010 * wz0 + s011 * wz1;
    const float c10 = s100 * wz0 + s101 * wz1, c11 = s110 * wz0 + s111 * wz1;
    const float c0 = c00 * wy0 + c01 * wy1, c1 = c10 * wy0 + c11 * wy1;
    const float val = c0 * wx0 + c1 * wx1;
    if (g3) {
        const float gx = xin ? (c1 - c0) : 0.f;
        const float gy = yin ? ((c01 - c00) * wx0 + (c11 - c10) * wx1) : 0.f;
        const float d00 = zin ? (s001 - s000) : 0.f, d01 = zin ? (s011 - s010) : 0.f;
        const float d10 = zin ? (s101 - s100) : 0.f, d11 = zin ? (s111 - s110) : 0.f;
        const float gz = (d00 * wy0 + d01 * wy1) * wx0 + (d10 * wy0 + d11 * wy1) * wx1;
        g3[0] = gx * gm[0];
        g3[1] = gy * gm[1];
        g3[2] = gz * gm[2];
    }
    return val;
}

}  // namespace psi
