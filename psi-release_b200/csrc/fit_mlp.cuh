// VPoser decoder forward / backward for the fused fitting loop, one CTA (512 threads) per body.
//
// The three Linear layers are matrix-vector products against weights shared by every body:
// each CTA STREAMS the weight matrix through shared memory with the TMA engine (cp.async.bulk,
// one 32 kB bulk copy per tile, 3-stage mbarrier ring) while its threads consume the previous tile
// -- no per-thread dependent-load chain (the first version was latency bound: 283 us for 64 bodies).
// Weights stay L2 resident (2.7 MB); reductions use a fixed order (bit-reproducible).
//
// included by fit.cu after gs_fwd / gs_bwd / lrelu are defined
#pragma once

namespace psi {

constexpr int kMlpThreads = 512;
constexpr int kMvStages = 3;
constexpr int kMvTileFloats = 8192;   // 32 kB

struct MvState {
    float *ring;       // kMvStages * kMvTileFloats floats, 128-byte aligned
    uint64_t *bars;    // kMvStages mbarriers
    float *red;        // kMlpThreads floats
    uint32_t tiles;    // tiles consumed so far (same value in every thread)
};

// y_s[n] = sum_k W[k*N+n] * v_s[k]   (W row-major [K][N] in global memory, N*4 % 16 == 0, N | 512,
// W 16-byte aligned).  Block-wide; y_s and v_s live in shared memory; ends with a __syncthreads.
__device__ __forceinline__ void matvec_stream(const float *__restrict__ W, int K, int N,
                                              const float *v_s, float *y_s, MvState &st) {
    const int tid = threadIdx.x;
    const int R = kMvTileFloats / N;                 // rows per tile
    const int ntiles = (K + R - 1) / R;
    const int groups = kMlpThreads / N, g = tid / N, col = tid - g * N;
    auto issue = [&](int t) {
        const uint32_t slot = (st.tiles + t) % kMvStages;
        const int rows = min(R, K - t * R);
        const uint32_t bytes = (uint32_t)rows * N * 4u;
        mbar_arrive_expect_tx(&st.bars[slot], bytes);
        tma_load_1d(st.ring + (size_t)slot * kMvTileFloats, W + (size_t)t * R * N, bytes, &st.bars[slot]);
    };
    if (tid == 0)
        for (int t = 0; t < kMvStages - 1 && t < ntiles; ++t) issue(t);
    float acc = 0.f;
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + kMvStages - 1 < ntiles) issue(t + kMvStages - 1);
        const uint32_t seq = st.tiles + t, slot = seq % kMvStages;
        mbar_wait(&st.bars[slot], (seq / kMvStages) & 1u);
        const float *tile = st.ring + (size_t)slot * kMvTileFloats;
        const int rows = min(R, K - t * R);
        const float *vv = v_s + t * R;
#pragma unroll 4
        for (int r = g; r < rows; r += groups) acc = fmaf(tile[r * N + col], vv[r], acc);
        __syncthreads();                              // the slot may be refilled
    }
    st.tiles += ntiles;
    st.red[tid] = acc;
    __syncthreads();
    if (tid < N) {
        float s = 0.f;
        for (int gg = 0; gg < groups; ++gg) s += st.red[gg * N + tid];
        y_s[tid] = s;
    }
    __syncthreads();
}

struct MlpSmem {
    float ring[kMvStages * kMvTileFloats];
    float red[kMlpThreads];
    float a[512], bvec[512], c[512];
    float sx[80], g[80];
    uint64_t bars[kMvStages];
};

__device__ __forceinline__ MvState mv_init(MlpSmem &S) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < kMvStages; ++i) mbar_init(&S.bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    MvState st;
    st.ring = S.ring;
    st.bars = S.bars;
    st.red = S.red;
    st.tiles = 0;
    return st;
}

// x layout [75]: t3 | 6D | betas10 | z32 | lh12 | rh12   (cvae.py:28-33 after convert_to_6D_rot)
// W1T [latent][512], W2T [512][512], W3Tp [512][128] (columns >= nbody*6 are zero)
__global__ void __launch_bounds__(kMlpThreads)
fit_prologue2_kernel(FitDims d, const float *__restrict__ x0, const float *__restrict__ x,
                     const float *__restrict__ W1T, const float *__restrict__ b1,
                     const float *__restrict__ W2T, const float *__restrict__ b2,
                     const float *__restrict__ W3Tp, const float *__restrict__ b3,
                     const float *__restrict__ hand_l, const float *__restrict__ hand_r,
                     const float *__restrict__ pose_mean, float w_rec, float w_vp,
                     float *__restrict__ rot, float *__restrict__ pose, float *__restrict__ shape,
                     float *__restrict__ transl, float *__restrict__ h1pre, float *__restrict__ h2pre,
                     float *__restrict__ o6, float *__restrict__ losses) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MlpSmem &S = *reinterpret_cast<MlpSmem *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int H = d.hidden, Lz = d.latent, NO = d.nbody * 6;
    const int xdim = 9 + 10 + Lz + 2 * d.ncomp;
    const int zoff = 19, lhoff = 19 + Lz, rhoff = lhoff + d.ncomp;
    MvState st = mv_init(S);
    for (int e = tid; e < xdim; e += blockDim.x) S.sx[e] = x[(size_t)b * xdim + e];
    __syncthreads();
    // layer 1
    matvec_stream(W1T, Lz, H, S.sx + zoff, S.a, st);
    for (int j = tid; j < H; j += blockDim.x) {
        const float a = S.a[j] + b1[j];
        h1pre[(size_t)b * H + j] = a;
        S.bvec[j] = lrelu(a);
    }
    __syncthreads();
    // layer 2
    matvec_stream(W2T, H, H, S.bvec, S.a, st);
    for (int j = tid; j < H; j += blockDim.x) {
        const float a = S.a[j] + b2[j];
        h2pre[(size_t)b * H + j] = a;
        S.c[j] = lrelu(a);
    }
    __syncthreads();
    // output layer (columns padded to 128)
    matvec_stream(W3Tp, H, 128, S.c, S.a, st);
    if (tid < NO) {
        const float a = S.a[tid] + b3[tid];
        S.a[tid] = a;
        o6[(size_t)b * NO + tid] = a;
    }
    __syncthreads();
    if (tid <= d.nbody) {
        float R[9];
        gs_fwd(tid == 0 ? S.sx + 3 : S.a + (tid - 1) * 6, R);
#pragma unroll
        for (int e = 0; e < 9; ++e) rot[((size_t)b * d.num_rot + tid) * 9 + e] = R[e];
    }
    const int hl0 = (d.J - 30) * 3, hr0 = (d.J - 15) * 3;
    for (int e = tid; e < d.J * 3; e += blockDim.x) {
        float v = pose_mean[e];
        if (e >= hr0) {
            for (int c = 0; c < d.ncomp; ++c) v = fmaf(S.sx[rhoff + c], hand_r[c * 45 + (e - hr0)], v);
        } else if (e >= hl0) {
            for (int c = 0; c < d.ncomp; ++c) v = fmaf(S.sx[lhoff + c], hand_l[c * 45 + (e - hl0)], v);
        } else if (e < d.num_rot * 3) {
            v = 0.f;
        }
        pose[(size_t)b * d.J * 3 + e] = v;
    }
    for (int e = tid; e < d.NB; e += blockDim.x) shape[(size_t)b * d.NB + e] = e < 10 ? S.sx[9 + e] : 0.f;
    if (tid < 3) transl[(size_t)b * 3 + tid] = S.sx[tid];
    if (tid == 0) {
        float r = 0.f, zz = 0.f;
        for (int e = 0; e < xdim; ++e) r += fabsf(x0[(size_t)b * xdim + e] - S.sx[e]);
        for (int i = 0; i < Lz; ++i) zz = fmaf(S.sx[zoff + i], S.sx[zoff + i], zz);
        losses[(size_t)b * 4 + 0] = w_rec * (r / (float)xdim);
        losses[(size_t)b * 4 + 1] = w_vp * (zz / (float)Lz);
    }
}

// W3 [nbody*6][512], W2 [512][512], W1 [512][latent=32]  (the reference's nn.Linear layout)
__global__ void __launch_bounds__(kMlpThreads)
fit_epilogue2_kernel(FitDims d, psi_fit_config cfg, int np_sdf, int nchunk, int num_contact,
                     const float *__restrict__ x0, float *__restrict__ x, float *__restrict__ am,
                     float *__restrict__ av, int *__restrict__ step, const float *__restrict__ W1,
                     const float *__restrict__ W2, const float *__restrict__ W3,
                     const float *__restrict__ hand_l, const float *__restrict__ hand_r,
                     const float *__restrict__ h1pre, const float *__restrict__ h2pre,
                     const float *__restrict__ o6, const float *__restrict__ grot,
                     const float *__restrict__ gpose, const float *__restrict__ gshape,
                     const float *__restrict__ gtransl, const float *__restrict__ partial,
                     const float *__restrict__ cpart, float *__restrict__ losses) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MlpSmem &S = *reinterpret_cast<MlpSmem *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int H = d.hidden, Lz = d.latent, NO = d.nbody * 6;
    const int xdim = 9 + 10 + Lz + 2 * d.ncomp;
    const int zoff = 19, lhoff = 19 + Lz, rhoff = lhoff + d.ncomp;
    MvState st = mv_init(S);
    for (int e = tid; e < xdim; e += blockDim.x) { S.sx[e] = x[(size_t)b * xdim + e]; S.g[e] = 0.f; }
    __syncthreads();
    if (tid <= d.nbody) {        // Gram-Schmidt backward: joint 0 -> x[3:9], others -> d o
        float dx6[6];
        gs_bwd(tid == 0 ? S.sx + 3 : o6 + (size_t)b * NO + (tid - 1) * 6,
               grot + ((size_t)b * d.num_rot + tid) * 9, dx6);
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            if (tid == 0) S.g[3 + e] = dx6[e];
            else S.a[(tid - 1) * 6 + e] = dx6[e];
        }
    }
    if (tid < 3) S.g[tid] = gtransl[(size_t)b * 3 + tid];
    if (tid >= 32 && tid < 42) S.g[9 + tid - 32] = gshape[(size_t)b * d.NB + tid - 32];
    if (tid >= 64 && tid < 64 + 2 * d.ncomp) {     // hand PCA backward
        const int u = tid - 64, c = u % d.ncomp;
        const bool right = u >= d.ncomp;
        const float *comp = (right ? hand_r : hand_l) + c * 45;
        const float *gp = gpose + (size_t)b * d.J * 3 + (right ? (d.J - 15) * 3 : (d.J - 30) * 3);
        float a = 0.f;
        for (int k = 0; k < 45; ++k) a = fmaf(comp[k], gp[k], a);
        S.g[(right ? rhoff : lhoff) + c] = a;
    }
    __syncthreads();
    // d h2 = W3^T d o (.) lrelu'
    matvec_stream(W3, NO, H, S.a, S.bvec, st);
    for (int i = tid; i < H; i += blockDim.x) S.c[i] = S.bvec[i] * lrelu_grad(h2pre[(size_t)b * H + i]);
    __syncthreads();
    matvec_stream(W2, H, H, S.c, S.bvec, st);
    for (int i = tid; i < H; i += blockDim.x) S.a[i] = S.bvec[i] * lrelu_grad(h1pre[(size_t)b * H + i]);
    __syncthreads();
    matvec_stream(W1, H, Lz, S.a, S.bvec, st);        // d z   (Lz == 32)
    if (tid < Lz) S.g[zoff + tid] = S.bvec[tid] + cfg.w_vposer * (2.0f * S.sx[zoff + tid] / (float)Lz);
    __syncthreads();
    // + d L_rec, then Adam (torch.optim.Adam defaults; bias corrections in double)
    const int t = step[b] + 1;
    if (tid < xdim) {
        const float xe = S.sx[tid], diff = xe - x0[(size_t)b * xdim + tid];
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float ge = S.g[tid] + cfg.w_rec * sgn / (float)xdim;
        const size_t o = (size_t)b * xdim + tid;
        const float m = cfg.beta1 * am[o] + (1.0f - cfg.beta1) * ge;
        const float v = cfg.beta2 * av[o] + (1.0f - cfg.beta2) * ge * ge;
        am[o] = m;
        av[o] = v;
        const double bc1 = 1.0 - pow((double)cfg.beta1, (double)t);
        const double bc2 = 1.0 - pow((double)cfg.beta2, (double)t);
        const float step_size = (float)((double)cfg.lr / bc1);
        const float denom = sqrtf(v) / (float)sqrt(bc2) + cfg.eps;
        x[o] = xe - (m / denom) * step_size;
    }
    if (tid == 0) {
        float sn = 0.f, cn = 0.f, cs = 0.f;
        for (int i = 0; i < np_sdf; ++i) {
            sn += partial[((size_t)b * np_sdf + i) * 2];
            cn += partial[((size_t)b * np_sdf + i) * 2 + 1];
        }
        for (int i = 0; i < nchunk; ++i) cs += cpart[(size_t)b * nchunk + i];
        losses[(size_t)b * 4 + 2] = cfg.w_contact * (cs / (float)num_contact);
        losses[(size_t)b * 4 + 3] = cfg.w_collision * (cn > 0.f ? sn / cn : 0.f);
    }
    __syncthreads();
    if (tid == 0) step[b] = t;
}

}  // namespace psi
