// VPoser decoder layers (vposer_smpl.py:107-121) and their backward for the fused fitting loop,
// as GEMMs over the WHOLE batch on the tensor cores.
//
// The first fused version ran one CTA per body and streamed every weight matrix through that CTA
// (64 x 1.37 MB of L2 traffic per pass, 37 + 40 us on 64 of the 148 SMs).  Here a layer is
//   Y[b][n] = epilogue( sum_k A[b][k] * W[n][k] )        b < B, n < N, k < K
// with the batch as the M dimension: a CTA owns 32 bodies x 8 outputs (N/8 x B/32 CTAs: 128 for the
// 512 x 512 layers at B = 64), both operands of its tile sit in shared memory (<= 80 kB, TMA bulk
// copies requested up front), FP32 operands are split on the fly (3xTF32, mma.sync m16n8k8, FP32
// accumulate).  Both operands use the swizzled chunk layout of common.cuh (a_index / swz);
// activations are written by the producing layer's epilogue directly in that layout.
// (Layer 1 and its backward, 32 <-> 512, are per-body matrix-vector products inside fit_step.)
#pragma once

namespace psi {

constexpr int kLT = 8;                                   // output columns per CTA (one n8 tile)
constexpr int kLB = 32;                                  // bodies per CTA (two m16 tiles)
constexpr int kLMaxChunks = 16;                          // K <= 512
constexpr int kLGroup = 4;                               // chunks per arrival barrier
constexpr int kLChunkBytes = (kLT + kLB) * kKC * 4;      // 1 kB of weights + 4 kB of activations per 32 k

struct LinearParams {
    const float *A;        // [body group][K/32][64][32], swizzled; rows of bodies >= B are zero
    const float *W;        // [N/8][K/32][8][32], swizzled: W[n][k], n = output, k = reduction index
    const float *bias;     // [n_valid] or null
    const float *gate;     // row-major [B][n_valid] pre-activations: multiply by lrelu'(gate); or null
    float *pre_out;        // row-major [B][n_valid]: value after the bias (saved for the backward); or null
    float *out_rm;         // row-major out_rm[b*ld + n], n < n_valid; or null
    float *outA;           // next layer's A operand (reduction length outA_kpad >= N); or null
    int K, N, B, n_valid, ld, outA_kpad, act;            // act 1: leaky ReLU 0.2 after the bias
};

// The kernel is latency bound, so the work of a layer is cut small: 4 warps = 2 body tiles x 2 halves of K; each
// warp runs K/16 k-steps of three independent MMA chains (the 3xTF32 terms keep their own accumulators), the two
// K halves meet through shared memory in a fixed order.  (The first GEMM form, 64 bodies x 16 outputs per CTA
// with every warp walking the whole K, spent 3 of its 7 us issuing 384 dependent-by-six MMAs per warp.)
__global__ void __launch_bounds__(128) fit_linear_kernel(const LinearParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];     // W [K/32][8][32] | A [K/32][32][32]
    __shared__ __align__(8) uint64_t full[2][kLMaxChunks / 2 / kLGroup];
    __shared__ float s_red[2][32][4];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x, bg = blockIdx.y >> 1, half = blockIdx.y & 1;
    if (bg * kBG + half * kLB >= p.B) return;                      // no body in this half (uniform over the CTA)
    const int nchunks = p.K / kKC, nh = nchunks / 2, ngroups = (nh + kLGroup - 1) / kLGroup;
    const float *__restrict__ w_t = p.W + (size_t)tile * p.K * kLT;                          // [chunk][8][32]
    const float *__restrict__ a_g = p.A + (size_t)bg * p.K * kBG + (size_t)half * (kLB * kKC);   // + chunk * 64 * 32
    unsigned char *sW = smem_raw, *sA = smem_raw + (size_t)nchunks * kLT * kKC * 4;

    if (tid == 0) {
        for (int h = 0; h < 2; ++h)
            for (int i = 0; i < ngroups; ++i) mbar_init(&full[h][i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    if (tid == 0) {
        for (int g = 0; g < ngroups; ++g)
            for (int h = 0; h < 2; ++h) {          // both halves of K start early
                const int c0 = h * nh + g * kLGroup, nc = min(kLGroup, nh - g * kLGroup);
                mbar_arrive_expect_tx(&full[h][g], (uint32_t)(nc * kLChunkBytes));
                tma_load_1d(sW + (size_t)c0 * kLT * kKC * 4, w_t + (size_t)c0 * kLT * kKC, nc * kLT * kKC * 4, &full[h][g]);
                for (int c = c0; c < c0 + nc; ++c)
                    tma_load_1d(sA + (size_t)c * kLB * kKC * 4, a_g + (size_t)c * kBG * kKC, kLB * kKC * 4, &full[h][g]);
            }
    }

    const int mt = w & 1, kh = w >> 1;
    float acc[3][4];           // [term: lo*hi, hi*lo, hi*hi][fragment]
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const int l7 = lane & 7;
    const uint32_t offA = (uint32_t)((mt * 16 + l7 + ((lane >> 3) & 1) * 8) * 128);
    const int kcA = lane >> 4;
    const uint32_t offB = (uint32_t)(l7 * 128);
    const int kcB = (lane >> 3) & 1;
    const uint32_t sW0 = smem_u32(sW), sA0 = smem_u32(sA);

    for (int ci = 0; ci < nh; ++ci) {
        if (ci % kLGroup == 0) mbar_wait(&full[kh][ci / kLGroup], 0);
        const int c = kh * nh + ci;
        const uint32_t sbW = sW0 + c * (kLT * kKC * 4), sbA = sA0 + c * (kLB * kKC * 4);
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {
            const uint32_t ca = (uint32_t)(((2 * s + kcA) ^ l7) << 4), cb = (uint32_t)(((2 * s + kcB) ^ l7) << 4);
            uint32_t a[4], ah[4], al[4], bb[2], bh[2], bl[2];
            ldsm_x4(a, sbA + offA + ca);
            ldsm_x2(bb, sbW + offB + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
            for (int i = 0; i < 2; ++i) split_tf32(bb[i], bh[i], bl[i]);
            mma_tf32(acc[0], al, bh[0], bh[1]); mma_tf32(acc[1], ah, bl[0], bl[1]); mma_tf32(acc[2], ah, bh[0], bh[1]);
        }
    }
    float v4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v4[e] = (acc[0][e] + acc[1][e]) + acc[2][e];       // small terms first
    if (kh == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s_red[mt][lane][e] = v4[e];
    }
    __syncthreads();
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    if (kh == 1) return;
    // accumulator fragment: rows g, g+8 (bodies), columns 2t, 2t+1 (outputs) of the n8 tile
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int b = bg * kBG + half * kLB + mt * 16 + g + 8 * h;
        if (b >= p.B) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = tile * kLT + 2 * t + e;
            float v = v4[2 * h + e] + s_red[mt][lane][2 * h + e];              // first half of K + second half
            if (n < p.n_valid) {
                if (p.bias) v += p.bias[n];
                if (p.pre_out) p.pre_out[(size_t)b * p.n_valid + n] = v;
                if (p.gate) v *= lrelu_grad(p.gate[(size_t)b * p.n_valid + n]);
                if (p.act) v = lrelu(v);
                if (p.out_rm) p.out_rm[(size_t)b * p.ld + n] = v;
            } else {
                v = 0.f;
            }
            if (p.outA && n < p.outA_kpad) p.outA[a_index(b, n, p.outA_kpad)] = v;
        }
    }
}

// host: W[n][k] (row-major [N][K]) -> the kernel's tile layout, N padded to 8, K padded to 32
static std::vector<float> linear_weight_tiles(const float *M, int N, int K, int *Np, int *Kp) {
    const int n16 = (N + kLT - 1) / kLT * kLT, k32 = (K + kKC - 1) / kKC * kKC;
    std::vector<float> out((size_t)n16 * k32, 0.f);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            out[(size_t)(n / kLT) * k32 * kLT + (size_t)(k / kKC) * (kLT * kKC) + (size_t)(n % kLT) * kKC +
                swz(n % kLT, k % kKC)] = M[(size_t)n * K + k];
    *Np = n16;
    *Kp = k32;
    return out;
}

}  // namespace psi
