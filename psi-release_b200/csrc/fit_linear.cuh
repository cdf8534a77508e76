// VPoser decoder layers (vposer_smpl.py:107-121) and their backward for the fused fitting loop,
// as GEMMs over the WHOLE batch on the tensor cores.
//
// The first fused version ran one CTA per body and streamed every weight matrix through that CTA
// (64 x 1.37 MB of L2 traffic per pass, 37 + 40 us on 64 of the 148 SMs).  Here a layer is
//   Y[b][n] = epilogue( sum_k A[b][k] * W[n][k] )        b < 64 per body group, n < N, k < K
// with the batch as the M dimension: weights are read once per layer, a CTA owns 64 bodies x 16
// outputs (N/16 CTAs), the K loop streams {2 kB weight tile, 8 kB activation tile} stages through
// a full/empty mbarrier ring fed by TMA bulk copies, FP32 operands are split on the fly (3xTF32,
// mma.sync m16n8k8, FP32 accumulate) -- the machinery of the LBS blend GEMM (lbs.cu).
// Both operands use the swizzled chunk layout of common.cuh (a_index / swz); activations are
// written by the producing layer's epilogue directly in that layout.
#pragma once

namespace psi {

constexpr int kLT = 16;                                  // output columns per CTA (2 n8 tiles)
constexpr int kLMaxChunks = 16;                          // K <= 512: both operands of a CTA sit in shared memory
constexpr int kLGroup = 4;                               // chunks per arrival barrier (compute starts on the first 40 kB)
constexpr int kLChunkBytes = (kLT + kBG) * kKC * 4;      // 2 kB of weights + 8 kB of activations per 32 k

struct LinearParams {
    const float *A;        // [body group][K/32][64][32], swizzled; rows of bodies >= B are zero
    const float *W;        // [N/16][K/32][16][32], swizzled: W[n][k], n = output, k = reduction index
    const float *bias;     // [n_valid] or null
    const float *gate;     // row-major [B][n_valid] pre-activations: multiply by lrelu'(gate); or null
    float *pre_out;        // row-major [B][n_valid]: value after the bias (saved for the backward); or null
    float *out_rm;         // row-major out_rm[b*ld + n], n < n_valid; or null
    float *outA;           // next layer's A operand (reduction length outA_kpad >= N); or null
    int K, N, B, n_valid, ld, outA_kpad, act;            // act 1: leaky ReLU 0.2 after the bias
};

// The kernel is latency bound (32 CTAs, 160 kB each): everything is requested up front -- per group
// of 4 chunks ONE bulk copy of weights and ONE of activations (both operands are contiguous over k)
// -- and the three 3xTF32 terms of each output tile accumulate in their own registers, so the
// 64-step K loop is six independent MMA chains per warp instead of two chains of 192.
__global__ void __launch_bounds__(128) fit_linear_kernel(const LinearParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];     // W [K/32][16][32] | A [K/32][64][32]
    __shared__ __align__(8) uint64_t full[kLMaxChunks / kLGroup];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x, bg = blockIdx.y;
    const int nchunks = p.K / kKC, ngroups = (nchunks + kLGroup - 1) / kLGroup;
    const float *__restrict__ w_t = p.W + (size_t)tile * p.K * kLT;      // [chunk][16][32]
    const float *__restrict__ a_g = p.A + (size_t)bg * p.K * kBG;        // [chunk][64][32]
    unsigned char *sW = smem_raw, *sA = smem_raw + (size_t)nchunks * kLT * kKC * 4;

    if (tid == 0) {
        for (int i = 0; i < ngroups; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    if (tid == 0) {
        for (int g = 0; g < ngroups; ++g) {
            const int c0 = g * kLGroup, nc = min(kLGroup, nchunks - c0);
            mbar_arrive_expect_tx(&full[g], (uint32_t)(nc * kLChunkBytes));
            tma_load_1d(sW + (size_t)c0 * kLT * kKC * 4, w_t + (size_t)c0 * kLT * kKC, nc * kLT * kKC * 4, &full[g]);
            tma_load_1d(sA + (size_t)c0 * kBG * kKC * 4, a_g + (size_t)c0 * kBG * kKC, nc * kBG * kKC * 4, &full[g]);
        }
    }

    float acc[2][3][4];        // [n8 tile][term: lo*hi, hi*lo, hi*hi][fragment]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const int l7 = lane & 7;
    const uint32_t offA = (uint32_t)((w * 16 + l7 + ((lane >> 3) & 1) * 8) * 128);
    const int kcA = lane >> 4;
    const uint32_t offB = (uint32_t)((l7 + ((lane >> 4) & 1) * 8) * 128);
    const int kcB = (lane >> 3) & 1;
    const uint32_t sW0 = smem_u32(sW), sA0 = smem_u32(sA);

    for (int c = 0; c < nchunks; ++c) {
        if (c % kLGroup == 0) mbar_wait(&full[c / kLGroup], 0);
        const uint32_t sbW = sW0 + c * (kLT * kKC * 4), sbA = sA0 + c * (kBG * kKC * 4);
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {
            const uint32_t ca = (uint32_t)(((2 * s + kcA) ^ l7) << 4), cb = (uint32_t)(((2 * s + kcB) ^ l7) << 4);
            uint32_t a[4], ah[4], al[4], bb[4], bh[4], bl[4];
            ldsm_x4(a, sbA + offA + ca);
            ldsm_x4(bb, sbW + offB + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) { split_tf32(a[i], ah[i], al[i]); split_tf32(bb[i], bh[i], bl[i]); }
            mma_tf32(acc[0][0], al, bh[0], bh[1]); mma_tf32(acc[0][1], ah, bl[0], bl[1]); mma_tf32(acc[0][2], ah, bh[0], bh[1]);
            mma_tf32(acc[1][0], al, bh[2], bh[3]); mma_tf32(acc[1][1], ah, bl[2], bl[3]); mma_tf32(acc[1][2], ah, bh[2], bh[3]);
        }
    }

    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    // accumulator fragment: rows g, g+8 (bodies), columns 2t, 2t+1 (outputs) of each n8 tile
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int b = bg * kBG + w * 16 + g + 8 * h;
        if (b >= p.B) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = tile * kLT + nt * 8 + 2 * t + e;
                float v = (acc[nt][0][2 * h + e] + acc[nt][1][2 * h + e]) + acc[nt][2][2 * h + e];   // small terms first
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

// host: W[n][k] (row-major [N][K]) -> the kernel's tile layout, N padded to 16, K padded to 32
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
