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
constexpr int kLStages = 16;                             // K <= 512: every stage of a CTA is in flight at once
constexpr int kLStageBytes = (kLT + kBG) * kKC * 4;      // 10 kB (the ring is latency bound: a 4-stage ring
                                                         // measured 8-10 us per layer, 30 kB in flight per SM)

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

__global__ void __launch_bounds__(128) fit_linear_kernel(const LinearParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];     // min(K/32, kLStages) stages
    __shared__ __align__(8) uint64_t full[kLStages], empty[kLStages];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x, bg = blockIdx.y;
    const int nchunks = p.K / kKC;
    const float *__restrict__ w_t = p.W + (size_t)tile * p.K * kLT;      // [chunk][16][32]
    const float *__restrict__ a_g = p.A + (size_t)bg * p.K * kBG;        // [chunk][64][32]

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kLStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {   // thread 0 only
        const int st = c % kLStages;
        mbar_wait(&empty[st], (uint32_t)(((c / kLStages) & 1) ^ 1));
        unsigned char *dst = smem_raw + st * kLStageBytes;
        mbar_arrive_expect_tx(&full[st], (uint32_t)kLStageBytes);
        tma_load_1d(dst, w_t + (size_t)c * kLT * kKC, kLT * kKC * 4, &full[st]);
        tma_load_1d(dst + kLT * kKC * 4, a_g + (size_t)c * kBG * kKC, kBG * kKC * 4, &full[st]);
    };
    if (tid == 0)
        for (int c = 0; c < kLStages - 1 && c < nchunks; ++c) issue(c);

    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const int l7 = lane & 7;
    const uint32_t offA = (uint32_t)(kLT * kKC * 4 + (w * 16 + l7 + ((lane >> 3) & 1) * 8) * 128);
    const int kcA = lane >> 4;
    const uint32_t offB = (uint32_t)((l7 + ((lane >> 4) & 1) * 8) * 128);
    const int kcB = (lane >> 3) & 1;
    const uint32_t smem0 = smem_u32(smem_raw);

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + kLStages - 1 < nchunks) issue(c + kLStages - 1);
        __syncwarp();
        const int st = c % kLStages;
        mbar_wait(&full[st], (uint32_t)((c / kLStages) & 1));
        const uint32_t sb = smem0 + st * kLStageBytes;
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {
            const uint32_t ca = (uint32_t)(((2 * s + kcA) ^ l7) << 4), cb = (uint32_t)(((2 * s + kcB) ^ l7) << 4);
            uint32_t a[4], ah[4], al[4], bb[4], bh[4], bl[4];
            ldsm_x4(a, sb + offA + ca);
            ldsm_x4(bb, sb + offB + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) { split_tf32(a[i], ah[i], al[i]); split_tf32(bb[i], bh[i], bl[i]); }
            mma_3xtf32(acc[0], ah, al, bh[0], bh[1], bl[0], bl[1]);
            mma_3xtf32(acc[1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }

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
                float v = acc[nt][2 * h + e];
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
