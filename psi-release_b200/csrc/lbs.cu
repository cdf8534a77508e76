// SMPL-X linear blend skinning, forward and backward, for sm_100a.
//
// Replaces smplx.lbs.lbs (math vendored at human_body_prior/body_model/lbs.py:34-262),
// the translation of body_model.py:246-247 and GeometryTransformer.verts_transform
// (source/cvae.py:141-149), which the reference runs as ~25 torch kernels plus a 54-step
// python loop of 4x4 matmuls.
//
// HBM layout owned by the model handle:
//   basis_fwd [tile n/128][chunk k/32][128 rows][32 k]   rows of k: 0..P-1 = posedirs (P=(J-1)*9), P..P+NB-1 =
//                       shapedirs^T, zero up to Kpad; shape and pose blend shapes become ONE contraction
//   basis_bwd [chunk n/32][Kpad rows][32 n]              the same matrix with the coordinates as the reduction index
//                       (both with 128-byte-swizzled rows: the tcgen05 / ldmatrix operand layout)
//   Jt [J,3], Jdirs [J,3,NB]  joint regressor folded with the template / shape basis
//                       (J = Jt + Jdirs*beta; the reference regresses from 10 475 vertices)
//   skin_j/skin_w [V][KW]  the non-zeros of the dense [V,J] skinning weights (ascending joint order)
//   ch_seg/ch_lv/ch_w, ch_ju/unit_desc   the same non-zeros per 256-vertex chunk, bucketed by joint and cut
//                       into work units, for the backward's dA partial sums
//
// Kernels:
//   lbs_pose_fwd      one CTA per body: Rodrigues / 6D Gram-Schmidt, joints, kinematic chain, A = G - G*J,
//                     blend coefficients (pose feature | beta) as a GEMM operand
//   lbs_blend_fwd_tc5 v_posed = basis * coef on tcgen05 + TMEM (3xTF32)   [lbs_blend_fwd: mma.sync variant]
//   lbs_skin_fwd      sparse skinning + translation + camera transform (+ fused scene-SDF sample)
//   lbs_vertex_bwd    d verts -> d v_posed (GEMM operand) + per-chunk dA partial sums (+ fused loss gradient)
//   lbs_dcoef_tc5     d coef partials = basis^T * gvp on tcgen05 + TMEM        [lbs_dcoef: mma.sync variant]
//   lbs_reduce2       fixed-order sums of the coordinate splits and of the chunk partials
//   lbs_pose_bwd      chain / Rodrigues / Gram-Schmidt backward -> d beta, d pose, d transl, d 6D
// All reductions use fixed orders: results are bit-reproducible run to run.
#include "common.cuh"
#include "rot6d.cuh"
#include "fit_fuse.cuh"
#include "tc5.cuh"
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

#ifndef PSI_LBS_GEMM_DEFAULT
#define PSI_LBS_GEMM_DEFAULT kGemmBf3
#endif

namespace psi {

constexpr int kMaxJ = 64;
constexpr int kJdMaxNB = 32;     // shape coefficients up to which the pose kernels keep Jdirs in shared memory
constexpr int kUnitLen = 15;      // entries per work unit of the dA partial sums (lbs_vertex_bwd); ODD: with a joint-coherent vertex
                                  // order a unit walks consecutive vertices, and units 16 apart would collide on the same banks
constexpr int kMaxUnits = 160;    // units per 256-vertex chunk held in shared memory; more -> per-joint loop
constexpr int kFT = 72;         // forward tile: 72 vertex coordinates (9 n8 tiles); 3V/72 = 437 tiles at
                                // V = 10475 = 2.95 per SM -> one balanced wave at 3 CTAs per SM
constexpr int kFStages = 4;     // forward ring: 4 x 17 kB
constexpr int kDK = 128;        // dcoef tile: 128 coefficients x 64 bodies per CTA
constexpr int kDStages = 4;     // dcoef ring: 4 x 24 kB, 2 CTAs per SM
constexpr int kNSplit = 74;     // dcoef split of the coordinate reduction: 4 k-tiles x 74 = 2 CTAs per SM (mma.sync path)
constexpr int kNSplitTc5 = 37;  // tcgen05 path: 4 k-tiles x 37 = one persistent CTA per SM

}  // namespace psi

struct psi_lbs_tree {          // kinematic tree by levels (root = level 0) + children lists (device)
    int nlev;
    const int *lvl_start, *lvl_joint, *child_start, *child_list;
};

// the same tree BY VALUE (a kernel parameter: constant bank, no global round trip before the pose kernels can start)
struct TreeParam {
    int nlev;
    unsigned char lvl_start[64 + 1], lvl_joint[64], child_start[64 + 1], child_list[64], parents[64];   // 64 = kMaxJ
};

struct psi_lbs_model {
    psi_lbs_tree tree;
    TreeParam tp;
    int *tree_buf;
    int V, J, NB, P, K, Kpad, Npad, KW, NC, NT, FT;   // NC coordinate chunks of 32 (Npad = 32 NC), NT forward tiles of 72
    long nnz;
    float *basis_fwd, *basis_bwd, *v_template, *Jt, *Jdirs, *Jdirs_p, *skin_w, *ch_w;
    unsigned short *basis_fwd3, *basis_bwd3;   // bf3 path: the basis as three bfloat16 terms (see psi_lbs_model_create)
    int *skin_j, *parents, *ch_seg, *ch_ju, *unit_desc;
    int max_units;                 // most work units in one chunk
    unsigned char *ch_lv;
    int max_ent;                   // most skinning entries in one chunk
    int NCH;                       // 256-vertex chunks of the vertex kernels (covers Npad/3 rows)
    size_t bytes;
};

namespace psi {

// The two blend GEMMs, three builds of the same contraction (PSI_LBS_GEMM, read once per process: the model's
// basis layout depends on it):
//   bf3 (default)  tcgen05 + TMEM, kind::f16: every FP32 operand as THREE bfloat16 terms (x = b1 + b2 + b3, exact to
//                  2^-24), the basis pre-split in HBM, the per-body operand pre-split by its producer; the six leading
//                  products come from three MMAs per k step with the per-body terms stacked along N.  A pure
//                  TMA -> MMA pipeline: no thread touches an operand.
//   tc5            tcgen05 + TMEM, kind::tf32, 3xTF32 with the FP32 basis split (hi | lo) in shared memory by four
//                  worker warps.  Measured bound by shared-memory bandwidth: 136 kB through the SM per 16 kB of basis
//                  (profiles/r02i_tc5_pipeline_trace.md).
//   mma            the legacy mma.sync 3xTF32 kernels (tensor-pipe bound; the baseline).
enum { kGemmMma = 0, kGemmTc5 = 1, kGemmBf3 = 2 };
static int lbs_gemm_mode() {
    static const int mode = [] {
        const char *e = getenv("PSI_LBS_GEMM");
        if (e && e[0] == 'm') return (int)kGemmMma;
        if (e && e[0] == 't') return (int)kGemmTc5;
        if (e && e[0] == 'b') return (int)kGemmBf3;
        return (int)PSI_LBS_GEMM_DEFAULT;
    }();
    return mode;
}
static bool lbs_gemm_tc5() { return lbs_gemm_mode() == kGemmTc5; }

struct SavedLayout {
    size_t R, Jr, Gr, Gt, A, vp, coef, coef_lo, total;
};
__host__ __device__ inline SavedLayout saved_layout(int B, int J, int V, int Kpad) {
    SavedLayout s;
    size_t o = 0;
    s.R = o;    o += (size_t)B * J * 9;
    s.Jr = o;   o += (size_t)B * J * 3;
    s.Gr = o;   o += (size_t)B * J * 9;
    s.Gt = o;   o += (size_t)B * J * 3;
    s.A = o;    o += (size_t)B * J * 12;
    o = (o + 3) & ~(size_t)3;
    s.vp = o;   o += (size_t)B * V * 3;
    o = (o + 31) & ~(size_t)31;                 // 128-byte aligned for the bulk copies
    s.coef = o; o += (size_t)((B + kBG - 1) / kBG) * Kpad * kBG;
    s.coef_lo = o; o += (size_t)((B + kBG - 1) / kBG) * Kpad * kBG;   // tcgen05 path: coef = hi, coef_lo = lo (3xTF32 split)
    s.total = o;
    return s;
}

// The kinematic tree (levels, children, parents: constants of the model) staged into shared memory: walking it
// from global memory cost three dependent L2 round trips per tree level (40 % of the pose kernels' stalls).
struct TreeSmem {
    int lvl_start[kMaxJ + 1], lvl_joint[kMaxJ], child_start[kMaxJ + 1], child_list[kMaxJ], parents[kMaxJ];
};
__device__ __forceinline__ void stage_tree(TreeSmem &t, const TreeParam &tp, int J) {
    for (int i = threadIdx.x; i <= J; i += blockDim.x) {
        t.lvl_start[i] = tp.lvl_start[i];
        t.child_start[i] = tp.child_start[i];
        if (i < J) { t.lvl_joint[i] = tp.lvl_joint[i]; t.child_list[i] = tp.child_list[i]; t.parents[i] = tp.parents[i]; }
    }
}

__device__ __forceinline__ void rodrigues(const float *r, float *R) {
    // lbs.py:177-191: eps is added to the vector inside the norm, direction uses raw r
    const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
    const float a = sqrtf(ex * ex + ey * ey + ez * ez);
    const float nx = r[0] / a, ny = r[1] / a, nz = r[2] / a;
    float s, c;
    sincosf(a, &s, &c);
    const float t = 1.0f - c;
    // K = [0 -nz ny; nz 0 -nx; -ny nx 0],  K^2 = n n^T - |n|^2 I
    const float nn = nx * nx + ny * ny + nz * nz;
    R[0] = 1.0f + t * (nx * nx - nn);
    R[1] = -s * nz + t * (nx * ny);
    R[2] = s * ny + t * (nx * nz);
    R[3] = s * nz + t * (nx * ny);
    R[4] = 1.0f + t * (ny * ny - nn);
    R[5] = -s * nx + t * (ny * nz);
    R[6] = -s * ny + t * (nx * nz);
    R[7] = s * nx + t * (ny * nz);
    R[8] = 1.0f + t * (nz * nz - nn);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lbs_pose_fwd_kernel(int J, int NB, int P, int Kpad, const float *__restrict__ Jt,
                    const float *__restrict__ Jdirs, const int *__restrict__ parents, int B,
                    const float *__restrict__ betas, const float *__restrict__ pose,
                    const float *__restrict__ transl, float *__restrict__ saved, SavedLayout L,
                    float *__restrict__ joints_out, const float *__restrict__ rot_in, int num_rot,
                    const float *__restrict__ rot6d, int split, const TreeParam tree,
                    const float *__restrict__ Jdirs_p) {
    __shared__ float sR[kMaxJ * 9], sJ[kMaxJ * 3], sGr[kMaxJ * 9], sGt[kMaxJ * 3];
    __shared__ __align__(16) float sJd[kMaxJ * 3 * (kJdMaxNB + 1)];
    __shared__ float sbeta[kJdMaxNB];
    __shared__ __align__(8) uint64_t jbar;
    __shared__ TreeSmem st;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int S = NB | 1;
    if (Jdirs_p && tid == 0) {                 // constants: requested before the dependency wait
        const uint32_t bytes = (uint32_t)(((size_t)J * 3 * S + 3) / 4 * 16);
        mbar_init(&jbar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&jbar, bytes);
        tma_load_1d(sJd, Jdirs_p, bytes, &jbar);
    }
    stage_tree(st, tree, J);
    float jt[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) jt[i] = tid + i * 128 < J * 3 ? Jt[tid + i * 128] : 0.f;
    pdl_wait();
    if (Jdirs_p && tid < NB) sbeta[tid] = betas[(size_t)b * NB + tid];
    {
        // joint rotations.  Axis-angle joints on the low threads, 6D / matrix joints from thread 64 on: the two code
        // paths run in different warps instead of one after the other
        int j = -1;
        if (tid < J - num_rot) j = num_rot + tid;
        else if (tid >= 64 && tid - 64 < num_rot) j = tid - 64;
        if (j >= 0) {
            if (j < num_rot && rot6d) {   // 6D representation -> R by Gram-Schmidt (cvae.py:46-55), fused here
                float R[9];
                gs_fwd(rot6d + ((size_t)b * num_rot + j) * 6, R);
#pragma unroll
                for (int e = 0; e < 9; ++e) sR[j * 9 + e] = R[e];
            } else if (j < num_rot) {     // rotation matrices handed over directly (no axis-angle round trip)
#pragma unroll
                for (int e = 0; e < 9; ++e) sR[j * 9 + e] = rot_in[((size_t)b * num_rot + j) * 9 + e];
            } else {
                rodrigues(pose + ((size_t)b * J + j) * 3, sR + j * 9);
            }
        }
    }
    if (Jdirs_p) {
        __syncthreads();                       // sbeta, the barrier's initialisation
        mbar_wait(&jbar, 0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int e = tid + i * 128;
            if (e < J * 3) {
                float v = jt[i];
                for (int l = 0; l < NB; ++l) v = fmaf(sJd[e * S + l], sbeta[l], v);
                sJ[e] = v;
            }
        }
    } else {
        for (int e = tid; e < J * 3; e += blockDim.x) {
            float v = Jt[e];
            for (int l = 0; l < NB; ++l) v = fmaf(Jdirs[(size_t)e * NB + l], betas[(size_t)b * NB + l], v);
            sJ[e] = v;
        }
    }
    __syncthreads();
    // kinematic chain, one tree level at a time (joints of a level are independent), one thread per entry of
    // [Gr | Gt] (12 per joint): a level costs one short dependent chain instead of 36 multiply-adds in one thread
    for (int l = 0; l < tree.nlev; ++l) {
        const int q0 = st.lvl_start[l], nj = st.lvl_start[l + 1] - q0;
        for (int u = tid; u < nj * 12; u += blockDim.x) {
            const int j = st.lvl_joint[q0 + u / 12], o = u % 12;
            if (j == 0) {
                if (o < 9) sGr[o] = sR[o];
                else sGt[o - 9] = sJ[o - 9];
                continue;
            }
            const int p = st.parents[j];
            const float *Gp = sGr + p * 9;
            if (o < 9) {
                const int r = o / 3, c = o % 3;
                const float *Rj = sR + j * 9;
                sGr[j * 9 + o] = Gp[r * 3] * Rj[c] + Gp[r * 3 + 1] * Rj[3 + c] + Gp[r * 3 + 2] * Rj[6 + c];
            } else {
                const int r = o - 9;
                const float rel[3] = {sJ[j * 3] - sJ[p * 3], sJ[j * 3 + 1] - sJ[p * 3 + 1], sJ[j * 3 + 2] - sJ[p * 3 + 2]};
                sGt[j * 3 + r] = Gp[r * 3] * rel[0] + Gp[r * 3 + 1] * rel[1] + Gp[r * 3 + 2] * rel[2] + sGt[p * 3 + r];
            }
        }
        __syncthreads();
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    float *oR = saved + L.R + (size_t)b * J * 9, *oJ = saved + L.Jr + (size_t)b * J * 3;
    float *oGr = saved + L.Gr + (size_t)b * J * 9, *oGt = saved + L.Gt + (size_t)b * J * 3;
    float *oA = saved + L.A + (size_t)b * J * 12;
    for (int e = tid; e < J * 9; e += blockDim.x) { oR[e] = sR[e]; oGr[e] = sGr[e]; }
    for (int e = tid; e < J * 3; e += blockDim.x) { oJ[e] = sJ[e]; oGt[e] = sGt[e]; }
    for (int j = tid; j < J; j += blockDim.x) {
        // A = [Gr | Gt - Gr*J]   (lbs.py:257-260), stored row-major 3x4
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float *g = sGr + j * 9 + r * 3;
            oA[j * 12 + r * 4 + 0] = g[0];
            oA[j * 12 + r * 4 + 1] = g[1];
            oA[j * 12 + r * 4 + 2] = g[2];
            oA[j * 12 + r * 4 + 3] =
                sGt[j * 3 + r] - (g[0] * sJ[j * 3] + g[1] * sJ[j * 3 + 1] + g[2] * sJ[j * 3 + 2]);
        }
        if (joints_out) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
                joints_out[((size_t)b * J + j) * 3 + r] =
                    sGt[j * 3 + r] + (transl ? transl[(size_t)b * 3 + r] : 0.f);
        }
    }
    // blend coefficients = the A operand of the blend GEMM: [body group][k chunk][64 bodies][32 k, swizzled]
    float *coef = saved + L.coef + (size_t)(b / kBG) * Kpad * kBG + (size_t)(b % kBG) * kKC;
    for (int k = tid; k < Kpad; k += blockDim.x) {
        float v = 0.f;
        if (k < P) {
            const int j = k / 9 + 1, e = k % 9;
            v = sR[j * 9 + e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
        } else if (k < P + NB) {
            v = Jdirs_p ? sbeta[k - P] : betas[(size_t)b * NB + (k - P)];
        }
        if (split == 2) {     // bf16x3 GEMM: three bfloat16 terms, 192-row tiles (common.cuh)
            unsigned short b1, b2, b3;
            split_bf16x3(v, b1, b2, b3);
            unsigned short *c3 = reinterpret_cast<unsigned short *>(saved + L.coef);
            c3[a3_index(b, 0, k, Kpad)] = b1;
            c3[a3_index(b, 1, k, Kpad)] = b2;
            c3[a3_index(b, 2, k, Kpad)] = b3;
            continue;
        }
        const size_t at = (size_t)(k / kKC) * (kBG * kKC) + swz(b % kBG, k % kKC);
        if (split) {     // tcgen05 GEMM: operands pre-split into TF32 hi + lo (x = hi + lo exactly)
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            coef[at] = hi;
            coef[at + (L.coef_lo - L.coef)] = v - hi;
        } else {
            coef[at] = v;
        }
    }
}

// coefficient rows of bodies past B inside the last body group must read as zero
__global__ void lbs_zero_coef_pad_kernel(float *coef, int B, int Kpad) {
    pdl_wait();
    const int nbg = (B + kBG - 1) / kBG;
    const int first = B % kBG;
    if (first == 0) return;
    float *base = coef + (size_t)(nbg - 1) * Kpad * kBG;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Kpad * kBG; i += gridDim.x * blockDim.x)
        if (((i / kKC) % kBG) >= first) base[i] = 0.f;
}

// the same for the bf16x3 operand: rows of bodies >= B of all three terms, every chunk of the last body group
__global__ void lbs_zero_coef_pad3_kernel(unsigned short *coef3, int B, int Kpad) {
    pdl_wait();
    const int nbg = (B + kBG - 1) / kBG, first = B % kBG;
    if (first == 0) return;
    unsigned short *base = coef3 + (size_t)(nbg - 1) * (Kpad / kKC3) * (3 * kBG * kKC3);
    const int total = (Kpad / kKC3) * 3 * kBG * kKC3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        if (((i / kKC3) % kBG) >= first) base[i] = 0;
}

// ---------------------------------------------------------------------------------------------
struct BlendFwdParams {
    const float *basis_fwd, *v_template, *coef;
    float *vp_out;
    int V, Kpad, B;
};

constexpr int kFStageBytes = (kFT + kBG) * kKC * 4;

// v_posed[b][n] = v_template[n] + sum_k coef[b][k] * basis[k][n]      (lbs.py:88-106)
// as a [64 bodies x 72 coordinates x Kpad] GEMM per CTA on the tensor cores: 4 warps, warp w owns
// bodies 16w..16w+15 (one m16 tile) x 9 n8 tiles; FP32 operands are split on the fly (3xTF32).
// Thread 0 feeds a full/empty mbarrier ring with bulk copies (one 9 kB + one 8 kB block a stage).
__global__ void __launch_bounds__(128, 3) lbs_blend_fwd_kernel(const BlendFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kFStages], empty[kFStages];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x, bg = blockIdx.y;
    const int nchunks = p.Kpad / kKC;
    const float *__restrict__ basis_t = p.basis_fwd + (size_t)tile * p.Kpad * kFT;   // [chunk][72][32]
    const float *__restrict__ coef_g = p.coef + (size_t)bg * p.Kpad * kBG;           // [chunk][64][32]

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kFStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    auto issue = [&](int c) {   // thread 0 only
        const int st = c % kFStages;
        mbar_wait(&empty[st], (uint32_t)(((c / kFStages) & 1) ^ 1));   // first use: passes at once
        unsigned char *dst = smem_raw + st * kFStageBytes;
        mbar_arrive_expect_tx(&full[st], (uint32_t)kFStageBytes);
        tma_load_1d(dst, basis_t + (size_t)c * kFT * kKC, kFT * kKC * 4, &full[st]);
        tma_load_1d(dst + kFT * kKC * 4, coef_g + (size_t)c * kBG * kKC, kBG * kKC * 4, &full[st]);
    };
    if (tid == 0)
        for (int c = 0; c < kFStages - 1 && c < nchunks; ++c) issue(c);

    float acc[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    // ldmatrix lane roles (row = lane & 7 inside an 8-row block, so the swizzle key is lane & 7)
    const int l7 = lane & 7;
    const uint32_t offA = (uint32_t)(kFT * kKC * 4 + (w * 16 + l7 + ((lane >> 3) & 1) * 8) * 128);   // coef rows
    const int kcA = lane >> 4;
    const uint32_t offB = (uint32_t)((l7 + ((lane >> 4) & 1) * 8) * 128);                             // basis rows
    const int kcB = (lane >> 3) & 1;
    const uint32_t offB8 = (uint32_t)((64 + l7) * 128);                                                // 9th n tile
    const uint32_t smem0 = smem_u32(smem_raw);

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + kFStages - 1 < nchunks) issue(c + kFStages - 1);
        __syncwarp();
        const int st = c % kFStages;
        mbar_wait(&full[st], (uint32_t)((c / kFStages) & 1));
        const uint32_t sb = smem0 + st * kFStageBytes;
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {
            const uint32_t ca = (uint32_t)(((2 * s + kcA) ^ l7) << 4), cb = (uint32_t)(((2 * s + kcB) ^ l7) << 4);
            uint32_t a[4], ah[4], al[4];
            ldsm_x4(a, sb + offA + ca);
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bb[4], bh[4], bl[4];
                ldsm_x4(bb, sb + offB + np * (16 * 128) + cb);
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(bb[i], bh[i], bl[i]);
                mma_3xtf32(acc[2 * np], ah, al, bh[0], bh[1], bl[0], bl[1]);
                mma_3xtf32(acc[2 * np + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
            }
            uint32_t b2[2], bh0, bh1, bl0, bl1;
            ldsm_x2(b2, sb + offB8 + cb);
            split_tf32(b2[0], bh0, bl0);
            split_tf32(b2[1], bh1, bl1);
            mma_3xtf32(acc[8], ah, al, bh0, bh1, bl0, bl1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);   // this warp is done with stage st
    }

    // accumulator fragment: rows g, g+8 (bodies), columns 2t, 2t+1 (coordinates) of each n8 tile
    const int g = lane >> 2, t = lane & 3;
    const int N = 3 * p.V;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int b = bg * kBG + w * 16 + g + 8 * h;
        if (b >= p.B) continue;
        float *vp = p.vp_out + (size_t)b * N;
#pragma unroll
        for (int nt = 0; nt < 9; ++nt) {
            const int n = tile * kFT + nt * 8 + 2 * t;
            if (n < N) vp[n] = acc[nt][2 * h] + p.v_template[n];
            if (n + 1 < N) vp[n + 1] = acc[nt][2 * h + 1] + p.v_template[n + 1];
        }
    }
}

// skinning + translation + camera transform, one thread per (body, vertex): a streaming pass over
// v_posed with gathers from the [B,J,12] transform table (169 kB at B=64, cache resident).
// (lbs.py:108-116, body_model.py:246-247, cvae.py:141-149)
// SDF = true (fitting loop): the scene-SDF sample + gradient + collision partial sums of the fresh
// vertex are taken in the same thread (sdf_sample.cuh) -- no second pass over the vertices.
template <bool SDF>
__global__ void __launch_bounds__(256)
lbs_skin_fwd_kernel(int V, int J, int KW, const int *__restrict__ skin_j, const float *__restrict__ skin_w,
                    const float *__restrict__ A, const float *__restrict__ vp_in,
                    const float *__restrict__ transl, const float *__restrict__ cam, long cam_bstride,
                    float *__restrict__ verts, const SdfFuse sf) {
    pdl_wait();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    float neg_sum = 0.f, neg_cnt = 0.f;
    // the body's [J,3,4] transform table (2.6 kB) staged in shared memory: every thread gathers 3 x KW rows of
    // it, and from global memory those scattered 16-byte loads were what the kernel waited on
    __shared__ float4 sA[kMaxJ * 3];
    {
        const float4 *__restrict__ Ag = reinterpret_cast<const float4 *>(A + (size_t)b * J * 12);
        for (int i = threadIdx.x; i < J * 3; i += blockDim.x) sA[i] = Ag[i];
    }
    float x = 0.f, y = 0.f, z = 0.f;
    int4 jj = make_int4(0, 0, 0, 0);
    float4 ww = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < V) {
        const float *vp = vp_in + ((size_t)b * V + v) * 3;
        x = vp[0]; y = vp[1]; z = vp[2];
        if (KW == 4) {     // SMPL-X: <= 4 weights per vertex: one 16-byte load each for ids and weights
            jj = __ldg(reinterpret_cast<const int4 *>(skin_j) + v);
            ww = __ldg(reinterpret_cast<const float4 *>(skin_w) + v);
        }
    }
    __syncthreads();
    if (v < V) {
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
        const float4 *Ab = sA;
        auto blend = [&](int j, float wt, float4 r0, float4 r1, float4 r2) {
            (void)j;
            T[0] = fmaf(wt, r0.x, T[0]); T[1] = fmaf(wt, r0.y, T[1]); T[2] = fmaf(wt, r0.z, T[2]); T[3] = fmaf(wt, r0.w, T[3]);
            T[4] = fmaf(wt, r1.x, T[4]); T[5] = fmaf(wt, r1.y, T[5]); T[6] = fmaf(wt, r1.z, T[6]); T[7] = fmaf(wt, r1.w, T[7]);
            T[8] = fmaf(wt, r2.x, T[8]); T[9] = fmaf(wt, r2.y, T[9]); T[10] = fmaf(wt, r2.z, T[10]); T[11] = fmaf(wt, r2.w, T[11]);
        };
        if (KW == 4) {     // 12 independent shared-memory gathers
            const float4 a0 = (Ab[jj.x * 3]), a1 = (Ab[jj.x * 3 + 1]), a2 = (Ab[jj.x * 3 + 2]);
            const float4 b0 = (Ab[jj.y * 3]), b1 = (Ab[jj.y * 3 + 1]), b2 = (Ab[jj.y * 3 + 2]);
            const float4 c0 = (Ab[jj.z * 3]), c1 = (Ab[jj.z * 3 + 1]), c2 = (Ab[jj.z * 3 + 2]);
            const float4 d0 = (Ab[jj.w * 3]), d1 = (Ab[jj.w * 3 + 1]), d2 = (Ab[jj.w * 3 + 2]);
            blend(jj.x, ww.x, a0, a1, a2); blend(jj.y, ww.y, b0, b1, b2); blend(jj.z, ww.z, c0, c1, c2); blend(jj.w, ww.w, d0, d1, d2);
        } else {
            for (int k = 0; k < KW; ++k) {
                const int j = skin_j[(size_t)v * KW + k];
                const float wt = skin_w[(size_t)v * KW + k];
                blend(j, wt, Ab[j * 3], Ab[j * 3 + 1], Ab[j * 3 + 2]);
            }
        }
        float ox = T[0] * x + T[1] * y + T[2] * z + T[3];
        float oy = T[4] * x + T[5] * y + T[6] * z + T[7];
        float oz = T[8] * x + T[9] * y + T[10] * z + T[11];
        if (transl) {
            ox += transl[(size_t)b * 3]; oy += transl[(size_t)b * 3 + 1]; oz += transl[(size_t)b * 3 + 2];
        }
        if (cam) {
            const float *C = cam + (size_t)b * cam_bstride;
            const float cx = C[0] * ox + C[1] * oy + C[2] * oz + C[3];
            const float cy = C[4] * ox + C[5] * oy + C[6] * oz + C[7];
            const float cz = C[8] * ox + C[9] * oy + C[10] * oz + C[11];
            ox = cx; oy = cy; oz = cz;
        }
        const size_t o = (size_t)b * V + v;
        verts[o * 3] = ox; verts[o * 3 + 1] = oy; verts[o * 3 + 2] = oz;
        if (SDF) {
            float g3[3];
            const float val = sdf_sample(sf.g, ox, oy, oz, g3);
            sf.sdfv[o] = val;
            sf.sdfg[o * 3] = g3[0]; sf.sdfg[o * 3 + 1] = g3[1]; sf.sdfg[o * 3 + 2] = g3[2];
            if (val < 0.f) { neg_sum -= val; neg_cnt += 1.f; }
        }
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    if (SDF) {   // per-(body, chunk) partial sums, fixed order
        __shared__ float red[2][8];
        neg_sum = warp_sum(neg_sum);
        neg_cnt = warp_sum(neg_cnt);
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = neg_sum; red[1][threadIdx.x >> 5] = neg_cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f, c = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { s += red[0][i]; c += red[1][i]; }
            sf.partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = s;
            sf.partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = c;
            if (sf.neg_cnt && c > 0.f) atomicAdd(sf.neg_cnt + (sf.step[0] & 1), (int)c);   // integer: order-independent
            if (sf.body_cnt && c > 0.f) atomicAdd(sf.body_cnt + (size_t)(sf.step[0] & 1) * sf.nbodies + b, (int)c);
        }
    }
}

// The fitting loop's skinning + SDF pass with TWO vertices per thread (v and v + 256 of a 512-vertex block): the
// pass is latency bound (eight scattered grid reads per vertex behind a dependent chain v_posed -> skin -> sample),
// so each thread first issues the loads of both vertices, skins both, puts all sixteen SDF gathers in flight and
// only then interpolates.  Same arithmetic and the same per-256-vertex partial sums as lbs_skin_fwd_kernel<true>.
#ifndef PSI_SKIN2_MINB
#define PSI_SKIN2_MINB 4
#endif
__global__ void __launch_bounds__(256, PSI_SKIN2_MINB)
lbs_skin_sdf2_kernel(int V, int J, const int *__restrict__ skin_j, const float *__restrict__ skin_w,
                     const float *__restrict__ A, const float *__restrict__ vp_in,
                     const float *__restrict__ transl, const float *__restrict__ cam, long cam_bstride,
                     float *__restrict__ verts, const SdfFuse sf) {
    pdl_wait();
    const int b = blockIdx.y, tid = threadIdx.x;
    __shared__ float4 sA[kMaxJ * 3];
    __shared__ float red[2][2][8];
    {
        const float4 *__restrict__ Ag = reinterpret_cast<const float4 *>(A + (size_t)b * J * 12);
        for (int i = tid; i < J * 3; i += blockDim.x) sA[i] = Ag[i];
    }
    int v[2];
    float x[2], y[2], z[2];
    int4 jj[2];
    float4 ww[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        v[h] = blockIdx.x * 512 + h * 256 + tid;
        x[h] = y[h] = z[h] = 0.f;
        jj[h] = make_int4(0, 0, 0, 0);
        ww[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v[h] < V) {
            const float *vp = vp_in + ((size_t)b * V + v[h]) * 3;
            x[h] = vp[0]; y[h] = vp[1]; z[h] = vp[2];
            jj[h] = __ldg(reinterpret_cast<const int4 *>(skin_j) + v[h]);
            ww[h] = __ldg(reinterpret_cast<const float4 *>(skin_w) + v[h]);
        }
    }
    float tr[3] = {0.f, 0.f, 0.f}, C[12];
    if (transl) { tr[0] = transl[(size_t)b * 3]; tr[1] = transl[(size_t)b * 3 + 1]; tr[2] = transl[(size_t)b * 3 + 2]; }
    if (cam) {
#pragma unroll
        for (int e = 0; e < 12; ++e) C[e] = cam[(size_t)b * cam_bstride + e];
    }
    __syncthreads();
    float o[2][3];
    SdfCell cell[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
        auto blend = [&](float wt, float4 r0, float4 r1, float4 r2) {
            T[0] = fmaf(wt, r0.x, T[0]); T[1] = fmaf(wt, r0.y, T[1]); T[2] = fmaf(wt, r0.z, T[2]); T[3] = fmaf(wt, r0.w, T[3]);
            T[4] = fmaf(wt, r1.x, T[4]); T[5] = fmaf(wt, r1.y, T[5]); T[6] = fmaf(wt, r1.z, T[6]); T[7] = fmaf(wt, r1.w, T[7]);
            T[8] = fmaf(wt, r2.x, T[8]); T[9] = fmaf(wt, r2.y, T[9]); T[10] = fmaf(wt, r2.z, T[10]); T[11] = fmaf(wt, r2.w, T[11]);
        };
        const float4 a0 = sA[jj[h].x * 3], a1 = sA[jj[h].x * 3 + 1], a2 = sA[jj[h].x * 3 + 2];
        const float4 b0 = sA[jj[h].y * 3], b1 = sA[jj[h].y * 3 + 1], b2 = sA[jj[h].y * 3 + 2];
        const float4 c0 = sA[jj[h].z * 3], c1 = sA[jj[h].z * 3 + 1], c2 = sA[jj[h].z * 3 + 2];
        const float4 d0 = sA[jj[h].w * 3], d1 = sA[jj[h].w * 3 + 1], d2 = sA[jj[h].w * 3 + 2];
        blend(ww[h].x, a0, a1, a2); blend(ww[h].y, b0, b1, b2); blend(ww[h].z, c0, c1, c2); blend(ww[h].w, d0, d1, d2);
        float ox = T[0] * x[h] + T[1] * y[h] + T[2] * z[h] + T[3];
        float oy = T[4] * x[h] + T[5] * y[h] + T[6] * z[h] + T[7];
        float oz = T[8] * x[h] + T[9] * y[h] + T[10] * z[h] + T[11];
        if (transl) { ox += tr[0]; oy += tr[1]; oz += tr[2]; }
        if (cam) {
            const float cx = C[0] * ox + C[1] * oy + C[2] * oz + C[3];
            const float cy = C[4] * ox + C[5] * oy + C[6] * oz + C[7];
            const float cz = C[8] * ox + C[9] * oy + C[10] * oz + C[11];
            ox = cx; oy = cy; oz = cz;
        }
        o[h][0] = ox; o[h][1] = oy; o[h][2] = oz;
        cell[h] = sdf_prepare(sf.g, ox, oy, oz);
    }
    float sv[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (v[h] < V) sdf_gather(sf.g, cell[h], sv[h]);          // sixteen independent gathers in flight
    float neg_sum[2] = {0.f, 0.f}, neg_cnt[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (v[h] < V) {
            const size_t at = (size_t)b * V + v[h];
            verts[at * 3] = o[h][0]; verts[at * 3 + 1] = o[h][1]; verts[at * 3 + 2] = o[h][2];
            float g3[3];
            const float val = sdf_finish(cell[h], sv[h], g3);
            sf.sdfv[at] = val;
            sf.sdfg[at * 3] = g3[0]; sf.sdfg[at * 3 + 1] = g3[1]; sf.sdfg[at * 3 + 2] = g3[2];
            if (val < 0.f) { neg_sum[h] -= val; neg_cnt[h] += 1.f; }
        }
    }
    pdl_launch_dependents();
    // per-(body, 256-vertex chunk) partial sums, fixed order (the layout lbs_vertex_bwd<FIT> and fit_step read)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float s_ = warp_sum(neg_sum[h]), c_ = warp_sum(neg_cnt[h]);
        if ((tid & 31) == 0) { red[h][0][tid >> 5] = s_; red[h][1][tid >> 5] = c_; }
    }
    __syncthreads();
    if (tid < 2) {
        const int h = tid, np = (V + 255) / 256, chunk = blockIdx.x * 2 + h;
        if (chunk < np) {
            float s_ = 0.f, c_ = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { s_ += red[h][0][i]; c_ += red[h][1][i]; }
            sf.partial[((size_t)b * np + chunk) * 2 + 0] = s_;
            sf.partial[((size_t)b * np + chunk) * 2 + 1] = c_;
            if (sf.neg_cnt && c_ > 0.f) atomicAdd(sf.neg_cnt + (sf.step[0] & 1), (int)c_);   // integer: order-independent
            if (sf.body_cnt && c_ > 0.f) atomicAdd(sf.body_cnt + (size_t)(sf.step[0] & 1) * sf.nbodies + b, (int)c_);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward, vertex side, one CTA per (256-vertex chunk, body):
//   g   = dL/dverts: read from gverts, or (FIT) evaluated here from the NN / SDF results
//   gw  = Rc^T g ;  gvp = Tr^T gw, written as the A operand of the dcoef GEMM
//         ([body group][coordinate chunk][64 bodies][32 coordinates, swizzled]; rows of bodies >= B zero)
//   dA partials of the chunk: dApart[chunk][b][j][:] = sum over the chunk's vertices skinned to j of
//         w * [gw (x) vp | gw]  (row J: sum of gw = d translation).  gw and vp of the chunk sit in
//         shared memory; work item (j, entry e of 12) walks the chunk's entry list of joint j
//         (ch_seg / ch_lv / ch_w, built once per model) in a fixed order.  lbs_pose_bwd adds the
//         chunks up.  (This replaced a per-(joint, body) gather kernel that re-read gw and vp from
//         L2 for every skinning entry: 37 us at B = 64.)
template <bool FIT>
__global__ void __launch_bounds__(256)      // (minBlocks 5 / 6 measured slower: 59.4 vs 56.1 us)
lbs_vertex_bwd_kernel(int V, int J, int KW, int Npad, int B, const int *__restrict__ skin_j,
                      const float *__restrict__ skin_w, const float *__restrict__ A,
                      const float *__restrict__ vp_in, const float *__restrict__ cam, long cam_bstride,
                      const float *__restrict__ gverts, const int *__restrict__ ch_seg,
                      const unsigned char *__restrict__ ch_lv, const float *__restrict__ ch_w, int stage_cap,
                      const int *__restrict__ ch_ju, const int *__restrict__ unit_desc, int use_units,
                      float *__restrict__ gvp_out, float *__restrict__ gvp_lo, int gmode, float *__restrict__ dApart,
                      const VGradFuse fg) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];   // the chunk's entries: stage_cap x (weight, local vertex)
    __shared__ float4 s_gw4[256], s_vp4[256];         // gw and v_posed of the chunk's vertices (xyz, pad): one 16-byte read each
    __shared__ float s_cnt, red[8], red3[24];
    __shared__ float4 s_part[kMaxUnits * 3];          // per (unit, row): partial sums of up to kUnitLen entries
    const int tid = threadIdx.x;
    // the chunk's skinning entries are constants of the model: staged before the dependency wait
    const int *seg = ch_seg + (size_t)blockIdx.x * (J + 1);
    const int e0 = seg[0], nent = seg[J] - e0;
    float *s_w = reinterpret_cast<float *>(dyn_smem);
    unsigned char *s_lv = dyn_smem + (size_t)stage_cap * 4;
    const bool staged = nent <= stage_cap;
    __shared__ __align__(8) uint64_t ent_bar;
    if (staged && nent > 0 && blockIdx.y < B && tid == 0) {
        // two TMA bulk copies instead of a load/store loop through registers (7.5 % of the kernel's stall samples sat
        // on that store); consumed after the block barriers below, so the latency is hidden
        const uint32_t np = (uint32_t)((nent + 15) & ~15);
        mbar_init(&ent_bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&ent_bar, np * 5u);
        tma_load_1d(s_w, ch_w + e0, np * 4u, &ent_bar);
        tma_load_1d(s_lv, ch_lv + e0, np, &ent_bar);
    }
    pdl_wait();
    __shared__ float4 sA[kMaxJ * 3];                  // the body's transform table: gathered 3 x KW rows per thread
    if (blockIdx.y < B) {
        const float4 *__restrict__ Ag = reinterpret_cast<const float4 *>(A + (size_t)blockIdx.y * J * 12);
        for (int i = tid; i < J * 3; i += blockDim.x) sA[i] = Ag[i];
    }
    const int v = blockIdx.x * blockDim.x + tid;
    const int b = blockIdx.y;
    float *gvp_g = gvp_out + (size_t)(b / kBG) * Npad * kBG + (size_t)(b % kBG) * kKC;
    const size_t lo_off = gvp_lo ? (size_t)(gvp_lo - gvp_out) : 0;   // tcgen05 GEMM: operand pre-split into TF32 hi + lo
    // bf16x3 operand: this body's row of term 0 in chunk 0 ([group][chunk n/64][term][64 bodies][64 n])
    unsigned short *g3_body = reinterpret_cast<unsigned short *>(gvp_out) + (size_t)(b / kBG) * (Npad / kKC3) * (3 * kBG * kKC3) +
                              (size_t)(b % kBG) * kKC3;
    auto put = [&](int n, float x) {
        if (n >= Npad) return;
        if (gmode == 2) {        // bf16x3 GEMM operand: three bfloat16 terms (a3_index, common.cuh); Npad is a multiple of 64
            unsigned short b1, b2, b3;
            split_bf16x3(x, b1, b2, b3);
            const int col = n & (kKC3 - 1);
            unsigned short *q = g3_body + (size_t)(n >> 6) * (3 * kBG * kKC3) + (((((col >> 3) ^ b) & 7) << 3) | (col & 7));
            q[0] = b1;
            q[kBG * kKC3] = b2;
            q[2 * kBG * kKC3] = b3;
            return;
        }
        const size_t at = (size_t)(n / kKC) * (kBG * kKC) + swz(b % kBG, n % kKC);
        if (lo_off) {
            const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
            gvp_g[at] = hi;
            gvp_g[at + lo_off] = x - hi;
        } else {
            gvp_g[at] = x;
        }
    };
    if (b >= B) {        // padding rows of the GEMM operand (whole CTA)
        put(3 * v, 0.f); put(3 * v + 1, 0.f); put(3 * v + 2, 0.f);
        return;
    }
    // work-unit tables of the dA partial sums (constants of the model): fetched now, used after the barriers
    const int *ju = ch_ju + (size_t)blockIdx.x * (J + 1);
    int u0 = 0, nunits = 0, d_first = 0, ja = 0, jb = 0;
    if (use_units && staged) {
        u0 = ju[0];
        nunits = ju[J] - u0;
        if (tid < nunits) d_first = unit_desc[u0 + tid];
        if (tid < J * 3) { ja = ju[tid / 3] - u0; jb = ju[tid / 3 + 1] - u0; }
    }
    // The kernel is latency bound: every global load is issued at the earliest point its address is known --
    // level 1 (addresses from v alone) before the block barrier that publishes the body's penetration count,
    // level 2 (NN result of the vertex's contact slot), level 3 (the matched scene point).
    const bool act = v < V;
    const size_t o = (size_t)b * V + (act ? v : 0);
    float sv = 0.f, sg0 = 0.f, sg1 = 0.f, sg2 = 0.f, wgt = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    int slot = -1;
    float px = 0.f, py = 0.f, pz = 0.f;
    int4 jj = make_int4(0, 0, 0, 0);
    float4 ww = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
        const float *vp = vp_in + o * 3;
        px = vp[0]; py = vp[1]; pz = vp[2];
        if (KW == 4) {
            jj = __ldg(reinterpret_cast<const int4 *>(skin_j) + v);
            ww = __ldg(reinterpret_cast<const float4 *>(skin_w) + v);
        }
        if (FIT) {
            sv = fg.sdfv[o];
            sg0 = fg.sdfg[o * 3]; sg1 = fg.sdfg[o * 3 + 1]; sg2 = fg.sdfg[o * 3 + 2];
            slot = fg.cslot[v];
            wgt = fg.cweight[v];
            vx = fg.verts[o * 3]; vy = fg.verts[o * 3 + 1]; vz = fg.verts[o * 3 + 2];
        }
    }
    if (FIT) {
        if (fg.neg_cnt) {    // batch-coupled loss: penetrating vertices of the WHOLE batch (fitting_proxe.py:155-158)
            if (tid == 0) s_cnt = (float)fg.neg_cnt[fg.step[0] & 1];
        } else if (fg.body_cnt) {   // the skinning kernel counted this body's penetrating vertices (same integer, one load)
            if (tid == 0) s_cnt = (float)fg.body_cnt[(size_t)(fg.step[0] & 1) * B + b];
        } else if (tid < 32) {      // number of penetrating vertices of the body: lane-strided loads, shuffle tree (fixed order)
            float c = 0.f;
            for (int i = tid; i < fg.np_sdf; i += 32) c += fg.partial[((size_t)b * fg.np_sdf + i) * 2 + 1];
            c = warp_sum(c);
            if (tid == 0) s_cnt = c;
        }
    }
    float dd = 0.f;
    int ni = 0;
    if (FIT && slot >= 0) {                                    // level 2
        dd = fg.nnd[(size_t)b * fg.nu + slot];
        ni = fg.nni[(size_t)b * fg.nu + slot];
    }
    __syncthreads();
    float gx = 0.f, gy = 0.f, gz = 0.f, closs = 0.f;
    if (act) {
        if (FIT) {
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (slot >= 0) {                                   // level 3
                const float *q = fg.scene + (size_t)ni * 3;
                qx = q[0]; qy = q[1]; qz = q[2];
            }
            if (sv < 0.f) {   // d/dv [ w * sum(-sdf)/cnt ]
                const float k = -fg.w_coll / s_cnt;
                gx = k * sg0;
                gy = k * sg1;
                gz = k * sg2;
            }
            if (slot >= 0) {
                const float s = sqrtf(dd + 1e-4f);
                const float den = s + fg.robust_c;
                closs = wgt * (s / den);
                // d/dd [ s/(s+c) ] = c/(s+c)^2 * 1/(2s);   d dd/dp = 2 (p - q)   (chamfer.cu:165-168)
                const float gd = (fg.w_contact / ((float)fg.num_contact * fg.bdiv)) * wgt * (fg.robust_c / (den * den)) * (0.5f / s);
                const float g2 = gd * 2.0f;
                gx += g2 * (vx - qx);
                gy += g2 * (vy - qy);
                gz += g2 * (vz - qz);
            }
        } else {
            gx = gverts[o * 3]; gy = gverts[o * 3 + 1]; gz = gverts[o * 3 + 2];
        }
        if (cam) {
            const float *C = cam + (size_t)b * cam_bstride;
            const float x = C[0] * gx + C[4] * gy + C[8] * gz;
            const float y = C[1] * gx + C[5] * gy + C[9] * gz;
            const float z = C[2] * gx + C[6] * gy + C[10] * gz;
            gx = x; gy = y; gz = z;
        }
        float T[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) T[e] = 0.f;
        const float4 *Ab = sA;
        auto blend = [&](float wt, float4 r0, float4 r1, float4 r2) {
            T[0] = fmaf(wt, r0.x, T[0]); T[1] = fmaf(wt, r0.y, T[1]); T[2] = fmaf(wt, r0.z, T[2]);
            T[3] = fmaf(wt, r1.x, T[3]); T[4] = fmaf(wt, r1.y, T[4]); T[5] = fmaf(wt, r1.z, T[5]);
            T[6] = fmaf(wt, r2.x, T[6]); T[7] = fmaf(wt, r2.y, T[7]); T[8] = fmaf(wt, r2.z, T[8]);
        };
        if (KW == 4) {     // ids and weights came with the level-1 loads; 12 independent gathers
            const float4 a0 = Ab[jj.x * 3], a1 = Ab[jj.x * 3 + 1], a2 = Ab[jj.x * 3 + 2];
            const float4 b0 = Ab[jj.y * 3], b1 = Ab[jj.y * 3 + 1], b2 = Ab[jj.y * 3 + 2];
            const float4 c0 = Ab[jj.z * 3], c1 = Ab[jj.z * 3 + 1], c2 = Ab[jj.z * 3 + 2];
            const float4 d0 = Ab[jj.w * 3], d1 = Ab[jj.w * 3 + 1], d2 = Ab[jj.w * 3 + 2];
            blend(ww.x, a0, a1, a2); blend(ww.y, b0, b1, b2); blend(ww.z, c0, c1, c2); blend(ww.w, d0, d1, d2);
        } else {
            for (int k = 0; k < KW; ++k) {
                const int j = skin_j[(size_t)v * KW + k];
                const float wt = skin_w[(size_t)v * KW + k];
                blend(wt, Ab[j * 3], Ab[j * 3 + 1], Ab[j * 3 + 2]);
            }
        }
        put(3 * v, T[0] * gx + T[3] * gy + T[6] * gz);
        put(3 * v + 1, T[1] * gx + T[4] * gy + T[7] * gz);
        put(3 * v + 2, T[2] * gx + T[5] * gy + T[8] * gz);
    } else {
        put(3 * v, 0.f); put(3 * v + 1, 0.f); put(3 * v + 2, 0.f);
    }
    s_gw4[tid] = make_float4(gx, gy, gz, 0.f);
    s_vp4[tid] = make_float4(px, py, pz, 0.f);
    {   // per-warp sums of gw (row J of the partials = d translation) and of the contact loss
        const float sx = warp_sum(gx), sy = warp_sum(gy), sz = warp_sum(gz);
        if ((tid & 31) == 0) { red3[(tid >> 5) * 3] = sx; red3[(tid >> 5) * 3 + 1] = sy; red3[(tid >> 5) * 3 + 2] = sz; }
        if (FIT) {
            closs = warp_sum(closs);
            if ((tid & 31) == 0) red[tid >> 5] = closs;
        }
    }
    __syncthreads();
    if (FIT && tid == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        fg.cpart[(size_t)b * gridDim.x + blockIdx.x] = s;
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    // dA partial sums of this chunk: row r of sum over the entries of joint j of (w gw_r) * [vp | 1], entries in
    // ascending vertex order.  Joints own very different numbers of entries (the root collects every
    // vertex skinned to an ancestor chain through it), so the entry lists are cut into UNITS of at most
    // kUnitLen entries (unit_desc, built with the model): phase 1 = one thread per (unit, row), phase 2 =
    // one thread per (joint, row) adds that joint's units in order.
    float *outp = dApart + ((size_t)blockIdx.x * B + b) * (J + 1) * 12;
    if (staged && nent > 0) mbar_wait(&ent_bar, 0);      // the chunk's entries have landed (initialised before the barriers above)
    if (use_units && staged) {
        // one thread per unit does all three rows: the entry's weight, gw and v_posed are read once (4 shared-memory
        // loads per entry instead of 18); every (unit, row) sum keeps its entry order, so the bits do not change
        for (int u = tid; u < nunits; u += blockDim.x) {
            const int d = u == tid ? d_first : unit_desc[u0 + u];
            const int k0 = d >> 8, len = d & 255;
            float4 acc[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int k = k0; k < k0 + len; ++k) {
                const int lv = s_lv[k];
                const float w = s_w[k];
                const float4 gw = s_gw4[lv], vp = s_vp4[lv];
                const float g3[3] = {w * gw.x, w * gw.y, w * gw.z};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    acc[r].x = fmaf(g3[r], vp.x, acc[r].x); acc[r].y = fmaf(g3[r], vp.y, acc[r].y);
                    acc[r].z = fmaf(g3[r], vp.z, acc[r].z); acc[r].w += g3[r];
                }
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) s_part[u * 3 + r] = acc[r];
        }
        __syncthreads();
        for (int item = tid; item < (J + 1) * 3; item += blockDim.x) {
            const int j = item / 3, r = item - j * 3;
            if (j < J) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const int ua = item == tid ? ja : ju[j] - u0, ub = item == tid ? jb : ju[j + 1] - u0;
                for (int u = ua; u < ub; ++u) {
                    const float4 p = s_part[u * 3 + r];
                    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
                }
                *reinterpret_cast<float4 *>(outp + j * 12 + r * 4) = acc;
            } else {
                float a0 = 0.f;
                for (int w8 = 0; w8 < 8; ++w8) a0 += red3[w8 * 3 + r];     // row J: entries 0..2 = sum of gw
                outp[J * 12 + r] = a0;
                if (r == 0)
                    for (int e = 3; e < 12; ++e) outp[J * 12 + e] = 0.f;
            }
        }
        return;
    }
    for (int item = tid; item < (J + 1) * 3; item += blockDim.x) {
        const int j = item / 3, r = item - j * 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (j < J) {
            const int k0 = seg[j] - e0, k1 = seg[j + 1] - e0;
            if (staged) {
#pragma unroll 4
                for (int k = k0; k < k1; ++k) {
                    const int lv = s_lv[k];
                    const float4 vp = s_vp4[lv];
                    const float g = s_w[k] * (&s_gw4[lv].x)[r];
                    a0 = fmaf(g, vp.x, a0); a1 = fmaf(g, vp.y, a1); a2 = fmaf(g, vp.z, a2);
                    a3 += g;
                }
            } else {
#pragma unroll 4
                for (int k = k0; k < k1; ++k) {
                    const int lv = ch_lv[e0 + k];
                    const float4 vp = s_vp4[lv];
                    const float g = ch_w[e0 + k] * (&s_gw4[lv].x)[r];
                    a0 = fmaf(g, vp.x, a0); a1 = fmaf(g, vp.y, a1); a2 = fmaf(g, vp.z, a2);
                    a3 += g;
                }
            }
        } else {
            for (int w8 = 0; w8 < 8; ++w8) a0 += red3[w8 * 3 + r];     // row J: entries 0..2 = sum of gw
        }
        if (j < J) {
            *reinterpret_cast<float4 *>(outp + j * 12 + r * 4) = make_float4(a0, a1, a2, a3);
        } else {
            outp[J * 12 + r] = a0;
            if (r == 0)
                for (int e = 3; e < 12; ++e) outp[J * 12 + e] = 0.f;
        }
    }
}

// d coef partials: part[ns][b][k] = sum_{n in split ns} gvp[b][n] * basis[k][n]   (backward of the blend GEMM)
// CTA = 64 bodies x 128 coefficients over a range of coordinate chunks, on the tensor cores
// (3xTF32): 8 warps, warp (wm, wn) owns 32 bodies x 32 coefficients = 2 m16 x 4 n8 tiles.
constexpr int kDStageBytes = (kDK + kBG) * kKC * 4;

__global__ void __launch_bounds__(256, 2)
lbs_dcoef_kernel(int Kpad, int NC, int Bpad, const float *__restrict__ basis_bwd,
                 const float *__restrict__ gvp, float *__restrict__ part, int nsplit) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kDStages], empty[kDStages];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, wm = w & 1, wn = w >> 1;
    const int kt = blockIdx.x, ns = blockIdx.y, bg = blockIdx.z;
    const int c_begin = (int)((long)ns * NC / nsplit), c_end = (int)((long)(ns + 1) * NC / nsplit);
    const int nchunks = c_end - c_begin;
    const float *__restrict__ gvp_g = gvp + (size_t)bg * NC * (kBG * kKC);      // [chunk][64][32]
    const float *__restrict__ basis_k = basis_bwd + (size_t)kt * kDK * kKC;     // [chunk][Kpad][32]

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kDStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    auto issue = [&](int c) {   // thread 0 only; c counts from 0 inside this split
        const int st = c % kDStages;
        mbar_wait(&empty[st], (uint32_t)(((c / kDStages) & 1) ^ 1));
        unsigned char *dst = smem_raw + st * kDStageBytes;
        mbar_arrive_expect_tx(&full[st], (uint32_t)kDStageBytes);
        tma_load_1d(dst, basis_k + (size_t)(c_begin + c) * Kpad * kKC, kDK * kKC * 4, &full[st]);
        tma_load_1d(dst + kDK * kKC * 4, gvp_g + (size_t)(c_begin + c) * (kBG * kKC), kBG * kKC * 4, &full[st]);
    };
    if (tid == 0)
        for (int c = 0; c < kDStages - 1 && c < nchunks; ++c) issue(c);

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    const int l7 = lane & 7;
    const uint32_t offA = (uint32_t)(kDK * kKC * 4 + (wm * 32 + l7 + ((lane >> 3) & 1) * 8) * 128);   // gvp rows (bodies)
    const int kcA = lane >> 4;
    const uint32_t offB = (uint32_t)((wn * 32 + l7 + ((lane >> 4) & 1) * 8) * 128);                    // basis rows (k)
    const int kcB = (lane >> 3) & 1;
    const uint32_t smem0 = smem_u32(smem_raw);

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + kDStages - 1 < nchunks) issue(c + kDStages - 1);
        __syncwarp();
        const int st = c % kDStages;
        mbar_wait(&full[st], (uint32_t)((c / kDStages) & 1));
        const uint32_t sb = smem0 + st * kDStageBytes;
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {
            const uint32_t ca = (uint32_t)(((2 * s + kcA) ^ l7) << 4), cb = (uint32_t)(((2 * s + kcB) ^ l7) << 4);
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                uint32_t a[4];
                ldsm_x4(a, sb + offA + mt * (16 * 128) + ca);
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[mt][i], al[mt][i]);
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t bb[4], bh[4], bl[4];
                ldsm_x4(bb, sb + offB + np * (16 * 128) + cb);
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(bb[i], bh[i], bl[i]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_3xtf32(acc[mt][2 * np], ah[mt], al[mt], bh[0], bh[1], bl[0], bl[1]);
                    mma_3xtf32(acc[mt][2 * np + 1], ah[mt], al[mt], bh[2], bh[3], bl[2], bl[3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }

    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = bg * kBG + wm * 32 + mt * 16 + g + 8 * h;
            float *o = part + ((size_t)ns * Bpad + b) * Kpad + kt * kDK + wn * 32 + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<float2 *>(o + nt * 8) = make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
        }
}

// ---------------------------------------------------------------------------------------------
// The two blend GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Both are D[128 x 64 bodies] = BASIS_TILE[128 x K] * ACT[64 x K]^T with the basis as the M = 128
// operand (full-rate datapath, cta_group::1) and the per-body operand (blend coefficients, or the
// vertex gradient gvp) as N = 64; the accumulator lives in tensor memory (64 columns), operands are
// read by the tensor core straight from 128-byte-swizzled shared-memory tiles (tc5.cuh).
//   forward : rows = 128 vertex coordinates, K = blend coefficients  -> v_posed
//   backward: rows = 128 blend coefficients, K = a range of vertex coordinates -> d coef partials
// 3xTF32: the per-body operand arrives pre-split from its producer (hi and lo arrays); the basis tile
// is split in shared memory by the CTA's 128 worker threads (raw -> hi | lo) while the previous
// chunk's MMAs run; per 8-wide k step the elected thread issues hi*lo, lo*hi, hi*hi into the same
// accumulator.  Three rings decouple the latencies: RAW basis tiles (5 x 16 kB, HBM stream, freed by
// the workers right after the split), ACT tiles (4 x 16 kB, L2, freed by tcgen05.commit) and the
// OPERAND slots the split writes (2 x 32 kB, freed by tcgen05.commit).  Warp 4 / warp 5 (one lane
// each) are the TMA producers of the RAW / ACT rings.  CTAs are persistent over work items; the
// epilogue reads TMEM with tcgen05.ld: lane = row, so a warp stores 32 consecutive floats per body.
// (mma.sync kept these GEMMs tensor-pipe bound at 41 / 37 us.)
constexpr int kTM = 128;                       // rows of a basis tile = M of the MMA
#ifndef PSI_TC5_RAW
#define PSI_TC5_RAW 5
#endif
#ifndef PSI_TC5_ACT
#define PSI_TC5_ACT 4
#endif
#ifndef PSI_TC5_OP
#define PSI_TC5_OP 2
#endif
constexpr int kTRaw = PSI_TC5_RAW, kTAct = PSI_TC5_ACT, kTOp = PSI_TC5_OP;  // ring depths
#ifndef PSI_TC5_WORKERS
#define PSI_TC5_WORKERS 4
#endif
constexpr int kTWarps = PSI_TC5_WORKERS;       // worker warps (4 or 8): split the raw basis tile, read the accumulator
constexpr int kTWork = kTWarps * 32;           // worker threads
constexpr int kTThreads = kTWork + 96;         // + 2 producer warps + the MMA issuer's warp
static_assert(kTWarps == 4 || kTWarps == 8, "a TMEM lane quarter is owned by warp % 4: 4 or 8 worker warps");
constexpr int kTTileB = kTM * kKC * 4;         // 16 kB: one basis tile (raw, hi or lo)
constexpr int kTTileA = kBG * kKC * 4;         // 8 kB: one per-body tile (hi or lo)
constexpr int kTSmem = kTRaw * kTTileB + kTAct * 2 * kTTileA + kTOp * 2 * kTTileB;   // 208 kB
constexpr int kTCols = 128;                    // TMEM columns: [0,64) = hi*hi + lo*hi, [64,128) = hi*lo

struct Tc5Bars {
    uint64_t full_raw[kTRaw], empty_raw[kTRaw], full_act[kTAct], empty_act[kTAct], full_op[kTOp], empty_op[kTOp];
    uint64_t accum, tmem_free;
    uint32_t tmem_slot;
};

// -DPSI_TC5_TRACE: per-chunk time stamps of CTA 0's pipeline roles (debug builds only; tools/tc5_trace.py)
#ifdef PSI_TC5_TRACE
__device__ unsigned long long g_tc5_trace[8 * 64];      // [event][chunk]: globaltimer ns
__device__ __forceinline__ unsigned long long tc5_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TC5_STAMP(ev, g) do { if (blockIdx.x == 0 && (g) < 64) g_tc5_trace[(ev) * 64 + (g)] = tc5_now(); } while (0)
#else
#define TC5_STAMP(ev, g) do { } while (0)
#endif

struct Tc5Item {                 // one accumulator's worth of work
    const float *basis;          // first chunk of the [chunk][128][32] tile stream
    size_t basis_stride;         // floats between chunks
    const float *act_hi, *act_lo;   // [chunk][64][32]
    int nchunks;
};

// Runs the K loop of `it`; g0 = number of chunks this CTA has processed before (ring phase), item_idx =
// number of items before.  Roles by warp: 0-3 workers (split the raw basis tile into hi | lo operand
// tiles), 4 raw-tile producer, 5 per-body-tile producer, 6 MMA issuer.  Worker threads return with the
// accumulator complete in TMEM; after reading it they must arrive on bars.tmem_free.
__device__ __forceinline__ void tc5_mainloop(unsigned char *smem, Tc5Bars &bars, uint32_t tmem_d, const Tc5Item &it,
                                             int g0, int item_idx) {
    const int tid = threadIdx.x;
    unsigned char *raw = smem, *act = smem + kTRaw * kTTileB, *op = act + kTAct * 2 * kTTileA;
    if (tid >= kTWork) {
        if (tid == kTWork) {                            // ---- producer: raw basis tiles (HBM), streamed evict-first
            const uint64_t pol = l2_policy_evict_first();
            for (int c = 0; c < it.nchunks; ++c) {
                const int g = g0 + c, st = g % kTRaw;
                mbar_wait(&bars.empty_raw[st], (uint32_t)(((g / kTRaw) & 1) ^ 1));
                TC5_STAMP(0, g);                          // raw producer: slot free, TMA issued now
                mbar_arrive_expect_tx(&bars.full_raw[st], (uint32_t)kTTileB);
                tma_load_1d_hint(raw + (size_t)st * kTTileB, it.basis + (size_t)c * it.basis_stride, kTTileB,
                                 &bars.full_raw[st], pol);
            }
        } else if (tid == kTWork + 32) {                // ---- producer: per-body tiles (L2); hi then lo = one 128-row tile
            for (int c = 0; c < it.nchunks; ++c) {
                const int g = g0 + c, st = g % kTAct;
                mbar_wait(&bars.empty_act[st], (uint32_t)(((g / kTAct) & 1) ^ 1));
                mbar_arrive_expect_tx(&bars.full_act[st], (uint32_t)(2 * kTTileA));
                tma_load_1d(act + (size_t)st * 2 * kTTileA, it.act_hi + (size_t)c * (kBG * kKC), kTTileA, &bars.full_act[st]);
                tma_load_1d(act + (size_t)st * 2 * kTTileA + kTTileA, it.act_lo + (size_t)c * (kBG * kKC), kTTileA,
                            &bars.full_act[st]);
            }
        } else if (tid == kTWork + 64) {                // ---- MMA issuer
            constexpr uint32_t idesc_n128 = tc5::idesc_tf32(kTM, 2 * kBG), idesc_n64 = tc5::idesc_tf32(kTM, kBG);
            mbar_wait(&bars.tmem_free, (uint32_t)((item_idx & 1) ^ 1));      // the previous item's epilogue has read TMEM
            tc5::fence_after_sync();
            for (int c = 0; c < it.nchunks; ++c) {
                const int g = g0 + c, so = g % kTOp, sa = g % kTAct;
                mbar_wait(&bars.full_op[so], (uint32_t)((g / kTOp) & 1));
                TC5_STAMP(4, g);                          // issuer: operand slot ready
                mbar_wait(&bars.full_act[sa], (uint32_t)((g / kTAct) & 1));
                TC5_STAMP(5, g);                          // issuer: per-body tile ready
                tc5::fence_after_sync();
                const uint32_t sop = smem_u32(op + (size_t)so * 2 * kTTileB), sac = smem_u32(act + (size_t)sa * 2 * kTTileA);
                const uint64_t dbh = tc5::smem_desc_k_sw128(sop), dbl = tc5::smem_desc_k_sw128(sop + kTTileB);
                const uint64_t dact = tc5::smem_desc_k_sw128(sac);      // rows 0..63 = hi, 64..127 = lo
#pragma unroll
                for (int k = 0; k < kKC / 8; ++k) {      // 32 bytes of every 128-byte row per step: start address + 2
                    tc5::mma_tf32(tmem_d, dbh + 2 * k, dact + 2 * k, idesc_n128, (c | k) != 0);   // hi*[hi | lo]
                    tc5::mma_tf32(tmem_d, dbl + 2 * k, dact + 2 * k, idesc_n64, 1u);              // lo*hi
                }
                tc5::commit(&bars.empty_op[so]);
                tc5::commit(&bars.empty_act[sa]);
                TC5_STAMP(6, g);                          // issuer: MMAs + commits issued
                if (c == it.nchunks - 1) tc5::commit(&bars.accum);
            }
        }
        return;
    }
    for (int c = 0; c < it.nchunks; ++c) {
        const int g = g0 + c, sr = g % kTRaw, so = g % kTOp;
        mbar_wait(&bars.full_raw[sr], (uint32_t)((g / kTRaw) & 1));
        if (tid == 0) TC5_STAMP(1, g);                  // worker: raw tile landed
        mbar_wait(&bars.empty_op[so], (uint32_t)(((g / kTOp) & 1) ^ 1));       // the MMAs that read this slot are done
        if (tid == 0) TC5_STAMP(2, g);                  // worker: operand slot free
        const uint4 *src = reinterpret_cast<const uint4 *>(raw + (size_t)sr * kTTileB);
        uint4 *hi = reinterpret_cast<uint4 *>(op + (size_t)so * 2 * kTTileB);
        float4 *lo = reinterpret_cast<float4 *>(op + (size_t)so * 2 * kTTileB + kTTileB);
#pragma unroll 4
        for (int e = tid; e < kTM * kKC / 4; e += kTWork) {   // element-wise: the swizzle does not matter
            const uint4 x = src[e];
            const uint4 h = make_uint4(x.x & 0xffffe000u, x.y & 0xffffe000u, x.z & 0xffffe000u, x.w & 0xffffe000u);
            hi[e] = h;
            lo[e] = make_float4(__uint_as_float(x.x) - __uint_as_float(h.x), __uint_as_float(x.y) - __uint_as_float(h.y),
                                __uint_as_float(x.z) - __uint_as_float(h.z), __uint_as_float(x.w) - __uint_as_float(h.w));
        }
        tc5::fence_proxy_async();
        if (tid == 0) TC5_STAMP(3, g);                  // worker: split + proxy fence done
        mbar_arrive(&bars.full_op[so]);                 // kTWork arrivals: the operand slot is ready for the issuer
        mbar_arrive(&bars.empty_raw[sr]);               // kTWork arrivals free the raw slot
    }
    mbar_wait(&bars.accum, (uint32_t)(item_idx & 1));
    tc5::fence_after_sync();
}

__device__ __forceinline__ uint32_t tc5_setup(Tc5Bars &bars) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTRaw; ++i) { mbar_init(&bars.full_raw[i], 1); mbar_init(&bars.empty_raw[i], kTWork); }
        for (int i = 0; i < kTAct; ++i) { mbar_init(&bars.full_act[i], 1); mbar_init(&bars.empty_act[i], 1); }
        for (int i = 0; i < kTOp; ++i) { mbar_init(&bars.full_op[i], kTWork); mbar_init(&bars.empty_op[i], 1); }
        mbar_init(&bars.accum, 1);
        mbar_init(&bars.tmem_free, kTWork);
        mbar_fence_init();
    }
    if (threadIdx.x >= kTWork + 32 && threadIdx.x < kTWork + 64) {   // the second producer warp owns the TMEM allocation
        tc5::tmem_alloc(&bars.tmem_slot, kTCols);
        tc5::tmem_relinquish();
    }
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    return bars.tmem_slot;
}

__device__ __forceinline__ void tc5_teardown(uint32_t tmem_d) {
    tc5::fence_before_sync();
    __syncthreads();
    if (threadIdx.x >= kTWork + 32 && threadIdx.x < kTWork + 64) tc5::tmem_dealloc(tmem_d, kTCols);
}

// accumulator row of this lane, bodies c8*8 .. c8*8+7: hi*hi + lo*hi (columns 0..63) + hi*lo (columns 64..127)
__device__ __forceinline__ void tc5_load8(uint32_t tmem_d, int w, int c8, float (&v)[8]) {
    float u[8];
    tc5::ld_32x32b_x8(tmem_d + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)(c8 * 8), v);      // a warp reads lane quarter w % 4
    tc5::ld_32x32b_x8(tmem_d + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)(kBG + c8 * 8), u);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += u[e];
}

struct BlendFwdTc5Params {
    const float *basis_fwd, *v_template, *coef_hi, *coef_lo;
    float *vp_out;
    int V, Kpad, B, ntiles, nbg;
};

// v_posed[b][n] = v_template[n] + sum_k coef[b][k] * basis[k][n]; work item = (128 coordinates, body group)
__global__ void __launch_bounds__(kTThreads, 1) lbs_blend_fwd_tc5_kernel(const BlendFwdTc5Params p) {
    extern __shared__ __align__(1024) unsigned char smem_tc5[];
    unsigned char *smem_raw = smem_tc5 + ((1024u - (smem_u32(smem_tc5) & 1023u)) & 1023u);   // swizzle atoms need 1024-byte alignment
    __shared__ __align__(8) Tc5Bars bars;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t tmem_d = tc5_setup(bars);
    pdl_wait();
    const int nchunks = p.Kpad / kKC, N = 3 * p.V;
    int g0 = 0, idx = 0;
    for (int item = blockIdx.x; item < p.ntiles * p.nbg; item += gridDim.x, g0 += nchunks, ++idx) {
        const int tile = item % p.ntiles, bg = item / p.ntiles;
        Tc5Item it;
        it.basis = p.basis_fwd + (size_t)tile * p.Kpad * kTM;
        it.basis_stride = (size_t)kTM * kKC;
        it.act_hi = p.coef_hi + (size_t)bg * p.Kpad * kBG;
        it.act_lo = p.coef_lo + (size_t)bg * p.Kpad * kBG;
        it.nchunks = nchunks;
        tc5_mainloop(smem_raw, bars, tmem_d, it, g0, idx);
        if (w < kTWarps) {
            const int n = tile * kTM + (w & 3) * 32 + lane;    // TMEM lane = accumulator row = coordinate
            const float vt = n < N ? p.v_template[n] : 0.f;
            // with 8 worker warps, warps w and w + 4 share a lane quarter and take half of the bodies each
            constexpr int kC8 = kBG / 8 / (kTWarps / 4);
#pragma unroll 2
            for (int c8 = (w >> 2) * kC8; c8 < (w >> 2) * kC8 + kC8; ++c8) {
                float v[8];
                tc5_load8(tmem_d, w, c8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int b = bg * kBG + c8 * 8 + e;
                    if (n < N && b < p.B) p.vp_out[(size_t)b * N + n] = v[e] + vt;
                }
            }
            tc5::fence_before_sync();
            mbar_arrive(&bars.tmem_free);                      // TMEM has been read: the next item may overwrite it
        }
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    tc5_teardown(tmem_d);
}

// part[ns][b][k] = sum_{n in split ns} gvp[b][n] * basis[k][n]; work item = (128 coefficients, split, body group)
__global__ void __launch_bounds__(kTThreads, 1)
lbs_dcoef_tc5_kernel(int Kpad, int NC, int Bpad, const float *__restrict__ basis_bwd, const float *__restrict__ gvp_hi,
                     const float *__restrict__ gvp_lo, float *__restrict__ part, int nsplit) {
    extern __shared__ __align__(1024) unsigned char smem_tc5[];
    unsigned char *smem_raw = smem_tc5 + ((1024u - (smem_u32(smem_tc5) & 1023u)) & 1023u);
    __shared__ __align__(8) Tc5Bars bars;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t tmem_d = tc5_setup(bars);
    pdl_wait();
    const int nkt = Kpad / kTM, nbg = Bpad / kBG;
    int g0 = 0, idx = 0;
    for (int item = blockIdx.x; item < nkt * nsplit * nbg; item += gridDim.x) {
        const int kt = item % nkt, ns = (item / nkt) % nsplit, bg = item / (nkt * nsplit);
        const int c_begin = (int)((long)ns * NC / nsplit), c_end = (int)((long)(ns + 1) * NC / nsplit);
        float *o = part + ((size_t)ns * Bpad + (size_t)bg * kBG) * Kpad + kt * kTM + (w & 3) * 32 + lane;
        if (c_end == c_begin) {                    // more splits than chunks (small models): an empty split
            if (w < 4)
                for (int b = 0; b < kBG; ++b) o[(size_t)b * Kpad] = 0.f;
            continue;
        }
        Tc5Item it;
        it.basis = basis_bwd + (size_t)c_begin * Kpad * kKC + (size_t)kt * kTM * kKC;
        it.basis_stride = (size_t)Kpad * kKC;
        it.act_hi = gvp_hi + (size_t)bg * NC * (kBG * kKC) + (size_t)c_begin * (kBG * kKC);
        it.act_lo = gvp_lo + (size_t)bg * NC * (kBG * kKC) + (size_t)c_begin * (kBG * kKC);
        it.nchunks = c_end - c_begin;
        tc5_mainloop(smem_raw, bars, tmem_d, it, g0, idx);
        g0 += it.nchunks;
        ++idx;
        if (w < kTWarps) {
            constexpr int kC8 = kBG / 8 / (kTWarps / 4);
#pragma unroll 2
            for (int c8 = (w >> 2) * kC8; c8 < (w >> 2) * kC8 + kC8; ++c8) {
                float v[8];
                tc5_load8(tmem_d, w, c8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[(size_t)(c8 * 8 + e) * Kpad] = v[e];     // lane = coefficient: coalesced per body
            }
            tc5::fence_before_sync();
            mbar_arrive(&bars.tmem_free);
        }
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    tc5_teardown(tmem_d);
}

// ---------------------------------------------------------------------------------------------
// The blend GEMMs as BF16x3 on tcgen05 + TMEM (kind::f16) -- the default path.
//
// Why: the 3xTF32 kernels above are bound by shared-memory bandwidth (the FP32 basis tile is split into hi | lo by
// worker warps: 136 kB move through the SM per 16 kB of basis).  Here every FP32 value is THREE bfloat16 terms,
// x = b1 + b2 + b3 (8 + 8 + 8 mantissa bits, exact to 2^-24).  The basis is split once at model load and streamed
// from HBM as bf16 (6 bytes per element instead of 4: the price of having no thread touch an operand), the per-body
// operand is written pre-split by its producer (lbs_pose_fwd / lbs_vertex_bwd).  Of the nine term products the six
// with weight >= 2^-16 are kept -- a1c1, a1c2, a1c3, a2c1, a2c2, a3c1 (the dropped ones are <= 2^-24) -- and they
// cost three MMAs per 16-wide k step because the per-body terms are stacked along N:
//     a1 x [c1 | c2 | c3]  (N = 192)      a2 x [c1 | c2]  (N = 128)      a3 x [c1]  (N = 64)
// all into one accumulator of 192 TMEM columns; the epilogue adds the three 64-column blocks.  bf16 x bf16 products
// are exact in FP32 and the accumulation is FP32, so the result carries FP32 accuracy.
// Pipeline: one producer thread (TMA: 48 kB of basis terms + 24 kB of per-body terms per 64-k stage, 3 stages),
// one MMA issuer thread, four epilogue warps; TWO accumulators (2 x 256 of the 512 TMEM columns), so the
// tcgen05.ld epilogue of item i runs under the main loop of item i + 1.
constexpr int kT3Stages = 3;
constexpr int kT3TileA = kTM * kKC3 * 2;            // 16 kB: one basis term tile, 128 rows x 64 bf16
constexpr int kT3TileB = 3 * kBG * kKC3 * 2;        // 24 kB: per-body tile, 3 terms x 64 bodies x 64 bf16
constexpr int kT3Stage = 3 * kT3TileA + kT3TileB;   // 72 kB
constexpr int kT3Smem = kT3Stages * kT3Stage;       // 216 kB
constexpr int kT3Threads = 192;                     // 4 epilogue warps + the producer's warp + the issuer's warp
constexpr int kT3Cols = 512;                        // two accumulators, 256 columns apart (192 used)

struct T3Bars {
    uint64_t full[kT3Stages], empty[kT3Stages], accum[2], tmem_free[2];
    uint32_t tmem_slot;
};
struct T3Item {
    const unsigned short *basis;     // first 48 kB stage block [3 terms][128 rows][64]
    size_t basis_stride;             // elements between consecutive chunks
    const unsigned short *act;       // first 24 kB per-body block [3 terms][64 bodies][64]; consecutive chunks are contiguous
    int nchunks;
};

__device__ __forceinline__ uint32_t t3_setup(T3Bars &bars) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < kT3Stages; ++i) { mbar_init(&bars.full[i], 1); mbar_init(&bars.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars.accum[i], 1); mbar_init(&bars.tmem_free[i], 128); }
        mbar_fence_init();
    }
    if (threadIdx.x >= 128 && threadIdx.x < 160) {      // the producer's warp owns the TMEM allocation
        tc5::tmem_alloc(&bars.tmem_slot, kT3Cols);
        tc5::tmem_relinquish();
    }
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    return bars.tmem_slot;
}
__device__ __forceinline__ void t3_teardown(uint32_t tmem_d) {
    tc5::fence_before_sync();
    __syncthreads();
    if (threadIdx.x >= 128 && threadIdx.x < 160) tc5::tmem_dealloc(tmem_d, kT3Cols);
}

// K loop of one item.  g0 = chunks this CTA has streamed before (ring phase), idx = items before (accumulator idx & 1).
// The producer and the issuer return as soon as their work is ISSUED; the epilogue warps return with the
// accumulator complete and must arrive on bars.tmem_free[idx & 1] after reading it.
__device__ __forceinline__ void t3_mainloop(unsigned char *smem, T3Bars &bars, uint32_t tmem_d, const T3Item &it, int g0, int idx) {
    const int tid = threadIdx.x, acc = idx & 1;
    if (tid == 128) {                                   // ---- TMA producer
        const uint64_t pol = l2_policy_evict_first();
        for (int c = 0; c < it.nchunks; ++c) {
            const int g = g0 + c, st = g % kT3Stages;
            mbar_wait(&bars.empty[st], (uint32_t)(((g / kT3Stages) & 1) ^ 1));
            TC5_STAMP(0, g);                          // producer: stage free, TMA issued now
            mbar_arrive_expect_tx(&bars.full[st], (uint32_t)kT3Stage);
            unsigned char *dst = smem + (size_t)st * kT3Stage;
#ifdef PSI_T3_NO_HINT
            (void)pol;
            tma_load_1d(dst, it.basis + (size_t)c * it.basis_stride, 3 * kT3TileA, &bars.full[st]);
#else
            tma_load_1d_hint(dst, it.basis + (size_t)c * it.basis_stride, 3 * kT3TileA, &bars.full[st], pol);   // HBM stream
#endif
            tma_load_1d(dst + 3 * kT3TileA, it.act + (size_t)c * (3 * kBG * kKC3), kT3TileB, &bars.full[st]);   // L2
        }
    } else if (tid == 160) {                            // ---- MMA issuer
        constexpr uint32_t i192 = tc5::idesc_bf16(kTM, 3 * kBG), i128 = tc5::idesc_bf16(kTM, 2 * kBG), i64 = tc5::idesc_bf16(kTM, kBG);
        mbar_wait(&bars.tmem_free[acc], (uint32_t)(((idx >> 1) & 1) ^ 1));     // the epilogue two items ago has read this accumulator
        tc5::fence_after_sync();
        const uint32_t d = tmem_d + (uint32_t)(acc * 256);
        for (int c = 0; c < it.nchunks; ++c) {
            const int g = g0 + c, st = g % kT3Stages;
            mbar_wait(&bars.full[st], (uint32_t)((g / kT3Stages) & 1));
            TC5_STAMP(4, g);                          // issuer: stage landed
            tc5::fence_after_sync();
            const uint32_t sb = smem_u32(smem + (size_t)st * kT3Stage);
            const uint64_t a1 = tc5::smem_desc_k_sw128(sb), a2 = tc5::smem_desc_k_sw128(sb + kT3TileA),
                           a3 = tc5::smem_desc_k_sw128(sb + 2 * kT3TileA), bc = tc5::smem_desc_k_sw128(sb + 3 * kT3TileA);
#pragma unroll
            for (int k = 0; k < kKC3 / 16; ++k) {       // 16 bf16 = 32 bytes of every 128-byte row per step: start address + 2
                tc5::mma_bf16(d, a1 + 2 * k, bc + 2 * k, i192, (c | k) != 0);   // a1 x [c1 | c2 | c3]
                tc5::mma_bf16(d, a2 + 2 * k, bc + 2 * k, i128, 1u);             // a2 x [c1 | c2]
                tc5::mma_bf16(d, a3 + 2 * k, bc + 2 * k, i64, 1u);              // a3 x [c1]
            }
            tc5::commit(&bars.empty[st]);
            TC5_STAMP(6, g);                          // issuer: MMAs + commit issued
            if (c == it.nchunks - 1) tc5::commit(&bars.accum[acc]);
        }
    } else if (tid < 128) {
        mbar_wait(&bars.accum[acc], (uint32_t)((idx >> 1) & 1));
        tc5::fence_after_sync();
    }
}

// accumulator row of this lane, bodies c8*8 .. c8*8+7: the three 64-column blocks added up (smallest terms first)
__device__ __forceinline__ void t3_load8(uint32_t tmem_acc, int w, int c8, float (&v)[8]) {
    float u[8], t[8];
    const uint32_t base = tmem_acc + ((uint32_t)(w * 32) << 16) + (uint32_t)(c8 * 8);
    tc5::ld_32x32b_x8(base + 2 * kBG, t);
    tc5::ld_32x32b_x8(base + kBG, u);
    tc5::ld_32x32b_x8(base, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += u[e] + t[e];
}

struct BlendFwdBf3Params {
    const unsigned short *basis3, *coef3;
    const float *v_template;
    float *vp_out;
    int V, Kpad, B, ntiles, nbg;
};

// v_posed[b][n] = v_template[n] + sum_k coef[b][k] * basis[k][n]; work item = (128 coordinates, body group)
__global__ void __launch_bounds__(kT3Threads, 1) lbs_blend_fwd_bf3_kernel(const BlendFwdBf3Params p) {
    extern __shared__ __align__(1024) unsigned char smem_t3[];
    unsigned char *smem_raw = smem_t3 + ((1024u - (smem_u32(smem_t3) & 1023u)) & 1023u);
    __shared__ __align__(8) T3Bars bars;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t tmem_d = t3_setup(bars);
    pdl_wait();
    const int nchunks = p.Kpad / kKC3, N = 3 * p.V;
    int g0 = 0, idx = 0;
    for (int item = blockIdx.x; item < p.ntiles * p.nbg; item += gridDim.x, g0 += nchunks, ++idx) {
        // body groups of one basis tile are neighbours in the item order: they run at the same time on neighbouring CTAs,
        // so a batch of more than 64 bodies streams the basis from HBM once (the second group's reads hit L2)
        const int bg = item % p.nbg, tile = item / p.nbg;
        T3Item it;
        it.basis = p.basis3 + (size_t)tile * nchunks * (3 * kTM * kKC3);
        it.basis_stride = (size_t)3 * kTM * kKC3;
        it.act = p.coef3 + (size_t)bg * nchunks * (3 * kBG * kKC3);
        it.nchunks = nchunks;
        t3_mainloop(smem_raw, bars, tmem_d, it, g0, idx);
        if (w < 4) {
            const int n = tile * kTM + w * 32 + lane;          // TMEM lane = accumulator row = coordinate
            const float vt = n < N ? p.v_template[n] : 0.f;
            const uint32_t acc = tmem_d + (uint32_t)((idx & 1) * 256);
#pragma unroll 2
            for (int c8 = 0; c8 < kBG / 8; ++c8) {
                float v[8];
                t3_load8(acc, w, c8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int b = bg * kBG + c8 * 8 + e;
                    if (n < N && b < p.B) p.vp_out[(size_t)b * N + n] = v[e] + vt;
                }
            }
            tc5::fence_before_sync();
            mbar_arrive(&bars.tmem_free[idx & 1]);             // this accumulator may be overwritten
        }
    }
    pdl_launch_dependents();
    t3_teardown(tmem_d);
}

// part[ns][b][k] = sum_{n in split ns} gvp[b][n] * basis[k][n]; work item = (128 coefficients, split, body group);
// NC3 = coordinate chunks of 64
__global__ void __launch_bounds__(kT3Threads, 1)
lbs_dcoef_bf3_kernel(int Kpad, int NC3, int Bpad, const unsigned short *__restrict__ basis3, const unsigned short *__restrict__ gvp3,
                     float *__restrict__ part, int nsplit) {
    extern __shared__ __align__(1024) unsigned char smem_t3[];
    unsigned char *smem_raw = smem_t3 + ((1024u - (smem_u32(smem_t3) & 1023u)) & 1023u);
    __shared__ __align__(8) T3Bars bars;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t tmem_d = t3_setup(bars);
    pdl_wait();
    const int nkt = Kpad / kTM, nbg = Bpad / kBG;
    int g0 = 0, idx = 0;
    for (int item = blockIdx.x; item < nkt * nsplit * nbg; item += gridDim.x) {
        const int bg = item % nbg, kt = (item / nbg) % nkt, ns = item / (nbg * nkt);     // body group fastest (see the forward GEMM)
        const int c_begin = (int)((long)ns * NC3 / nsplit), c_end = (int)((long)(ns + 1) * NC3 / nsplit);
        float *o = part + ((size_t)ns * Bpad + (size_t)bg * kBG) * Kpad + kt * kTM + (w & 3) * 32 + lane;
        if (c_end == c_begin) {                    // more splits than chunks (small models): an empty split
            if (w < 4)
                for (int b = 0; b < kBG; ++b) o[(size_t)b * Kpad] = 0.f;
            continue;
        }
        T3Item it;
        it.basis = basis3 + ((size_t)c_begin * nkt + kt) * (3 * kTM * kKC3);
        it.basis_stride = (size_t)nkt * (3 * kTM * kKC3);
        it.act = gvp3 + ((size_t)bg * NC3 + c_begin) * (3 * kBG * kKC3);
        it.nchunks = c_end - c_begin;
        t3_mainloop(smem_raw, bars, tmem_d, it, g0, idx);
        if (w < 4) {
            const uint32_t acc = tmem_d + (uint32_t)((idx & 1) * 256);
#pragma unroll 2
            for (int c8 = 0; c8 < kBG / 8; ++c8) {
                float v[8];
                t3_load8(acc, w, c8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[(size_t)(c8 * 8 + e) * Kpad] = v[e];     // lane = coefficient: coalesced per body
            }
            tc5::fence_before_sync();
            mbar_arrive(&bars.tmem_free[idx & 1]);
        }
        g0 += it.nchunks;
        ++idx;
    }
    pdl_launch_dependents();
    t3_teardown(tmem_d);
}

// out[i] = sum over the leading dimension of in[n][count], fixed order (4 interleaved partial sums);
// blockIdx.y selects the array: 0 = the dcoef GEMM's N-splits, 1 = the dA chunk partials
__global__ void __launch_bounds__(256)
lbs_reduce2_kernel(const float *__restrict__ in0, int n0, long count0, float *__restrict__ out0,
                   const float *__restrict__ in1, int n1, long count1, float *__restrict__ out1) {
    pdl_wait();
    const float *__restrict__ in = blockIdx.y ? in1 : in0;
    float *__restrict__ out = blockIdx.y ? out1 : out0;
    const int n = blockIdx.y ? n1 : n0;
    const long count = blockIdx.y ? count1 : count0;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        s0 += in[(size_t)k * count + i];
        s1 += in[(size_t)(k + 1) * count + i];
        s2 += in[(size_t)(k + 2) * count + i];
        s3 += in[(size_t)(k + 3) * count + i];
    }
    for (; k < n; ++k) s0 += in[(size_t)k * count + i];
    out[i] = (s0 + s1) + (s2 + s3);
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
}

// per body: run the chain and Rodrigues backward on the reduced d coef
__global__ void __launch_bounds__(128)
lbs_pose_bwd_kernel(int J, int NB, int P, int Kpad, int Bpad, int nsplit,
                    const float *__restrict__ Jdirs, const int *__restrict__ parents,
                    const float *__restrict__ pose, const float *__restrict__ saved, SavedLayout L,
                    const float *__restrict__ dAsum, const float *__restrict__ part,
                    const float *__restrict__ gjoints,
                    float *__restrict__ gbetas, float *__restrict__ gpose,
                    float *__restrict__ gtransl, float *__restrict__ grot, int num_rot,
                    const float *__restrict__ rot6d, float *__restrict__ g6_root, float *__restrict__ g6A,
                    int g6_kpad, const TreeParam tree, const float *__restrict__ Jdirs_p) {
    __shared__ float sR[kMaxJ * 9], sJ[kMaxJ * 3], sGr[kMaxJ * 9], drel_s[kMaxJ * 3];
    __shared__ float dGr[kMaxJ * 9], dGt[kMaxJ * 3], dR[kMaxJ * 9], dJ[kMaxJ * 3];
    __shared__ float sdA[(kMaxJ + 1) * 12], dGtF[kMaxJ * 3];   // dGtF: dGt after the pull (what a parent reads)
    __shared__ float dbeta_direct[64], dbeta_part[6][32];
    __shared__ __align__(16) float sJd[kMaxJ * 3 * (kJdMaxNB + 1)];
    __shared__ __align__(8) uint64_t jbar;
    __shared__ TreeSmem st;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int S = NB | 1;
    if (Jdirs_p && tid == 0) {                 // constants: requested before the dependency wait
        const uint32_t bytes = (uint32_t)(((size_t)J * 3 * S + 3) / 4 * 16);
        mbar_init(&jbar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&jbar, bytes);
        tma_load_1d(sJd, Jdirs_p, bytes, &jbar);
    }
    stage_tree(st, tree, J);
    pdl_wait();
    const float *dA = dAsum + (size_t)b * (J + 1) * 12;      // chunk partials already summed; row J = d translation
    const float *iR = saved + L.R + (size_t)b * J * 9, *iJ = saved + L.Jr + (size_t)b * J * 3;
    const float *iGr = saved + L.Gr + (size_t)b * J * 9;
    // the joint whose rotation gradient this thread finishes at the end (axis-angle joints on the low threads, 6D /
    // matrix joints from thread 64 on: different warps, no divergence), and that joint's input
    const int jr = tid < J - num_rot ? num_rot + tid : (tid >= 64 && tid - 64 < num_rot ? tid - 64 : -1);
    float pin[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // everything this body reads from global memory in ONE round trip: all loads are issued before the first store
    // (loops with run-time bounds are not unrolled across their loads: 16 dependent round trips before)
    {
        constexpr int n9 = (kMaxJ * 9 + 127) / 128, n3 = (kMaxJ * 3 + 127) / 128, nA = ((kMaxJ + 1) * 12 + 127) / 128,
                      nP = ((kMaxJ - 1) * 9 + 64 + 127) / 128;
        float rR[n9], rG[n9], rJ[n3], rA[nA], rP[nP];
        const float *pb = part + (size_t)b * Kpad;           // already summed over the splits
        (void)nsplit; (void)Bpad;
        if (jr >= 0 && jr < num_rot && rot6d) {
#pragma unroll
            for (int e = 0; e < 6; ++e) pin[e] = rot6d[((size_t)b * num_rot + jr) * 6 + e];
        } else if (jr >= num_rot) {
#pragma unroll
            for (int e = 0; e < 3; ++e) pin[e] = pose[((size_t)b * J + jr) * 3 + e];
        }
#pragma unroll
        for (int i = 0; i < n9; ++i) {
            const int e = tid + i * 128;
            rR[i] = e < J * 9 ? iR[e] : 0.f;
            rG[i] = e < J * 9 ? iGr[e] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < n3; ++i) { const int e = tid + i * 128; rJ[i] = e < J * 3 ? iJ[e] : 0.f; }
#pragma unroll
        for (int i = 0; i < nA; ++i) { const int e = tid + i * 128; rA[i] = e < (J + 1) * 12 ? dA[e] : 0.f; }
#pragma unroll
        for (int i = 0; i < nP; ++i) { const int k = tid + i * 128; rP[i] = k < P + NB ? pb[k] : 0.f; }
#pragma unroll
        for (int i = 0; i < n9; ++i) {
            const int e = tid + i * 128;
            if (e < J * 9) { sR[e] = rR[i]; sGr[e] = rG[i]; }
        }
#pragma unroll
        for (int i = 0; i < n3; ++i) { const int e = tid + i * 128; if (e < J * 3) sJ[e] = rJ[i]; }
#pragma unroll
        for (int i = 0; i < nA; ++i) { const int e = tid + i * 128; if (e < (J + 1) * 12) sdA[e] = rA[i]; }
        // d pose-feature (added to dR below) and the direct d beta
#pragma unroll
        for (int i = 0; i < nP; ++i) {
            const int k = tid + i * 128;
            if (k < P) dR[9 + k] = rP[i];           // joint j = k/9+1, entry k%9
            else if (k < P + NB) dbeta_direct[k - P] = rP[i];
        }
    }
    if (tid < 9) dR[tid] = 0.f;
    __syncthreads();
    const float *dtr = sdA + J * 12;
    // dGr = dAr - dAt J^T ; dGt = dAt (+ d posed joints) ; dJ = -Gr^T dAt
    for (int j = tid; j < J; j += blockDim.x) {
        const float *a = sdA + j * 12;
        const float at[3] = {a[3], a[7], a[11]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) dGr[j * 9 + r * 3 + c] = a[r * 4 + c] - at[r] * sJ[j * 3 + c];
            dGt[j * 3 + r] = at[r] + (gjoints ? gjoints[((size_t)b * J + j) * 3 + r] : 0.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            dJ[j * 3 + c] = -(sGr[j * 9 + c] * at[0] + sGr[j * 9 + 3 + c] * at[1] + sGr[j * 9 + 6 + c] * at[2]);
    }
    __syncthreads();
    // chain backward, deepest tree level first.  A joint first PULLS from its children (fixed child order:
    // deterministic, race free), then finishes its own dR / d rel.  Three threads per joint, one per COLUMN cc of
    // dGr[j]: column cc of the pulled sum is all that dR[j][.][cc] needs, so the three never exchange anything.
    for (int l = tree.nlev - 1; l >= 0; --l) {
        const int q0 = st.lvl_start[l], nj = st.lvl_start[l + 1] - q0;
        for (int u = tid; u < nj * 3; u += blockDim.x) {
            const int j = st.lvl_joint[q0 + u / 3], cc = u % 3;
            float g[3], gt[3], djc = dJ[j * 3 + cc];
#pragma unroll
            for (int r = 0; r < 3; ++r) { g[r] = dGr[j * 9 + r * 3 + cc]; gt[r] = dGt[j * 3 + r]; }
            for (int ci = st.child_start[j]; ci < st.child_start[j + 1]; ++ci) {
                const int c = st.child_list[ci];
                const float *Rc = sR + c * 9, *gc = dGr + c * 9, *gtc = dGtF + c * 3;
                const float relc = sJ[c * 3 + cc] - sJ[j * 3 + cc];
#pragma unroll
                for (int r = 0; r < 3; ++r) {      // dGr[j] += dGr[c] Rc^T + dGt[c] rel^T
                    g[r] += gc[r * 3] * Rc[cc * 3] + gc[r * 3 + 1] * Rc[cc * 3 + 1] + gc[r * 3 + 2] * Rc[cc * 3 + 2] +
                            gtc[r] * relc;
                    gt[r] += gtc[r];
                }
                djc -= drel_s[c * 3 + cc];
            }
            float drel = 0.f;
            if (j == 0) {
#pragma unroll
                for (int r = 0; r < 3; ++r) dR[r * 3 + cc] += g[r];
                djc += gt[cc];
            } else {
                const float *Gp = sGr + st.parents[j] * 9;
#pragma unroll
                for (int r = 0; r < 3; ++r)        // dR[j] += Gp^T dGr[j]
                    dR[j * 9 + r * 3 + cc] += Gp[r] * g[0] + Gp[3 + r] * g[1] + Gp[6 + r] * g[2];
                drel = Gp[cc] * gt[0] + Gp[3 + cc] * gt[1] + Gp[6 + cc] * gt[2];
                djc += drel;
            }
            // a thread overwrites only what no other thread of this level reads (its own column of dGr[j], its own
            // entries of the other arrays; dGt[j], read by all three, stays as it was)
#pragma unroll
            for (int r = 0; r < 3; ++r) dGr[j * 9 + r * 3 + cc] = g[r];
            drel_s[j * 3 + cc] = drel;
            dGtF[j * 3 + cc] = gt[cc];
            dJ[j * 3 + cc] = djc;
        }
        __syncthreads();
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    // Rodrigues backward (lbs.py:177-191); joints given as matrices export dR itself
    for (int j = jr < 0 ? J : jr; j < J; j = J) {
        float *o = gpose + ((size_t)b * J + j) * 3;
        if (j < num_rot && rot6d) {
            // Gram-Schmidt backward fused here: joint 0 -> g6_root [B,6]; joints 1.. -> g6A, the
            // A operand ([chunk][64 bodies][32 k, swizzled], k = (j-1)*6+e, zero up to g6_kpad) of the
            // decoder's backward GEMM (fit.cu)
            float dx6[6];
            gs_bwd(pin, dR + j * 9, dx6);
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                if (j == 0) g6_root[(size_t)b * 6 + e] = dx6[e];
                else g6A[a_index(b, (j - 1) * 6 + e, g6_kpad)] = dx6[e];
            }
            if (j == num_rot - 1)
                for (int k = (num_rot - 1) * 6; k < g6_kpad; ++k) g6A[a_index(b, k, g6_kpad)] = 0.f;
            o[0] = o[1] = o[2] = 0.f;
            continue;
        }
        if (j < num_rot) {
#pragma unroll
            for (int e = 0; e < 9; ++e) grot[((size_t)b * num_rot + j) * 9 + e] = dR[j * 9 + e];
            o[0] = o[1] = o[2] = 0.f;
            continue;
        }
        const float *r = pin;
        const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
        const float a = sqrtf(ex * ex + ey * ey + ez * ez);
        const float nx = r[0] / a, ny = r[1] / a, nz = r[2] / a;
        float s, c;
        sincosf(a, &s, &c);
        const float t = 1.0f - c;
        const float K[9] = {0.f, -nz, ny, nz, 0.f, -nx, -ny, nx, 0.f};
        float K2[9];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                K2[rr * 3 + cc] = K[rr * 3] * K[cc] + K[rr * 3 + 1] * K[3 + cc] + K[rr * 3 + 2] * K[6 + cc];
        const float *d = dR + j * 9;
        float sK = 0.f, sK2 = 0.f;
#pragma unroll
        for (int e = 0; e < 9; ++e) { sK += d[e] * K[e]; sK2 += d[e] * K2[e]; }
        float da = sK * c + sK2 * s;
        float dK[9];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                // (dR K^T + K^T dR)[rr][cc]
                const float m1 = d[rr * 3] * K[cc * 3] + d[rr * 3 + 1] * K[cc * 3 + 1] + d[rr * 3 + 2] * K[cc * 3 + 2];
                const float m2 = K[rr] * d[cc] + K[3 + rr] * d[3 + cc] + K[6 + rr] * d[6 + cc];
                dK[rr * 3 + cc] = s * d[rr * 3 + cc] + t * (m1 + m2);
            }
        const float dn0 = dK[7] - dK[5], dn1 = dK[2] - dK[6], dn2 = dK[3] - dK[1];
        da += -(dn0 * r[0] + dn1 * r[1] + dn2 * r[2]) / (a * a);
        o[0] = dn0 / a + da * ex / a;
        o[1] = dn1 / a + da * ey / a;
        o[2] = dn2 / a + da * ez / a;
    }
    // d beta = direct term + Jdirs^T dJ: (l, slice of the joint coordinates) per thread, slices summed in order
    if (Jdirs_p) {
        const int l = tid % 32, part = tid / 32;                 // one slice per warp, Jdirs from shared memory
        mbar_wait(&jbar, 0);
        float s0 = 0.f, s1 = 0.f;
        if (l < NB) {
            int e = part;
            for (; e + 4 < J * 3; e += 8) {
                s0 = fmaf(sJd[e * S + l], dJ[e], s0);
                s1 = fmaf(sJd[(e + 4) * S + l], dJ[e + 4], s1);
            }
            if (e < J * 3) s0 = fmaf(sJd[e * S + l], dJ[e], s0);
        }
        dbeta_part[part][l] = s0 + s1;
        __syncthreads();
        if (tid < NB)
            gbetas[(size_t)b * NB + tid] =
                dbeta_direct[tid] + ((dbeta_part[0][tid] + dbeta_part[1][tid]) + (dbeta_part[2][tid] + dbeta_part[3][tid]));
    } else if (NB <= 32) {
        const int l = tid % 32, part = tid / 32;                 // 128 threads: 4 slices active, parts 4, 5 done by a second pass
        for (int pp = part; pp < 6; pp += 4) {
            float s = 0.f;
            if (l < NB)
                for (int e = pp; e < J * 3; e += 6) s = fmaf(Jdirs[(size_t)e * NB + l], dJ[e], s);
            dbeta_part[pp][l] = s;
        }
        __syncthreads();
        if (tid < NB) {
            float s = dbeta_direct[tid];
            for (int pp = 0; pp < 6; ++pp) s += dbeta_part[pp][tid];
            gbetas[(size_t)b * NB + tid] = s;
        }
    } else {
        for (int l = tid; l < NB; l += blockDim.x) {
            float s = dbeta_direct[l];
            for (int e = 0; e < J * 3; ++e) s = fmaf(Jdirs[(size_t)e * NB + l], dJ[e], s);
            gbetas[(size_t)b * NB + l] = s;
        }
    }
    if (gtransl && tid < 3) {
        float s = dtr[tid];                      // row J of the chunk sums: sum over the vertices of gw
        if (gjoints)
            for (int j = 0; j < J; ++j) s += gjoints[((size_t)b * J + j) * 3 + tid];
        gtransl[(size_t)b * 3 + tid] = s;
    }
}

struct BwdLayout {
    size_t gvp, gvp_lo, dApart, part, dsum, dA, total;  // in floats
    int Bpad;
};
static BwdLayout bwd_layout(const psi_lbs_model *m, int B) {
    BwdLayout l;
    l.Bpad = ((B + kBG - 1) / kBG) * kBG;
    size_t o = 0;
    l.gvp = o;     o += (size_t)l.Bpad * m->Npad;
    l.gvp_lo = o;  o += (size_t)l.Bpad * m->Npad;
    l.dApart = o;  o += (size_t)m->NCH * B * (m->J + 1) * 12;
    o = (o + 3) & ~(size_t)3;
    l.part = o;    o += (size_t)kNSplit * l.Bpad * m->Kpad;
    l.dsum = o;    o += (size_t)l.Bpad * m->Kpad;
    l.dA = o;      o += (size_t)B * (m->J + 1) * 12;
    l.total = o;
    return l;
}

template <typename T>
static int upload(T **dst, const std::vector<T> &h, cudaStream_t st, size_t *bytes) {
    const size_t nb = h.size() * sizeof(T);
    cudaError_t e = cudaMalloc((void **)dst, nb ? nb : sizeof(T));
    if (e != cudaSuccess) return PSI_ERR_ALLOC;
    if (nb) {
        e = cudaMemcpyAsync(*dst, h.data(), nb, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    *bytes += nb;
    return PSI_OK;
}

}  // namespace psi

extern "C" {

void psi_lbs_model_destroy(psi_lbs_model *m) {
    if (!m) return;
    cudaFree(m->basis_fwd); cudaFree(m->basis_bwd); cudaFree(m->basis_fwd3); cudaFree(m->basis_bwd3); cudaFree(m->v_template); cudaFree(m->Jt); cudaFree(m->Jdirs); cudaFree(m->Jdirs_p);
    cudaFree(m->skin_w); cudaFree(m->ch_w); cudaFree(m->skin_j); cudaFree(m->parents);
    cudaFree(m->ch_seg); cudaFree(m->ch_lv); cudaFree(m->tree_buf); cudaFree(m->ch_ju); cudaFree(m->unit_desc);
    delete m;
}

int psi_lbs_model_create(psi_lbs_model **out, int V, int J, int NB, const float *h_v_template,
                         const float *h_shapedirs, const float *h_posedirs,
                         const float *h_J_regressor, const float *h_weights, const int *h_parents,
                         psi_stream_t stream) {
    psi::Range nvtx_range("psi_lbs_model_create");
    using namespace psi;
    if (!out || V < 1 || J < 1 || NB < 0 || !h_v_template || !h_shapedirs || !h_posedirs ||
        !h_J_regressor || !h_weights || !h_parents)
        return PSI_ERR_BAD_ARG;
    if (J > kMaxJ || NB > 64) return PSI_ERR_UNSUPPORTED;
    for (int j = 1; j < J; ++j)
        if (h_parents[j] < 0 || h_parents[j] >= j) return PSI_ERR_BAD_ARG;  // parent before child
    cudaStream_t st = (cudaStream_t)stream;
    psi_lbs_model *m = new (std::nothrow) psi_lbs_model();
    if (!m) return PSI_ERR_ALLOC;
    m->V = V; m->J = J; m->NB = NB; m->P = (J - 1) * 9; m->K = m->P + NB;
    m->Kpad = ((m->K + kDK - 1) / kDK) * kDK;   // multiple of the dcoef k tile (and of kKC, kKC3)
    m->Npad = ((3 * V + kKC3 - 1) / kKC3) * kKC3;   // multiple of 64: whole 128-byte bf16 rows (and of kKC)
    m->NC = m->Npad / kKC;
    const int gmode = lbs_gemm_mode();
    const int FT = gmode == kGemmMma ? kFT : kTM;   // rows of a forward basis tile
    m->FT = FT;
    m->NT = (3 * V + FT - 1) / FT;
    m->bytes = 0;
    const int P = m->P, Kpad = m->Kpad, Npad = m->Npad;
    const size_t N = (size_t)3 * V;

    // The blend basis (posedirs rows, then shapedirs columns) twice, each as the B operand of its
    // GEMM with the reduction index contiguous in swizzled 128-byte rows:
    //   forward  [tile n/72][chunk k/32][row n%72][32 k]     (reduction over k)
    //   backward [chunk n/32][row k][32 n]                    (reduction over n)
    // bf3 path: the same matrix as three bfloat16 terms per element (split once, here), 128-byte rows of 64:
    //   forward  [tile n/128][chunk k/64][term][row n%128][64 k]      one 48 kB TMA block per pipeline stage
    //   backward [chunk n/64][k tile k/128][term][row k%128][64 n]    likewise
    const bool bf3 = gmode == kGemmBf3;
    std::vector<float> bf(bf3 ? 0 : (size_t)m->NT * Kpad * FT, 0.f), bb(bf3 ? 0 : (size_t)m->NC * Kpad * kKC, 0.f);
    const int Kc3 = Kpad / kKC3, nkt = Kpad / kTM;
    std::vector<unsigned short> bf3f(bf3 ? (size_t)m->NT * Kc3 * 3 * kTM * kKC3 : 0, 0),
        bf3b(bf3 ? (size_t)(Npad / kKC3) * nkt * 3 * kTM * kKC3 : 0, 0);
    auto put = [&](int k, size_t n, float x) {
        if (bf3) {
            unsigned short t[3];
            split_bf16x3(x, t[0], t[1], t[2]);
            const int rf = (int)(n % kTM), rb = k % kTM;
            for (int e = 0; e < 3; ++e) {
                bf3f[(((n / kTM) * Kc3 + (size_t)(k / kKC3)) * 3 + e) * (kTM * kKC3) + (size_t)rf * kKC3 + swz16(rf, k % kKC3)] = t[e];
                bf3b[(((n / kKC3) * nkt + (size_t)(k / kTM)) * 3 + e) * (kTM * kKC3) + (size_t)rb * kKC3 + swz16(rb, (int)(n % kKC3))] = t[e];
            }
            return;
        }
        const int r = (int)(n % FT);
        bf[(n / FT) * (size_t)Kpad * FT + (size_t)(k / kKC) * (FT * kKC) + (size_t)r * kKC + swz(r, k % kKC)] = x;
        bb[(n / kKC) * (size_t)Kpad * kKC + (size_t)k * kKC + swz(k, (int)(n % kKC))] = x;
    };
    for (int k = 0; k < P; ++k)
        for (size_t n = 0; n < N; ++n) put(k, n, h_posedirs[(size_t)k * N + n]);
    for (int l = 0; l < NB; ++l)
        for (size_t n = 0; n < N; ++n) put(P + l, n, h_shapedirs[n * NB + l]);
    std::vector<float> vt((size_t)m->NT * FT, 0.f);
    for (size_t n = 0; n < N; ++n) vt[n] = h_v_template[n];
    (void)Npad;

    // fold the joint regressor: Jt = Jreg * v_template, Jdirs = Jreg * shapedirs (double accum)
    std::vector<float> Jt((size_t)J * 3), Jdirs((size_t)J * 3 * NB);
    {
        std::vector<double> acc((size_t)3 * (NB + 1));
        for (int j = 0; j < J; ++j) {
            std::fill(acc.begin(), acc.end(), 0.0);
            for (int v = 0; v < V; ++v) {
                const double w = h_J_regressor[(size_t)j * V + v];
                if (w == 0.0) continue;
                for (int c = 0; c < 3; ++c) {
                    acc[(size_t)c * (NB + 1)] += w * h_v_template[(size_t)v * 3 + c];
                    for (int l = 0; l < NB; ++l)
                        acc[(size_t)c * (NB + 1) + 1 + l] += w * h_shapedirs[((size_t)v * 3 + c) * NB + l];
                }
            }
            for (int c = 0; c < 3; ++c) {
                Jt[(size_t)j * 3 + c] = (float)acc[(size_t)c * (NB + 1)];
                for (int l = 0; l < NB; ++l)
                    Jdirs[((size_t)j * 3 + c) * NB + l] = (float)acc[(size_t)c * (NB + 1) + 1 + l];
            }
        }
    }
    // sparse skinning weights
    int KW = 1;
    for (int v = 0; v < V; ++v) {
        int c = 0;
        for (int j = 0; j < J; ++j) c += (h_weights[(size_t)v * J + j] != 0.f);
        KW = c > KW ? c : KW;
    }
    m->KW = KW;
    std::vector<int> skin_j((size_t)V * KW, 0);
    std::vector<float> skin_w((size_t)V * KW, 0.f);
    for (int v = 0; v < V; ++v) {
        int c = 0;
        for (int j = 0; j < J; ++j) {
            const float w = h_weights[(size_t)v * J + j];
            if (w != 0.f) {
                skin_j[(size_t)v * KW + c] = j;
                skin_w[(size_t)v * KW + c] = w;
                ++c;
            }
        }
    }
    // per 256-vertex chunk: the skinning entries bucketed by joint (local vertex, weight), ascending
    // vertex order inside a bucket -- the fixed summation order of the dA partials
    m->NCH = ((m->Npad + 2) / 3 + 255) / 256;
    std::vector<int> ch_seg((size_t)m->NCH * (J + 1), 0);
    std::vector<unsigned char> ch_lv;
    std::vector<float> ch_w;
    m->nnz = 0;
    for (int c = 0; c < m->NCH; ++c)
        for (int j = 0; j <= J; ++j) {
            ch_seg[(size_t)c * (J + 1) + j] = (int)ch_lv.size();
            if (j == J) {
                // every chunk's list starts at a multiple of 16 entries: the vertex kernel stages it with two bulk
                // copies (16-byte aligned source, size a multiple of 16 bytes); the pad entries carry weight 0
                while (ch_lv.size() % 16) { ch_lv.push_back(0); ch_w.push_back(0.f); }
                break;
            }
            for (int lv = 0; lv < 256; ++lv) {
                const int v = c * 256 + lv;
                if (v >= V) break;
                const float w = h_weights[(size_t)v * J + j];
                if (w != 0.f) { ch_lv.push_back((unsigned char)lv); ch_w.push_back(w); }
            }
        }
    m->nnz = (long)ch_lv.size();
    m->max_ent = 0;
    for (int c = 0; c < m->NCH; ++c) {
        const int ne = ch_seg[(size_t)c * (J + 1) + J] - ch_seg[(size_t)c * (J + 1)];
        m->max_ent = ne > m->max_ent ? ne : m->max_ent;
    }
    // work units: every joint's entry list of a chunk cut into pieces of <= kUnitLen entries;
    // unit_desc = (first entry relative to the chunk) << 8 | length, ch_ju = per (chunk, joint) first unit
    std::vector<int> ch_ju((size_t)m->NCH * (J + 1), 0), unit_desc;
    m->max_units = 0;
    for (int c = 0; c < m->NCH; ++c) {
        const int e0 = ch_seg[(size_t)c * (J + 1)];
        const int ubegin = (int)unit_desc.size();
        for (int j = 0; j <= J; ++j) {
            ch_ju[(size_t)c * (J + 1) + j] = (int)unit_desc.size();
            if (j == J) break;
            for (int k = ch_seg[(size_t)c * (J + 1) + j]; k < ch_seg[(size_t)c * (J + 1) + j + 1]; k += kUnitLen) {
                const int len = std::min(kUnitLen, ch_seg[(size_t)c * (J + 1) + j + 1] - k);
                unit_desc.push_back(((k - e0) << 8) | len);
            }
        }
        m->max_units = std::max(m->max_units, (int)unit_desc.size() - ubegin);
    }
    if (unit_desc.empty()) unit_desc.push_back(0);
    std::vector<int> parents(h_parents, h_parents + J);
    parents[0] = -1;
    // tree levels and children lists, packed: [lvl_start (J+1) | lvl_joint (J) | child_start (J+1) | child_list (J)]
    std::vector<int> level(J, 0), treebuf((size_t)4 * J + 2, 0);
    int nlev = 1;
    for (int j = 1; j < J; ++j) { level[j] = level[parents[j]] + 1; nlev = level[j] + 1 > nlev ? level[j] + 1 : nlev; }
    {
        int *lvl_start = treebuf.data(), *lvl_joint = lvl_start + J + 1, *child_start = lvl_joint + J, *child_list = child_start + J + 1;
        for (int j = 0; j < J; ++j) ++lvl_start[level[j] + 1];
        for (int l = 0; l < J; ++l) lvl_start[l + 1] += lvl_start[l];
        std::vector<int> fillp(lvl_start, lvl_start + J);
        for (int j = 0; j < J; ++j) lvl_joint[fillp[level[j]]++] = j;
        for (int j = 1; j < J; ++j) ++child_start[parents[j] + 1];
        for (int j = 0; j < J; ++j) child_start[j + 1] += child_start[j];
        std::vector<int> fillc(child_start, child_start + J);
        for (int j = 1; j < J; ++j) child_list[fillc[parents[j]]++] = j;
    }

    int rc = PSI_OK;
    m->basis_fwd = m->basis_bwd = nullptr;
    m->basis_fwd3 = m->basis_bwd3 = nullptr;
    if (bf3) {
        if (rc == PSI_OK) rc = upload(&m->basis_fwd3, bf3f, st, &m->bytes);
        if (rc == PSI_OK) rc = upload(&m->basis_bwd3, bf3b, st, &m->bytes);
    } else {
        if (rc == PSI_OK) rc = upload(&m->basis_fwd, bf, st, &m->bytes);
        if (rc == PSI_OK) rc = upload(&m->basis_bwd, bb, st, &m->bytes);
    }
    if (rc == PSI_OK) rc = upload(&m->v_template, vt, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->Jt, Jt, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->Jdirs, Jdirs, st, &m->bytes);
    if (NB <= kJdMaxNB) {       // rows of an odd length: the copy the pose kernels stage in shared memory (one TMA bulk copy)
        const int S = NB | 1;
        std::vector<float> jp(((size_t)J * 3 * S + 3) / 4 * 4, 0.f);
        for (int e = 0; e < J * 3; ++e)
            for (int l = 0; l < NB; ++l) jp[(size_t)e * S + l] = Jdirs[(size_t)e * NB + l];
        if (rc == PSI_OK) rc = upload(&m->Jdirs_p, jp, st, &m->bytes);
    }
    if (rc == PSI_OK) rc = upload(&m->skin_j, skin_j, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->skin_w, skin_w, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->parents, parents, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->ch_seg, ch_seg, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->ch_lv, ch_lv, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->ch_w, ch_w, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->ch_ju, ch_ju, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->unit_desc, unit_desc, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->tree_buf, treebuf, st, &m->bytes);
    {
        m->tp.nlev = nlev;
        const int *tb = treebuf.data();
        for (int i = 0; i <= J; ++i) {
            m->tp.lvl_start[i] = (unsigned char)tb[i];
            m->tp.child_start[i] = (unsigned char)tb[2 * J + 1 + i];
        }
        for (int i = 0; i < J; ++i) {
            m->tp.lvl_joint[i] = (unsigned char)tb[J + 1 + i];
            m->tp.child_list[i] = (unsigned char)tb[3 * J + 2 + i];
            m->tp.parents[i] = (unsigned char)(parents[i] < 0 ? 0 : parents[i]);      // (the root's is never read)
        }
    }
    if (rc == PSI_OK) {
        m->tree.nlev = nlev;
        m->tree.lvl_start = m->tree_buf;
        m->tree.lvl_joint = m->tree_buf + J + 1;
        m->tree.child_start = m->tree_buf + 2 * J + 1;
        m->tree.child_list = m->tree_buf + 3 * J + 2;
    }
    if (rc == PSI_OK) {
        cudaError_t e = cudaStreamSynchronize(st);   // host vectors die with this scope
        if (e != cudaSuccess) rc = (int)e;
    }
    if (rc != PSI_OK) {
        psi_lbs_model_destroy(m);
        return rc;
    }
    *out = m;
    return PSI_OK;
}

size_t psi_lbs_model_bytes(const psi_lbs_model *m) { return m ? m->bytes : 0; }
}  // extern "C"
namespace psi {
int lbs_vertex_chunks(const psi_lbs_model *m) { return m ? m->NCH : 0; }
}
extern "C" {

size_t psi_lbs_saved_floats(const psi_lbs_model *m, int B) {
    if (!m || B <= 0) return 0;
    return psi::saved_layout(B, m->J, m->V, m->Kpad).total;
}

}  // extern "C"

namespace psi {
// psi_lbs_fwd with the leading num_rot joints given either as matrices (rot_in) or in the 6D
// representation (rot6d [B,num_rot,6], Gram-Schmidt fused into the pose kernel: the fitting loop)
int lbs_fwd_impl(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                 const float *transl, const float *cam, long cam_bstride, const float *rot_in,
                 const float *rot6d, int num_rot, float *verts, float *joints, float *saved,
                 const SdfFuse *sdf, cudaStream_t st) {
    psi::Range nvtx_range("psi_lbs_fwd");
    if (!m || B < 0) return PSI_ERR_BAD_ARG;
    if (B == 0) return PSI_OK;
    if (!betas || !pose || !verts || !saved) return PSI_ERR_BAD_ARG;
    if (num_rot < 0 || num_rot > m->J || (num_rot > 0 && !rot_in && !rot6d)) return PSI_ERR_BAD_ARG;
    if (B > 65535) return PSI_ERR_UNSUPPORTED;
    if (((uintptr_t)saved & 127u) != 0) return PSI_ERR_BAD_ARG;
    const SavedLayout L = saved_layout(B, m->J, m->V, m->Kpad);
    const bool tc5 = lbs_gemm_tc5();
    const int gmode = lbs_gemm_mode();
    (psi::skip_kernel("lbs_pose_fwd") ? cudaSuccess : launch_pdl(lbs_pose_fwd_kernel, dim3(B), dim3(128), 0, st, m->J, m->NB, m->P, m->Kpad, m->Jt, m->Jdirs, m->parents,
                                          B, betas, pose, transl, saved, L, joints, rot_in, num_rot, rot6d, gmode, m->tp, m->Jdirs_p));
    PSI_LAUNCHED_K("lbs_pose_fwd");
    if (B % kBG && gmode == kGemmBf3) {
        (psi::skip_kernel("lbs_zero_coef_pad") ? cudaSuccess : launch_pdl(lbs_zero_coef_pad3_kernel, dim3(8), dim3(256), 0, st,
                   reinterpret_cast<unsigned short *>(saved + L.coef), B, m->Kpad));
        PSI_LAUNCHED_K("lbs_zero_coef_pad");
    } else if (B % kBG) {
        (psi::skip_kernel("lbs_zero_coef_pad") ? cudaSuccess : launch_pdl(lbs_zero_coef_pad_kernel, dim3(8), dim3(256), 0, st, saved + L.coef, B, m->Kpad));
        PSI_LAUNCHED_K("lbs_zero_coef_pad");
        if (tc5) {
            (psi::skip_kernel("lbs_zero_coef_pad") ? cudaSuccess : launch_pdl(lbs_zero_coef_pad_kernel, dim3(8), dim3(256), 0, st, saved + L.coef_lo, B, m->Kpad));
            PSI_LAUNCHED_K("lbs_zero_coef_pad");
        }
    }
    dim3 grid((unsigned)m->NT, (unsigned)((B + kBG - 1) / kBG));
    if (gmode == kGemmBf3) {
        BlendFwdBf3Params p;
        p.basis3 = m->basis_fwd3; p.coef3 = reinterpret_cast<const unsigned short *>(saved + L.coef); p.v_template = m->v_template;
        p.vp_out = saved + L.vp; p.V = m->V; p.Kpad = m->Kpad; p.B = B; p.ntiles = m->NT; p.nbg = (B + kBG - 1) / kBG;
        const size_t smem = (size_t)kT3Smem + 1024;
        if (const int arc = ensure_max_dyn_smem(lbs_blend_fwd_bf3_kernel, smem)) return arc;
        const int items = p.ntiles * p.nbg;
        (psi::skip_kernel("lbs_blend_fwd_bf3") ? cudaSuccess : launch_pdl(lbs_blend_fwd_bf3_kernel, dim3((unsigned)(items < PSI_NUM_SMS ? items : PSI_NUM_SMS)), dim3(kT3Threads),
                   smem, st, p));
        PSI_LAUNCHED_K("lbs_blend_fwd_bf3");
    } else if (tc5) {
        BlendFwdTc5Params p;
        p.basis_fwd = m->basis_fwd; p.v_template = m->v_template; p.coef_hi = saved + L.coef; p.coef_lo = saved + L.coef_lo;
        p.vp_out = saved + L.vp; p.V = m->V; p.Kpad = m->Kpad; p.B = B; p.ntiles = m->NT; p.nbg = (B + kBG - 1) / kBG;
        const size_t smem = (size_t)kTSmem + 1024;
        if (const int arc = ensure_max_dyn_smem(lbs_blend_fwd_tc5_kernel, smem)) return arc;
        const int items = p.ntiles * p.nbg;
        (psi::skip_kernel("lbs_blend_fwd_tc5") ? cudaSuccess : launch_pdl(lbs_blend_fwd_tc5_kernel, dim3((unsigned)(items < PSI_NUM_SMS ? items : PSI_NUM_SMS)), dim3(kTThreads),
                   smem, st, p));
        PSI_LAUNCHED_K("lbs_blend_fwd_tc5");
    } else {
        BlendFwdParams p;
        p.basis_fwd = m->basis_fwd; p.v_template = m->v_template; p.coef = saved + L.coef;
        p.vp_out = saved + L.vp; p.V = m->V; p.Kpad = m->Kpad; p.B = B;
        const size_t smem = (size_t)kFStages * kFStageBytes;
        if (const int arc = ensure_max_dyn_smem(lbs_blend_fwd_kernel, smem)) return arc;
        (psi::skip_kernel("lbs_blend_fwd") ? cudaSuccess : launch_pdl(lbs_blend_fwd_kernel, grid, dim3(128), smem, st, p));
        PSI_LAUNCHED_K("lbs_blend_fwd");
    }
    {
        dim3 sgrid((unsigned)((m->V + 255) / 256), (unsigned)B);
        static const bool two = [] { const char *e = getenv("PSI_SKIN_SDF"); return !(e && e[0] == '1'); }();   // PSI_SKIN_SDF=1: one vertex per thread
        if (sdf && two && m->KW == 4)
            (psi::skip_kernel("lbs_skin_sdf_fwd") ? cudaSuccess : launch_pdl(lbs_skin_sdf2_kernel, dim3((unsigned)((m->V + 511) / 512), (unsigned)B), dim3(256), 0, st,
                       m->V, m->J, m->skin_j, m->skin_w, saved + L.A, saved + L.vp, transl, cam, cam_bstride, verts, *sdf));
        else if (sdf)
            (psi::skip_kernel(sdf ? "lbs_skin_sdf_fwd" : "lbs_skin_fwd") ? cudaSuccess : launch_pdl(lbs_skin_fwd_kernel<true>, sgrid, dim3(256), 0, st, m->V, m->J, m->KW, m->skin_j, m->skin_w,
                       saved + L.A, saved + L.vp, transl, cam, cam_bstride, verts, *sdf));
        else
            (psi::skip_kernel(sdf ? "lbs_skin_sdf_fwd" : "lbs_skin_fwd") ? cudaSuccess : launch_pdl(lbs_skin_fwd_kernel<false>, sgrid, dim3(256), 0, st, m->V, m->J, m->KW, m->skin_j, m->skin_w,
                       saved + L.A, saved + L.vp, transl, cam, cam_bstride, verts, SdfFuse()));
        PSI_LAUNCHED_K(sdf ? "lbs_skin_sdf_fwd" : "lbs_skin_fwd");
    }
    return PSI_OK;
}
}  // namespace psi

extern "C" {

int psi_lbs_fwd(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                const float *transl, const float *cam, long cam_bstride, const float *rot_in,
                int num_rot, float *verts, float *joints, float *saved, psi_stream_t stream) {
    if (num_rot > 0 && !rot_in) return PSI_ERR_BAD_ARG;
    return psi::lbs_fwd_impl(m, B, betas, pose, transl, cam, cam_bstride, rot_in, nullptr, num_rot, verts,
                             joints, saved, nullptr, (cudaStream_t)stream);
}

size_t psi_lbs_bwd_workspace_bytes(const psi_lbs_model *m, int B) {
    if (!m || B <= 0) return 0;
    return psi::bwd_layout(m, B).total * sizeof(float);
}

}  // extern "C"

namespace psi {
// psi_lbs_bwd2 + the 6D variant: with rot6d the gradient of the leading joints is returned through
// the Gram-Schmidt backward (g6_root [B,6], g6A = GEMM operand layout, see lbs_pose_bwd_kernel)
int lbs_bwd_impl(const psi_lbs_model *m, int B, const float *pose, const float *cam, long cam_bstride,
                 const float *saved, const float *grad_verts, const VGradFuse *vg, const float *grad_joints,
                 float *grad_betas, float *grad_pose, float *grad_transl, float *grad_rot, int num_rot,
                 const float *rot6d, float *g6_root, float *g6A, int g6_kpad, void *workspace,
                 size_t workspace_bytes, cudaStream_t st) {
    psi::Range nvtx_range("psi_lbs_bwd");
    if (!m || B < 0) return PSI_ERR_BAD_ARG;
    if (B == 0) return PSI_OK;
    if (!pose || !saved || (!grad_verts && !vg) || !grad_betas || !grad_pose || !workspace) return PSI_ERR_BAD_ARG;
    if (num_rot < 0 || num_rot > m->J || (num_rot > 0 && !grad_rot && !rot6d)) return PSI_ERR_BAD_ARG;
    if (rot6d && (!g6_root || !g6A || g6_kpad < (num_rot - 1) * 6)) return PSI_ERR_BAD_ARG;
    if (B > 65535) return PSI_ERR_UNSUPPORTED;
    const BwdLayout W = bwd_layout(m, B);
    if (workspace_bytes < W.total * sizeof(float)) return PSI_ERR_WORKSPACE;
    if (((uintptr_t)workspace & 15u) != 0) return PSI_ERR_BAD_ARG;
    float *ws = reinterpret_cast<float *>(workspace);
    const SavedLayout L = saved_layout(B, m->J, m->V, m->Kpad);
    const bool tc5 = lbs_gemm_tc5();
    const int gmode = lbs_gemm_mode();
    const int nsplit = gmode == kGemmMma ? kNSplit : kNSplitTc5;
    {
        dim3 grid((unsigned)m->NCH, (unsigned)W.Bpad);
        // the chunk's skinning entries are staged in shared memory when they fit (5 bytes each)
        const int cap = m->max_ent <= 8192 ? ((m->max_ent + 15) & ~15) : 0;
        const size_t smem = (size_t)cap * 5;
        const int units = (cap > 0 && m->max_units <= kMaxUnits) ? 1 : 0;
        if (vg)
            (psi::skip_kernel(vg ? "lbs_vertex_bwd_fit" : "lbs_vertex_bwd") ? cudaSuccess : launch_pdl_sel(lbs_vertex_bwd_kernel<true>, grid, dim3(256), smem, st, m->V, m->J, m->KW, m->Npad, B, m->skin_j,
                       m->skin_w, saved + L.A, saved + L.vp, cam, cam_bstride, grad_verts, m->ch_seg, m->ch_lv,
                       m->ch_w, cap, m->ch_ju, m->unit_desc, units, ws + W.gvp, tc5 ? ws + W.gvp_lo : nullptr, gmode, ws + W.dApart, *vg));
        else
            (psi::skip_kernel(vg ? "lbs_vertex_bwd_fit" : "lbs_vertex_bwd") ? cudaSuccess : launch_pdl(lbs_vertex_bwd_kernel<false>, grid, dim3(256), smem, st, m->V, m->J, m->KW, m->Npad, B, m->skin_j,
                       m->skin_w, saved + L.A, saved + L.vp, cam, cam_bstride, grad_verts, m->ch_seg, m->ch_lv,
                       m->ch_w, cap, m->ch_ju, m->unit_desc, units, ws + W.gvp, tc5 ? ws + W.gvp_lo : nullptr, gmode, ws + W.dApart, VGradFuse()));
        PSI_LAUNCHED_K(vg ? "lbs_vertex_bwd_fit" : "lbs_vertex_bwd");
    }
    {
        dim3 grid((unsigned)(m->Kpad / kDK), (unsigned)kNSplit, (unsigned)(W.Bpad / kBG));
        if (gmode == kGemmBf3) {
            const size_t smem = (size_t)kT3Smem + 1024;
            if (const int arc = ensure_max_dyn_smem(lbs_dcoef_bf3_kernel, smem)) return arc;
            const int items = (m->Kpad / kTM) * nsplit * (W.Bpad / kBG);
            (psi::skip_kernel("lbs_dcoef_bf3") ? cudaSuccess : launch_pdl(lbs_dcoef_bf3_kernel, dim3((unsigned)(items < PSI_NUM_SMS ? items : PSI_NUM_SMS)), dim3(kT3Threads),
                       smem, st, m->Kpad, m->Npad / kKC3, W.Bpad, m->basis_bwd3, reinterpret_cast<const unsigned short *>(ws + W.gvp), ws + W.part, nsplit));
            PSI_LAUNCHED_K("lbs_dcoef_bf3");
        } else if (tc5) {
            const size_t smem = (size_t)kTSmem + 1024;
            if (const int arc = ensure_max_dyn_smem(lbs_dcoef_tc5_kernel, smem)) return arc;
            const int items = (m->Kpad / kTM) * nsplit * (W.Bpad / kBG);
            (psi::skip_kernel("lbs_dcoef_tc5") ? cudaSuccess : launch_pdl(lbs_dcoef_tc5_kernel, dim3((unsigned)(items < PSI_NUM_SMS ? items : PSI_NUM_SMS)), dim3(kTThreads),
                       smem, st, m->Kpad, m->NC, W.Bpad, m->basis_bwd, ws + W.gvp, ws + W.gvp_lo, ws + W.part, nsplit));
            PSI_LAUNCHED_K("lbs_dcoef_tc5");
        } else {
            const size_t smem = (size_t)kDStages * kDStageBytes;
            if (const int arc = ensure_max_dyn_smem(lbs_dcoef_kernel, smem)) return arc;
            (psi::skip_kernel("lbs_dcoef") ? cudaSuccess : launch_pdl(lbs_dcoef_kernel, grid, dim3(256), smem, st, m->Kpad, m->NC, W.Bpad, m->basis_bwd, ws + W.gvp,
                       ws + W.part, kNSplit));
            PSI_LAUNCHED_K("lbs_dcoef");
        }
    }
    {
        const long c0 = (long)W.Bpad * m->Kpad, c1 = (long)B * (m->J + 1) * 12;
        dim3 grid((unsigned)(((c0 > c1 ? c0 : c1) + 255) / 256), 2);
        (psi::skip_kernel("lbs_reduce2") ? cudaSuccess : launch_pdl(lbs_reduce2_kernel, grid, dim3(256), 0, st, ws + W.part, nsplit, c0, ws + W.dsum, ws + W.dApart,
                   m->NCH, c1, ws + W.dA));
        PSI_LAUNCHED_K("lbs_reduce2");
    }
    (psi::skip_kernel("lbs_pose_bwd") ? cudaSuccess : launch_pdl(lbs_pose_bwd_kernel, dim3(B), dim3(128), 0, st, m->J, m->NB, m->P, m->Kpad, W.Bpad, nsplit, m->Jdirs,
               m->parents, pose, saved, L, ws + W.dA, ws + W.dsum, grad_joints, grad_betas,
               grad_pose, grad_transl, grad_rot, num_rot, rot6d, g6_root, g6A, g6_kpad, m->tp, m->Jdirs_p));
    PSI_LAUNCHED_K("lbs_pose_bwd");
    return PSI_OK;
}
}  // namespace psi

extern "C" {

int psi_lbs_bwd2(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                 const float *cam, long cam_bstride, const float *saved, const float *grad_verts,
                 const float *grad_joints, float *grad_betas, float *grad_pose, float *grad_transl,
                 float *grad_rot, int num_rot, void *workspace, size_t workspace_bytes,
                 psi_stream_t stream, psi_stream_t side_stream, void *ev_fork, void *ev_join) {
    (void)betas; (void)side_stream; (void)ev_fork; (void)ev_join;   // the backward is 3 launches on one stream now
    if (!grad_verts || (num_rot > 0 && !grad_rot)) return PSI_ERR_BAD_ARG;
    return psi::lbs_bwd_impl(m, B, pose, cam, cam_bstride, saved, grad_verts, nullptr, grad_joints, grad_betas,
                             grad_pose, grad_transl, grad_rot, num_rot, nullptr, nullptr, nullptr, 0, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int psi_lbs_bwd(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                const float *cam, long cam_bstride, const float *saved, const float *grad_verts,
                const float *grad_joints, float *grad_betas, float *grad_pose, float *grad_transl,
                float *grad_rot, int num_rot, void *workspace, size_t workspace_bytes,
                psi_stream_t stream) {
    return psi_lbs_bwd2(m, B, betas, pose, cam, cam_bstride, saved, grad_verts, grad_joints, grad_betas,
                        grad_pose, grad_transl, grad_rot, num_rot, workspace, workspace_bytes, stream,
                        nullptr, nullptr, nullptr);
}

}  // extern "C"

#ifdef PSI_TC5_TRACE
extern "C" __attribute__((visibility("default"))) int psi_debug_tc5_trace(unsigned long long *h_out) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(h_out, psi::g_tc5_trace, sizeof(unsigned long long) * 8 * 64);
}
#endif
