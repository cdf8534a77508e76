// SMPL-X linear blend skinning, forward and backward, for sm_100a.
//
// Replaces smplx.lbs.lbs (math vendored at human_body_prior/body_model/lbs.py:34-262),
// the translation of body_model.py:246-247 and GeometryTransformer.verts_transform
// (source/cvae.py:141-149), which the reference runs as ~25 torch kernels plus a 54-step
// python loop of 4x4 matmuls.
//
// HBM layout owned by the model handle (DESIGN.md "LBS"):
//   basis [Kpad][Npad]  rows 0..P-1 = posedirs (P=(J-1)*9), rows P..P+NB-1 = shapedirs^T,
//                       zero rows up to Kpad (multiple of 32); Npad = 3V rounded up to 96.
//                       Shape and pose blend shapes become ONE contraction over K = P+NB.
//   Jt [J,3], Jdirs [J,3,NB]  joint regressor folded with the template / shape basis
//                       (J = Jt + Jdirs*beta; the reference regresses from 10 475 vertices)
//   skin_j/skin_w [V][KW]  the non-zeros of the dense [V,J] skinning weights (ascending joint
//                       order, so the non-zero terms are summed in the reference's order)
//   jl_*                the same non-zeros bucketed by joint, for the backward gather-reduce
//
// Kernels:
//   lbs_pose_fwd     one CTA per body: Rodrigues, joints, kinematic chain, A = G - G*J,
//                    blend coefficients (pose feature | beta) in the vertex kernel's layout
//   lbs_vertex_fwd   CTA = 32 vertices x 32 bodies; the K-loop streams [32 x 96] basis tiles
//                    and [32 x 32] coefficient tiles through a 3-stage TMA (cp.async.bulk) +
//                    mbarrier ring; 48 FP32 accumulators per thread (3 coords x 16 bodies);
//                    epilogue = sparse skinning + translation + camera transform, all fused
//   lbs_vertex_bwd   d verts -> d v_posed (through T^T and the camera rotation)
//   lbs_dA           per (joint, body) gather-reduce of w * g (x) [v_posed;1] in fixed order
//   lbs_dcoef        split-N SIMT GEMM  d coef[b,k] = sum_n gvp[b,n] * basis[k,n]
//   lbs_pose_bwd     partial sums -> chain / Rodrigues backward -> d beta, d pose, d transl
// All reductions use fixed orders: results are bit-reproducible run to run.
#include "common.cuh"
#include <math.h>
#include <new>
#include <vector>

namespace psi {

constexpr int kMaxJ = 64;
constexpr int kKT = 16;         // basis rows per pipeline stage
constexpr int kTileN = 96;      // basis columns per CTA = 32 vertices
constexpr int kBG = 64;         // bodies per CTA (4 warps x 16)
constexpr int kStages = 6;        // 6 x 10 kB ring: 5 stages of prefetch cover the L2->smem latency
constexpr int kNSplit = 74;     // dcoef split of the N reduction: 4 k-tiles x 74 x (B/32) CTAs = 4 per SM at B=64

}  // namespace psi

struct psi_lbs_tree {          // kinematic tree by levels (root = level 0) + children lists (device)
    int nlev;
    const int *lvl_start, *lvl_joint, *child_start, *child_list;
};

struct psi_lbs_model {
    psi_lbs_tree tree;
    int *tree_buf;
    int V, J, NB, P, K, Kpad, Npad, KW;
    long nnz;
    float *basis, *v_template, *Jt, *Jdirs, *skin_w, *jl_w;
    int *skin_j, *parents, *jl_start, *jl_vert;
    size_t bytes;
};

namespace psi {

struct SavedLayout {
    size_t R, Jr, Gr, Gt, A, vp, coef, total;
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
    s.total = o;
    return s;
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
                    const psi_lbs_tree tree) {
    __shared__ float sR[kMaxJ * 9], sJ[kMaxJ * 3], sGr[kMaxJ * 9], sGt[kMaxJ * 3];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int j = tid; j < J; j += blockDim.x) {
        if (j < num_rot) {   // rotation matrices handed over directly (no axis-angle round trip)
#pragma unroll
            for (int e = 0; e < 9; ++e) sR[j * 9 + e] = rot_in[((size_t)b * num_rot + j) * 9 + e];
        } else {
            rodrigues(pose + ((size_t)b * J + j) * 3, sR + j * 9);
        }
    }
    for (int e = tid; e < J * 3; e += blockDim.x) {
        float v = Jt[e];
        for (int l = 0; l < NB; ++l) v = fmaf(Jdirs[(size_t)e * NB + l], betas[(size_t)b * NB + l], v);
        sJ[e] = v;
    }
    __syncthreads();
    // kinematic chain, one tree level at a time (joints of a level are independent)
    for (int l = 0; l < tree.nlev; ++l) {
        for (int q = tree.lvl_start[l] + tid; q < tree.lvl_start[l + 1]; q += blockDim.x) {
            const int j = tree.lvl_joint[q];
            if (j == 0) {
                for (int e = 0; e < 9; ++e) sGr[e] = sR[e];
                for (int e = 0; e < 3; ++e) sGt[e] = sJ[e];
                continue;
            }
            const int p = parents[j];
            const float *Gp = sGr + p * 9, *Rj = sR + j * 9;
            float rel[3] = {sJ[j * 3] - sJ[p * 3], sJ[j * 3 + 1] - sJ[p * 3 + 1],
                            sJ[j * 3 + 2] - sJ[p * 3 + 2]};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    sGr[j * 9 + r * 3 + c] =
                        Gp[r * 3] * Rj[c] + Gp[r * 3 + 1] * Rj[3 + c] + Gp[r * 3 + 2] * Rj[6 + c];
                sGt[j * 3 + r] =
                    Gp[r * 3] * rel[0] + Gp[r * 3 + 1] * rel[1] + Gp[r * 3 + 2] * rel[2] + sGt[p * 3 + r];
            }
        }
        __syncthreads();
    }
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
    // blend coefficients, layout [body group][Kpad][32]
    float *coef = saved + L.coef + (size_t)(b / kBG) * Kpad * kBG + (b % kBG);
    for (int k = tid; k < Kpad; k += blockDim.x) {
        float v = 0.f;
        if (k < P) {
            const int j = k / 9 + 1, e = k % 9;
            v = sR[j * 9 + e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
        } else if (k < P + NB) {
            v = betas[(size_t)b * NB + (k - P)];
        }
        coef[(size_t)k * kBG] = v;
    }
}

// coefficient rows of bodies past B inside the last body group must read as zero
__global__ void lbs_zero_coef_pad_kernel(float *coef, int B, int Kpad) {
    const int nbg = (B + kBG - 1) / kBG;
    const int first = B % kBG;
    if (first == 0) return;
    float *base = coef + (size_t)(nbg - 1) * Kpad * kBG;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Kpad * kBG; i += gridDim.x * blockDim.x)
        if ((i % kBG) >= first) base[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
struct VertexFwdParams {
    const float *basis, *v_template, *skin_w, *A, *coef, *transl, *cam;
    const int *skin_j;
    long cam_bstride;
    float *verts, *vp_out;
    int V, J, Kpad, Npad, KW, B;
};

constexpr int kBPT = 8;                   // bodies per thread (24 accumulators); kBG / kBPT warps per CTA
constexpr int kVfWarps = kBG / kBPT;      // 8 warps = 256 threads, 3 CTAs per SM: 17.7 warps/SM at B=64

__global__ void __launch_bounds__(kVfWarps * 32, 3) lbs_vertex_fwd_kernel(const VertexFwdParams p) {
    // 4 warps (16 bodies each, lane = vertex).  full[st]: bytes landed; empty[st]: all 4 warps
    // released the stage.  Lane 0 of warp 0 refills a stage kStages-2 chunks ahead, so there is no
    // block-wide barrier inside the K loop: warps may drift a chunk apart.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    constexpr int kStageFloats = kKT * kTileN + kKT * kBG;  // 2560 floats = 10 kB
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x, bg = blockIdx.y;
    const int nchunks = p.Kpad / kKT;
    // the basis is stored tile-major [tile][Kpad][96]: a stage is ONE contiguous 6 kB block
    const float *__restrict__ basis_t = p.basis + (size_t)tile * p.Kpad * kTileN;
    const float *__restrict__ coef_g = p.coef + (size_t)bg * p.Kpad * kBG;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kVfWarps); }
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int c) {   // warp 0, lane 0 only
        const int st = c % kStages;
        mbar_wait(&empty[st], (uint32_t)(((c / kStages) & 1) ^ 1));   // first use: passes at once
        float *dstB = sm + st * kStageFloats;
        mbar_arrive_expect_tx(&full[st], (uint32_t)(kStageFloats * 4));
        tma_load_1d(dstB, basis_t + (size_t)c * kKT * kTileN, kKT * kTileN * 4, &full[st]);
        tma_load_1d(dstB + kKT * kTileN, coef_g + (size_t)c * kKT * kBG, kKT * kBG * 4, &full[st]);
    };
    if (tid == 0)
        for (int c = 0; c < kStages - 2 && c < nchunks; ++c) issue(c);

    float acc[kBPT][3];
#pragma unroll
    for (int i = 0; i < kBPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + kStages - 2 < nchunks) issue(c + kStages - 2);
        __syncwarp();
        const int st = c % kStages;
        mbar_wait(&full[st], (uint32_t)((c / kStages) & 1));
        const float *bs = sm + st * kStageFloats + 3 * lane;
        const float4 *cs = reinterpret_cast<const float4 *>(sm + st * kStageFloats + kKT * kTileN) + (kBPT / 4) * w;  // [kk][64 bodies]
#pragma unroll 4
        for (int kk = 0; kk < kKT; ++kk) {
            const float b0 = bs[kk * kTileN], b1 = bs[kk * kTileN + 1], b2 = bs[kk * kTileN + 2];
            float cf[kBPT];
#pragma unroll
            for (int u = 0; u < kBPT / 4; ++u) {
                const float4 cc = cs[kk * (kBG / 4) + u];
                cf[4 * u] = cc.x; cf[4 * u + 1] = cc.y; cf[4 * u + 2] = cc.z; cf[4 * u + 3] = cc.w;
            }
#pragma unroll
            for (int i = 0; i < kBPT; ++i) {
                acc[i][0] = fmaf(cf[i], b0, acc[i][0]);
                acc[i][1] = fmaf(cf[i], b1, acc[i][1]);
                acc[i][2] = fmaf(cf[i], b2, acc[i][2]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);   // this warp is done with stage st
    }

    const int v = tile * 32 + lane;
    if (v >= p.V) return;
    const float t0 = p.v_template[3 * v], t1 = p.v_template[3 * v + 1], t2 = p.v_template[3 * v + 2];
    // phase A (unrolled, compile-time accumulator indices): v_posed = blend + template, stored once
    // (the backward needs it anyway); phase B (rolled): skinning reads it back, so the 24
    // accumulators are dead before the register-hungry part starts
#pragma unroll
    for (int i = 0; i < kBPT; ++i) {
        const int b = bg * kBG + w * kBPT + i;
        if (b >= p.B) continue;
        float *vp = p.vp_out + ((size_t)b * p.V + v) * 3;
        vp[0] = acc[i][0] + t0;
        vp[1] = acc[i][1] + t1;
        vp[2] = acc[i][2] + t2;
    }
}

// skinning + translation + camera transform, one thread per (body, vertex): a streaming pass over
// v_posed with gathers from the [B,J,12] transform table (169 kB at B=64, cache resident).
// (lbs.py:108-116, body_model.py:246-247, cvae.py:141-149)
__global__ void __launch_bounds__(256)
lbs_skin_fwd_kernel(int V, int J, int KW, const int *__restrict__ skin_j, const float *__restrict__ skin_w,
                    const float *__restrict__ A, const float *__restrict__ vp_in,
                    const float *__restrict__ transl, const float *__restrict__ cam, long cam_bstride,
                    float *__restrict__ verts) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (v >= V) return;
    const float *vp = vp_in + ((size_t)b * V + v) * 3;
    const float x = vp[0], y = vp[1], z = vp[2];
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    const float4 *__restrict__ Ab = reinterpret_cast<const float4 *>(A + (size_t)b * J * 12);
    for (int k = 0; k < KW; ++k) {
        const int j = skin_j[(size_t)v * KW + k];
        const float wt = skin_w[(size_t)v * KW + k];
        const float4 r0 = __ldg(Ab + j * 3), r1 = __ldg(Ab + j * 3 + 1), r2 = __ldg(Ab + j * 3 + 2);
        T[0] = fmaf(wt, r0.x, T[0]); T[1] = fmaf(wt, r0.y, T[1]); T[2] = fmaf(wt, r0.z, T[2]); T[3] = fmaf(wt, r0.w, T[3]);
        T[4] = fmaf(wt, r1.x, T[4]); T[5] = fmaf(wt, r1.y, T[5]); T[6] = fmaf(wt, r1.z, T[6]); T[7] = fmaf(wt, r1.w, T[7]);
        T[8] = fmaf(wt, r2.x, T[8]); T[9] = fmaf(wt, r2.y, T[9]); T[10] = fmaf(wt, r2.z, T[10]); T[11] = fmaf(wt, r2.w, T[11]);
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
    float *o = verts + ((size_t)b * V + v) * 3;
    o[0] = ox; o[1] = oy; o[2] = oz;
}

// ---------------------------------------------------------------------------------------------
// backward, vertex side: gw = Rc^T g ; gvp = Tr^T gw
__global__ void __launch_bounds__(256)
lbs_vertex_bwd_kernel(int V, int J, int KW, int Npad, int B, const int *__restrict__ skin_j,
                      const float *__restrict__ skin_w, const float *__restrict__ A,
                      const float *__restrict__ cam, long cam_bstride,
                      const float *__restrict__ gverts, float *__restrict__ gw_out,
                      float *__restrict__ gvp_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (v * 3 >= Npad) return;
    float *gvp = gvp_out + (size_t)b * Npad + (size_t)v * 3;
    if (v >= V) {
        gvp[0] = gvp[1] = gvp[2] = 0.f;
        return;
    }
    const float *g = gverts + ((size_t)b * V + v) * 3;
    float gx = g[0], gy = g[1], gz = g[2];
    if (cam) {
        const float *C = cam + (size_t)b * cam_bstride;
        const float x = C[0] * gx + C[4] * gy + C[8] * gz;
        const float y = C[1] * gx + C[5] * gy + C[9] * gz;
        const float z = C[2] * gx + C[6] * gy + C[10] * gz;
        gx = x; gy = y; gz = z;
    }
    float *gw = gw_out + ((size_t)b * V + v) * 3;
    gw[0] = gx; gw[1] = gy; gw[2] = gz;
    float T[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) T[e] = 0.f;
    const float4 *__restrict__ Ab = reinterpret_cast<const float4 *>(A + (size_t)b * J * 12);
    for (int k = 0; k < KW; ++k) {
        const int j = skin_j[(size_t)v * KW + k];
        const float wt = skin_w[(size_t)v * KW + k];
        const float4 r0 = __ldg(Ab + j * 3), r1 = __ldg(Ab + j * 3 + 1), r2 = __ldg(Ab + j * 3 + 2);
        T[0] = fmaf(wt, r0.x, T[0]); T[1] = fmaf(wt, r0.y, T[1]); T[2] = fmaf(wt, r0.z, T[2]);
        T[3] = fmaf(wt, r1.x, T[3]); T[4] = fmaf(wt, r1.y, T[4]); T[5] = fmaf(wt, r1.z, T[5]);
        T[6] = fmaf(wt, r2.x, T[6]); T[7] = fmaf(wt, r2.y, T[7]); T[8] = fmaf(wt, r2.z, T[8]);
    }
    gvp[0] = T[0] * gx + T[3] * gy + T[6] * gz;
    gvp[1] = T[1] * gx + T[4] * gy + T[7] * gz;
    gvp[2] = T[2] * gx + T[5] * gy + T[8] * gz;
}

// dA[b,j,:] = sum over the vertices skinned to j of w * [gw (x) vp | gw]; block (J, b) sums gw.
__global__ void __launch_bounds__(128)
lbs_dA_kernel(int V, int J, const int *__restrict__ jl_start, const int *__restrict__ jl_vert,
              const float *__restrict__ jl_w, const float *__restrict__ gw,
              const float *__restrict__ vp, float *__restrict__ dA, float *__restrict__ dtr) {
    const int j = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) acc[e] = 0.f;
    const float *gwb = gw + (size_t)b * V * 3, *vpb = vp + (size_t)b * V * 3;
    if (j < J) {
        const int s = jl_start[j], e_end = jl_start[j + 1];
        for (int e = s + tid; e < e_end; e += blockDim.x) {
            const int v = jl_vert[e];
            const float w = jl_w[e];
            const float g0 = w * gwb[v * 3], g1 = w * gwb[v * 3 + 1], g2 = w * gwb[v * 3 + 2];
            const float x = vpb[v * 3], y = vpb[v * 3 + 1], z = vpb[v * 3 + 2];
            acc[0] = fmaf(g0, x, acc[0]); acc[1] = fmaf(g0, y, acc[1]); acc[2] = fmaf(g0, z, acc[2]); acc[3] += g0;
            acc[4] = fmaf(g1, x, acc[4]); acc[5] = fmaf(g1, y, acc[5]); acc[6] = fmaf(g1, z, acc[6]); acc[7] += g1;
            acc[8] = fmaf(g2, x, acc[8]); acc[9] = fmaf(g2, y, acc[9]); acc[10] = fmaf(g2, z, acc[10]); acc[11] += g2;
        }
    } else {
        for (int v = tid; v < V; v += blockDim.x) {
            acc[0] += gwb[v * 3]; acc[1] += gwb[v * 3 + 1]; acc[2] += gwb[v * 3 + 2];
        }
    }
    __shared__ float red[4][12];
#pragma unroll
    for (int e = 0; e < 12; ++e) acc[e] = warp_sum(acc[e]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int e = 0; e < 12; ++e) red[tid >> 5][e] = acc[e];
    }
    __syncthreads();
    if (tid < 12) {
        const float s = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
        if (j < J) dA[((size_t)b * J + j) * 12 + tid] = s;
        else if (tid < 3) dtr[(size_t)b * 3 + tid] = s;
    }
}

// d coef partials: part[ns][b][k] = sum_{n in split ns} gvp[b][n] * basis[k][n]
// CTA: 128 k x 32 bodies, 128 threads, register tile 8 k x 4 bodies, n in chunks of 32.
__global__ void __launch_bounds__(128)
lbs_dcoef_kernel(int Kpad, int Npad, int B, int Bpad, const float *__restrict__ basis,
                 const float *__restrict__ gvp, float *__restrict__ part, int chunks_per_split) {
    __shared__ __align__(16) float BsT[32][132];
    __shared__ __align__(16) float GT[32][36];
    const int tid = threadIdx.x, tk = tid & 15, tb = tid >> 4;
    const int kbase = blockIdx.x * 128, ns = blockIdx.y, bbase = blockIdx.z * 32;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const int total_chunks = Npad / 32;
    const int c_begin = (int)((long)ns * total_chunks / chunks_per_split);           // balanced partition (chunks_per_split = number of splits)
    const int c_end = (int)((long)(ns + 1) * total_chunks / chunks_per_split);
    // software pipeline: the next chunk's global loads are in flight while this one is computed.
    // store mapping: lanes run over 16 rows (k or body) x 2 column groups -> conflict-free
    // transposed stores ((4n + row) % 32 covers all banks).
    const int lrow = tid & 15, col4 = tid >> 4;
    float4 pb[8], pg[2];
    auto fetch = [&](int c) {
        const int n0 = c * 32;
        const float *bt = basis + (size_t)(n0 / kTileN) * Kpad * kTileN + (n0 % kTileN) + col4 * 4;
#pragma unroll
        for (int r = 0; r < 8; ++r)
            pb[r] = __ldg(reinterpret_cast<const float4 *>(bt + (size_t)(kbase + r * 16 + lrow) * kTileN));
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = r * 16 + lrow;
            pg[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bbase + row < B)
                pg[r] = __ldg(reinterpret_cast<const float4 *>(gvp + (size_t)(bbase + row) * Npad + n0 + col4 * 4));
        }
    };
    if (c_begin < c_end) fetch(c_begin);
    for (int c = c_begin; c < c_end; ++c) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = r * 16 + lrow;
            BsT[col4 * 4 + 0][row] = pb[r].x; BsT[col4 * 4 + 1][row] = pb[r].y;
            BsT[col4 * 4 + 2][row] = pb[r].z; BsT[col4 * 4 + 3][row] = pb[r].w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = r * 16 + lrow;
            GT[col4 * 4 + 0][row] = pg[r].x; GT[col4 * 4 + 1][row] = pg[r].y;
            GT[col4 * 4 + 2][row] = pg[r].z; GT[col4 * 4 + 3][row] = pg[r].w;
        }
        __syncthreads();
        if (c + 1 < c_end) fetch(c + 1);
#pragma unroll 8
        for (int n = 0; n < 32; ++n) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&BsT[n][4 * tk]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&BsT[n][64 + 4 * tk]);
            const float4 g = *reinterpret_cast<const float4 *>(&GT[n][4 * tb]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i], gg[jj], acc[i][jj]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int b = bbase + tb * 4 + jj;
        if (b >= Bpad) continue;
        float *o = part + ((size_t)ns * Bpad + b) * Kpad + kbase;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[4 * tk + i] = acc[i][jj];
            o[64 + 4 * tk + i] = acc[4 + i][jj];
        }
    }
}

// dsum[b][k] = sum over the N-splits, in a fixed order (4 interleaved partial sums)
__global__ void __launch_bounds__(256)
lbs_dcoef_reduce_kernel(const float *__restrict__ part, int nsplit, long per_split, float *__restrict__ dsum) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_split) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int ns = 0;
    for (; ns + 4 <= nsplit; ns += 4) {
        s0 += part[(size_t)ns * per_split + i];
        s1 += part[(size_t)(ns + 1) * per_split + i];
        s2 += part[(size_t)(ns + 2) * per_split + i];
        s3 += part[(size_t)(ns + 3) * per_split + i];
    }
    for (; ns < nsplit; ++ns) s0 += part[(size_t)ns * per_split + i];
    dsum[i] = (s0 + s1) + (s2 + s3);
}

// per body: run the chain and Rodrigues backward on the reduced d coef
__global__ void __launch_bounds__(128)
lbs_pose_bwd_kernel(int J, int NB, int P, int Kpad, int Bpad, int nsplit,
                    const float *__restrict__ Jdirs, const int *__restrict__ parents,
                    const float *__restrict__ pose, const float *__restrict__ saved, SavedLayout L,
                    const float *__restrict__ dA, const float *__restrict__ dtr,
                    const float *__restrict__ part, const float *__restrict__ gjoints,
                    float *__restrict__ gbetas, float *__restrict__ gpose,
                    float *__restrict__ gtransl, float *__restrict__ grot, int num_rot,
                    const psi_lbs_tree tree) {
    __shared__ float sR[kMaxJ * 9], sJ[kMaxJ * 3], sGr[kMaxJ * 9], drel_s[kMaxJ * 3];
    __shared__ float dGr[kMaxJ * 9], dGt[kMaxJ * 3], dR[kMaxJ * 9], dJ[kMaxJ * 3];
    __shared__ float dbeta_direct[64];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *iR = saved + L.R + (size_t)b * J * 9, *iJ = saved + L.Jr + (size_t)b * J * 3;
    const float *iGr = saved + L.Gr + (size_t)b * J * 9;
    for (int e = tid; e < J * 9; e += blockDim.x) { sR[e] = iR[e]; sGr[e] = iGr[e]; }
    for (int e = tid; e < J * 3; e += blockDim.x) sJ[e] = iJ[e];
    // d pose-feature (added to dR below) and the direct d beta, summed over splits in order
    for (int k = tid; k < P + NB; k += blockDim.x) {
        const float s = part[(size_t)b * Kpad + k];          // already summed over the splits
        (void)nsplit; (void)Bpad;
        if (k < P) dR[9 + k] = s;           // joint j = k/9+1, entry k%9
        else dbeta_direct[k - P] = s;
    }
    if (tid < 9) dR[tid] = 0.f;
    __syncthreads();
    // dGr = dAr - dAt J^T ; dGt = dAt (+ d posed joints) ; dJ = -Gr^T dAt
    for (int j = tid; j < J; j += blockDim.x) {
        const float *a = dA + ((size_t)b * J + j) * 12;
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
    // chain backward, deepest tree level first.  A joint first PULLS from its children (fixed
    // child order: deterministic, race free), then finishes its own dR / d rel.
    for (int l = tree.nlev - 1; l >= 0; --l) {
        for (int q = tree.lvl_start[l] + tid; q < tree.lvl_start[l + 1]; q += blockDim.x) {
            const int j = tree.lvl_joint[q];
            float g[9], gt[3], dj[3];
#pragma unroll
            for (int e = 0; e < 9; ++e) g[e] = dGr[j * 9 + e];
#pragma unroll
            for (int e = 0; e < 3; ++e) { gt[e] = dGt[j * 3 + e]; dj[e] = dJ[j * 3 + e]; }
            for (int ci = tree.child_start[j]; ci < tree.child_start[j + 1]; ++ci) {
                const int c = tree.child_list[ci];
                const float *Rc = sR + c * 9, *gc = dGr + c * 9, *gtc = dGt + c * 3;
                const float rel[3] = {sJ[c * 3] - sJ[j * 3], sJ[c * 3 + 1] - sJ[j * 3 + 1], sJ[c * 3 + 2] - sJ[j * 3 + 2]};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)   // dGr[j] += dGr[c] Rc^T + dGt[c] rel^T
                        g[r * 3 + cc] += gc[r * 3] * Rc[cc * 3] + gc[r * 3 + 1] * Rc[cc * 3 + 1] +
                                         gc[r * 3 + 2] * Rc[cc * 3 + 2] + gtc[r] * rel[cc];
                    gt[r] += gtc[r];
                    dj[r] -= drel_s[c * 3 + r];
                }
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) dGr[j * 9 + e] = g[e];
            if (j == 0) {
#pragma unroll
                for (int e = 0; e < 9; ++e) dR[e] += g[e];
#pragma unroll
                for (int e = 0; e < 3; ++e) dj[e] += gt[e];
            } else {
                const float *Gp = sGr + parents[j] * 9;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)   // dR[j] += Gp^T dGr[j]
                        dR[j * 9 + r * 3 + cc] += Gp[r] * g[cc] + Gp[3 + r] * g[3 + cc] + Gp[6 + r] * g[6 + cc];
                    const float drel = Gp[r] * gt[0] + Gp[3 + r] * gt[1] + Gp[6 + r] * gt[2];
                    drel_s[j * 3 + r] = drel;
                    dj[r] += drel;
                }
            }
#pragma unroll
            for (int e = 0; e < 3; ++e) { dGt[j * 3 + e] = gt[e]; dJ[j * 3 + e] = dj[e]; }
        }
        __syncthreads();
    }
    // Rodrigues backward (lbs.py:177-191); joints given as matrices export dR itself
    for (int j = tid; j < J; j += blockDim.x) {
        float *o = gpose + ((size_t)b * J + j) * 3;
        if (j < num_rot) {
#pragma unroll
            for (int e = 0; e < 9; ++e) grot[((size_t)b * num_rot + j) * 9 + e] = dR[j * 9 + e];
            o[0] = o[1] = o[2] = 0.f;
            continue;
        }
        const float *r = pose + ((size_t)b * J + j) * 3;
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
    for (int l = tid; l < NB; l += blockDim.x) {
        float s = dbeta_direct[l];
        for (int e = 0; e < J * 3; ++e) s = fmaf(Jdirs[(size_t)e * NB + l], dJ[e], s);
        gbetas[(size_t)b * NB + l] = s;
    }
    if (gtransl && tid < 3) {
        float s = dtr[(size_t)b * 3 + tid];
        if (gjoints)
            for (int j = 0; j < J; ++j) s += gjoints[((size_t)b * J + j) * 3 + tid];
        gtransl[(size_t)b * 3 + tid] = s;
    }
}

struct BwdLayout {
    size_t gw, gvp, dA, dtr, part, dsum, total;  // in floats
    int Bpad;
};
static BwdLayout bwd_layout(const psi_lbs_model *m, int B) {
    BwdLayout l;
    l.Bpad = ((B + 31) / 32) * 32;
    size_t o = 0;
    l.gw = o;   o += (size_t)B * m->V * 3;
    o = (o + 3) & ~(size_t)3;
    l.gvp = o;  o += (size_t)l.Bpad * m->Npad;
    l.dA = o;   o += (size_t)B * m->J * 12;
    l.dtr = o;  o += (size_t)B * 3;
    o = (o + 3) & ~(size_t)3;
    l.part = o; o += (size_t)kNSplit * l.Bpad * m->Kpad;
    l.dsum = o; o += (size_t)l.Bpad * m->Kpad;
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
    cudaFree(m->basis); cudaFree(m->v_template); cudaFree(m->Jt); cudaFree(m->Jdirs);
    cudaFree(m->skin_w); cudaFree(m->jl_w); cudaFree(m->skin_j); cudaFree(m->parents);
    cudaFree(m->jl_start); cudaFree(m->jl_vert); cudaFree(m->tree_buf);
    delete m;
}

int psi_lbs_model_create(psi_lbs_model **out, int V, int J, int NB, const float *h_v_template,
                         const float *h_shapedirs, const float *h_posedirs,
                         const float *h_J_regressor, const float *h_weights, const int *h_parents,
                         psi_stream_t stream) {
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
    m->Kpad = ((m->K + 127) / 128) * 128;   // multiple of the dcoef k tile (and of kKT)
    m->Npad = ((3 * V + kTileN - 1) / kTileN) * kTileN;
    m->bytes = 0;
    const int P = m->P, Kpad = m->Kpad, Npad = m->Npad;
    const size_t N = (size_t)3 * V;

    // tile-major: element (k, n) lives at [n / 96][k][n % 96]
    std::vector<float> basis((size_t)Kpad * Npad, 0.f);
    auto bidx = [&](int k, size_t n) { return (n / kTileN) * (size_t)Kpad * kTileN + (size_t)k * kTileN + (n % kTileN); };
    for (int k = 0; k < P; ++k)
        for (size_t n = 0; n < N; ++n) basis[bidx(k, n)] = h_posedirs[(size_t)k * N + n];
    for (int l = 0; l < NB; ++l)
        for (size_t n = 0; n < N; ++n) basis[bidx(P + l, n)] = h_shapedirs[n * NB + l];
    std::vector<float> vt((size_t)Npad, 0.f);
    for (size_t n = 0; n < N; ++n) vt[n] = h_v_template[n];

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
    std::vector<int> count(J + 1, 0);
    for (int v = 0; v < V; ++v) {
        int c = 0;
        for (int j = 0; j < J; ++j) {
            const float w = h_weights[(size_t)v * J + j];
            if (w != 0.f) {
                skin_j[(size_t)v * KW + c] = j;
                skin_w[(size_t)v * KW + c] = w;
                ++c;
                ++count[j + 1];
            }
        }
    }
    std::vector<int> jl_start(J + 1, 0);
    for (int j = 0; j < J; ++j) jl_start[j + 1] = jl_start[j] + count[j + 1];
    m->nnz = jl_start[J];
    std::vector<int> jl_vert((size_t)m->nnz), fill(jl_start.begin(), jl_start.end() - 1);
    std::vector<float> jl_w((size_t)m->nnz);
    for (int v = 0; v < V; ++v)
        for (int j = 0; j < J; ++j) {
            const float w = h_weights[(size_t)v * J + j];
            if (w != 0.f) {
                jl_vert[fill[j]] = v;
                jl_w[fill[j]] = w;
                ++fill[j];
            }
        }
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
    if (rc == PSI_OK) rc = upload(&m->basis, basis, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->v_template, vt, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->Jt, Jt, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->Jdirs, Jdirs, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->skin_j, skin_j, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->skin_w, skin_w, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->parents, parents, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->jl_start, jl_start, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->jl_vert, jl_vert, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->jl_w, jl_w, st, &m->bytes);
    if (rc == PSI_OK) rc = upload(&m->tree_buf, treebuf, st, &m->bytes);
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

size_t psi_lbs_saved_floats(const psi_lbs_model *m, int B) {
    if (!m || B <= 0) return 0;
    return psi::saved_layout(B, m->J, m->V, m->Kpad).total;
}

int psi_lbs_fwd(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                const float *transl, const float *cam, long cam_bstride, const float *rot_in,
                int num_rot, float *verts, float *joints, float *saved, psi_stream_t stream) {
    using namespace psi;
    if (!m || B < 0) return PSI_ERR_BAD_ARG;
    if (B == 0) return PSI_OK;
    if (!betas || !pose || !verts || !saved) return PSI_ERR_BAD_ARG;
    if (num_rot < 0 || num_rot > m->J || (num_rot > 0 && !rot_in)) return PSI_ERR_BAD_ARG;
    if (B > 65535) return PSI_ERR_UNSUPPORTED;
    if (((uintptr_t)saved & 127u) != 0) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const SavedLayout L = saved_layout(B, m->J, m->V, m->Kpad);
    lbs_pose_fwd_kernel<<<B, 128, 0, st>>>(m->J, m->NB, m->P, m->Kpad, m->Jt, m->Jdirs, m->parents,
                                          B, betas, pose, transl, saved, L, joints, rot_in, num_rot, m->tree);
    PSI_LAUNCHED();
    if (B % kBG) {
        lbs_zero_coef_pad_kernel<<<8, 256, 0, st>>>(saved + L.coef, B, m->Kpad);
        PSI_LAUNCHED();
    }
    VertexFwdParams p;
    p.basis = m->basis; p.v_template = m->v_template; p.skin_w = m->skin_w; p.skin_j = m->skin_j;
    p.A = saved + L.A; p.coef = saved + L.coef; p.transl = transl; p.cam = cam;
    p.cam_bstride = cam_bstride; p.verts = verts; p.vp_out = saved + L.vp;
    p.V = m->V; p.J = m->J; p.Kpad = m->Kpad; p.Npad = m->Npad; p.KW = m->KW; p.B = B;
    const size_t smem = (size_t)kStages * (kKT * kTileN + kKT * kBG) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(lbs_vertex_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    dim3 grid((unsigned)(m->Npad / kTileN), (unsigned)((B + kBG - 1) / kBG));
    lbs_vertex_fwd_kernel<<<grid, kVfWarps * 32, smem, st>>>(p);
    PSI_LAUNCHED();
    {
        dim3 sgrid((unsigned)((m->V + 255) / 256), (unsigned)B);
        lbs_skin_fwd_kernel<<<sgrid, 256, 0, st>>>(m->V, m->J, m->KW, m->skin_j, m->skin_w, saved + L.A, saved + L.vp,
                                                   transl, cam, cam_bstride, verts);
        PSI_LAUNCHED();
    }
    return PSI_OK;
}

int psi_lbs_bwd(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                const float *cam, long cam_bstride, const float *saved, const float *grad_verts,
                const float *grad_joints, float *grad_betas, float *grad_pose, float *grad_transl,
                float *grad_rot, int num_rot, void *workspace, size_t workspace_bytes,
                psi_stream_t stream);

size_t psi_lbs_bwd_workspace_bytes(const psi_lbs_model *m, int B) {
    if (!m || B <= 0) return 0;
    return psi::bwd_layout(m, B).total * sizeof(float);
}

int psi_lbs_bwd2(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                 const float *cam, long cam_bstride, const float *saved, const float *grad_verts,
                 const float *grad_joints, float *grad_betas, float *grad_pose, float *grad_transl,
                 float *grad_rot, int num_rot, void *workspace, size_t workspace_bytes,
                 psi_stream_t stream, psi_stream_t side_stream, void *ev_fork, void *ev_join) {
    using namespace psi;
    (void)betas;
    // optional side stream: lbs_dA runs next to lbs_dcoef (both only depend on lbs_vertex_bwd)
    cudaStream_t side = (side_stream && ev_fork && ev_join) ? (cudaStream_t)side_stream : nullptr;
    if (!m || B < 0) return PSI_ERR_BAD_ARG;
    if (B == 0) return PSI_OK;
    if (!pose || !saved || !grad_verts || !grad_betas || !grad_pose || !workspace) return PSI_ERR_BAD_ARG;
    if (num_rot < 0 || num_rot > m->J || (num_rot > 0 && !grad_rot)) return PSI_ERR_BAD_ARG;
    if (B > 65535) return PSI_ERR_UNSUPPORTED;
    const BwdLayout W = bwd_layout(m, B);
    if (workspace_bytes < W.total * sizeof(float)) return PSI_ERR_WORKSPACE;
    if (((uintptr_t)workspace & 15u) != 0) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    float *ws = reinterpret_cast<float *>(workspace);
    const SavedLayout L = saved_layout(B, m->J, m->V, m->Kpad);
    {
        dim3 grid((unsigned)((m->Npad / 3 + 255) / 256), (unsigned)B);
        lbs_vertex_bwd_kernel<<<grid, 256, 0, st>>>(m->V, m->J, m->KW, m->Npad, B, m->skin_j,
                                                    m->skin_w, saved + L.A, cam, cam_bstride,
                                                    grad_verts, ws + W.gw, ws + W.gvp);
        PSI_LAUNCHED();
    }
    {
        cudaStream_t sa = st;
        if (side) {
            if (cudaEventRecord((cudaEvent_t)ev_fork, st) != cudaSuccess ||
                cudaStreamWaitEvent(side, (cudaEvent_t)ev_fork, 0) != cudaSuccess)
                return PSI_ERR_BAD_ARG;
            sa = side;
        }
        dim3 grid((unsigned)(m->J + 1), (unsigned)B);
        lbs_dA_kernel<<<grid, 128, 0, sa>>>(m->V, m->J, m->jl_start, m->jl_vert, m->jl_w,
                                            ws + W.gw, saved + L.vp, ws + W.dA, ws + W.dtr);
        PSI_LAUNCHED();
        if (side && cudaEventRecord((cudaEvent_t)ev_join, side) != cudaSuccess) return PSI_ERR_BAD_ARG;
    }
    {
        const int total_chunks = m->Npad / 32;
        const int cps = kNSplit;
        dim3 grid((unsigned)(m->Kpad / 128), (unsigned)kNSplit, (unsigned)(W.Bpad / 32));
        lbs_dcoef_kernel<<<grid, 128, 0, st>>>(m->Kpad, m->Npad, B, W.Bpad, m->basis, ws + W.gvp,
                                               ws + W.part, cps);
        PSI_LAUNCHED();
        const long per_split = (long)W.Bpad * m->Kpad;
        lbs_dcoef_reduce_kernel<<<(unsigned)((per_split + 255) / 256), 256, 0, st>>>(ws + W.part, kNSplit, per_split, ws + W.dsum);
        PSI_LAUNCHED();
    }
    if (side && cudaStreamWaitEvent(st, (cudaEvent_t)ev_join, 0) != cudaSuccess) return PSI_ERR_BAD_ARG;
    lbs_pose_bwd_kernel<<<B, 128, 0, st>>>(m->J, m->NB, m->P, m->Kpad, W.Bpad, kNSplit, m->Jdirs,
                                          m->parents, pose, saved, L, ws + W.dA, ws + W.dtr,
                                          ws + W.dsum, grad_joints, grad_betas, grad_pose,
                                          grad_transl, grad_rot, num_rot, m->tree);
    PSI_LAUNCHED();
    return PSI_OK;
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
