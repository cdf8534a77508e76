// Fused scene-fitting loop: the whole iteration of FittingOP.cal_loss + backward + Adam
// (source/fitting_habitat.py:103-164,177-191) as 11 kernel launches, captured once in a CUDA graph.
//
//   fit_prologue   per body: xhr_rec[75] -> translation, 6D -> R (cvae.py:46-55), betas, VPoser
//                  decode (MLP 32->512->512->126, leaky 0.2, vposer_smpl.py:107-121) -> 21 x 6D -> R,
//                  hand PCA + pose_mean (smplx), L_rec and L_vposer of this body
//   psi_lbs_fwd    (lbs.cu) joints 0..21 take R directly: the reference's R -> axis-angle
//                  (torchgeometry) -> Rodrigues round trip is the identity on SO(3) up to rounding,
//                  and so is its Jacobian on the tangent directions that reach it
//   psi_nn_index_query / psi_sdf_fwd     contact NN distance, SDF value + gradient + partials
//   fit_vertex_grad  dL/dverts of the contact robustifier (fitting_habitat.py:141) and of the
//                  collision mean (:155-160) + per-chunk contact-loss partials
//   psi_lbs_bwd    -> d betas, d hand axis-angles, d translation, d R (joints 0..21)
//   fit_epilogue   per body: Gram-Schmidt / MLP / PCA backward, loss-term gradients, Adam step
//                  (torch.optim.Adam defaults, fitting_habitat.py:76), loss values
// Loss semantics: the SUM over bodies of the reference's B=1 loss ('independent' mode) -- every
// body is optimised exactly as the shipped batch_size-1 scripts do, bodies never interact.
// All reductions have a fixed order: results are bit-reproducible and independent of B.
#include "common.cuh"
#include <math.h>
#include <new>
#include <vector>

struct psi_fit_ctx {
    const psi_lbs_model *model;
    const psi_nn_index *index;
    psi_fit_config cfg;
    int B, V, J, NB, latent, hidden, nbody, ncomp, num_rot, nu, num_contact, D, np_sdf, nchunk;
    // constants
    float *W1, *W1T, *b1, *W2, *W2T, *b2, *W3, *W3T, *W3Tp, *b3, *hand_l, *hand_r, *pose_mean, *cweight;
    int *csel, *cslot;
    const float *sdf, *scene_pts;
    float gmin[3], gmax[3];
    // state + scratch
    float *x0, *x, *am, *av, *cam, *rot, *pose, *shape, *transl, *h1pre, *h2pre, *o6, *verts, *saved,
        *sdfv, *sdfg, *partial, *nnd, *gverts, *cpart, *gshape, *gpose, *grot, *gtransl, *lbs_ws, *losses;
    int *nni, *step, *nnhint;
    size_t lbs_ws_bytes;
    cudaGraphExec_t exec;
    int pending_join;
    cudaStream_t gstream;          // graphs cannot be captured on the legacy default stream
    cudaEvent_t ev_in, ev_out;
    std::vector<void *> owned;
};

namespace psi {

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.2f * v; }
__device__ __forceinline__ float lrelu_grad(float pre) { return pre > 0.f ? 1.0f : 0.2f; }

// cvae.py:46-55: x6 viewed [3,2]; columns a1 = (x0,x2,x4), a2 = (x1,x3,x5)
__device__ void gs_fwd(const float *x6, float *R) {
    const float a1[3] = {x6[0], x6[2], x6[4]}, a2[3] = {x6[1], x6[3], x6[5]};
    const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    const float s = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const float u[3] = {a2[0] - s * b1[0], a2[1] - s * b1[1], a2[2] - s * b1[2]};
    const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        R[r * 3 + 0] = b1[r];
        R[r * 3 + 1] = b2[r];
        R[r * 3 + 2] = b3[r];
    }
}

__device__ void gs_bwd(const float *x6, const float *dR, float *dx6) {
    const float a1[3] = {x6[0], x6[2], x6[4]}, a2[3] = {x6[1], x6[3], x6[5]};
    const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    const float s = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const float u[3] = {a2[0] - s * b1[0], a2[1] - s * b1[1], a2[2] - s * b1[2]};
    const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    float db1[3] = {dR[0], dR[3], dR[6]}, db2[3] = {dR[1], dR[4], dR[7]};
    const float db3[3] = {dR[2], dR[5], dR[8]};
    // b3 = b1 x b2
    db1[0] += b2[1] * db3[2] - b2[2] * db3[1];
    db1[1] += b2[2] * db3[0] - b2[0] * db3[2];
    db1[2] += b2[0] * db3[1] - b2[1] * db3[0];
    db2[0] += db3[1] * b1[2] - db3[2] * b1[1];
    db2[1] += db3[2] * b1[0] - db3[0] * b1[2];
    db2[2] += db3[0] * b1[1] - db3[1] * b1[0];
    // b2 = u / |u|
    const float p2 = b2[0] * db2[0] + b2[1] * db2[1] + b2[2] * db2[2];
    const float du[3] = {(db2[0] - p2 * b2[0]) / n2, (db2[1] - p2 * b2[1]) / n2, (db2[2] - p2 * b2[2]) / n2};
    // u = a2 - (b1.a2) b1
    const float dub1 = du[0] * b1[0] + du[1] * b1[1] + du[2] * b1[2];
    const float da2[3] = {du[0] - dub1 * b1[0], du[1] - dub1 * b1[1], du[2] - dub1 * b1[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r) db1[r] += -s * du[r] - dub1 * a2[r];
    // b1 = a1 / |a1|
    const float p1 = b1[0] * db1[0] + b1[1] * db1[1] + b1[2] * db1[2];
    const float da1[3] = {(db1[0] - p1 * b1[0]) / n1, (db1[1] - p1 * b1[1]) / n1, (db1[2] - p1 * b1[2]) / n1};
    dx6[0] = da1[0]; dx6[2] = da1[1]; dx6[4] = da1[2];
    dx6[1] = da2[0]; dx6[3] = da2[1]; dx6[5] = da2[2];
}

struct FitDims {
    int B, V, J, NB, latent, hidden, nbody, ncomp, num_rot;
};

}  // namespace psi
#include "fit_mlp.cuh"
namespace psi {

// x layout [75]: t3 | 6D | betas10 | z32 | lh12 | rh12   (cvae.py:28-33 after convert_to_6D_rot)
__global__ void __launch_bounds__(256)
fit_prologue_kernel(FitDims d, const float *__restrict__ x0, const float *__restrict__ x,
                    const float *__restrict__ W1T, const float *__restrict__ b1,
                    const float *__restrict__ W2T, const float *__restrict__ b2,
                    const float *__restrict__ W3T, const float *__restrict__ b3,
                    const float *__restrict__ hand_l, const float *__restrict__ hand_r,
                    const float *__restrict__ pose_mean, float w_rec, float w_vp,
                    float *__restrict__ rot, float *__restrict__ pose, float *__restrict__ shape,
                    float *__restrict__ transl, float *__restrict__ h1pre, float *__restrict__ h2pre,
                    float *__restrict__ o6, float *__restrict__ losses) {
    __shared__ float sx[80], sh1[512], sh2[512], so[128];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int H = d.hidden, Lz = d.latent, NO = d.nbody * 6;
    const int xdim = 9 + 10 + Lz + 2 * d.ncomp;   // 75
    const int zoff = 19, lhoff = 19 + Lz, rhoff = lhoff + d.ncomp;
    for (int e = tid; e < xdim; e += blockDim.x) sx[e] = x[(size_t)b * xdim + e];
    __syncthreads();
    for (int j = tid; j < H; j += blockDim.x) {
        float a = b1[j];
        for (int i = 0; i < Lz; ++i) a = fmaf(W1T[(size_t)i * H + j], sx[zoff + i], a);
        h1pre[(size_t)b * H + j] = a;
        sh1[j] = lrelu(a);
    }
    __syncthreads();
    for (int j = tid; j < H; j += blockDim.x) {
        float a0 = b2[j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int i = 0; i < H; i += 4) {
            a0 = fmaf(W2T[(size_t)i * H + j], sh1[i], a0);
            a1 = fmaf(W2T[(size_t)(i + 1) * H + j], sh1[i + 1], a1);
            a2 = fmaf(W2T[(size_t)(i + 2) * H + j], sh1[i + 2], a2);
            a3 = fmaf(W2T[(size_t)(i + 3) * H + j], sh1[i + 3], a3);
        }
        const float a = (a0 + a1) + (a2 + a3);
        h2pre[(size_t)b * H + j] = a;
        sh2[j] = lrelu(a);
    }
    __syncthreads();
    for (int k = tid; k < NO; k += blockDim.x) {
        float a0 = b3[k], a1 = 0.f;
        for (int i = 0; i < H; i += 2) {
            a0 = fmaf(W3T[(size_t)i * NO + k], sh2[i], a0);
            a1 = fmaf(W3T[(size_t)(i + 1) * NO + k], sh2[i + 1], a1);
        }
        const float a = a0 + a1;
        so[k] = a;
        o6[(size_t)b * NO + k] = a;
    }
    __syncthreads();
    // rotations of joints 0..nbody: global orientation from x[3:9], body joints from the decoder
    if (tid <= d.nbody) {
        float R[9];
        gs_fwd(tid == 0 ? sx + 3 : so + (tid - 1) * 6, R);
#pragma unroll
        for (int e = 0; e < 9; ++e) rot[((size_t)b * d.num_rot + tid) * 9 + e] = R[e];
    }
    // axis-angle pose vector [J*3]: only joints >= num_rot are read by the LBS kernels
    const int hl0 = (d.J - 30) * 3, hr0 = (d.J - 15) * 3;
    for (int e = tid; e < d.J * 3; e += blockDim.x) {
        float v = pose_mean[e];
        if (e >= hr0) {
            for (int c = 0; c < d.ncomp; ++c) v = fmaf(sx[rhoff + c], hand_r[c * 45 + (e - hr0)], v);
        } else if (e >= hl0) {
            for (int c = 0; c < d.ncomp; ++c) v = fmaf(sx[lhoff + c], hand_l[c * 45 + (e - hl0)], v);
        } else if (e < d.num_rot * 3) {
            v = 0.f;
        }
        pose[(size_t)b * d.J * 3 + e] = v;
    }
    for (int e = tid; e < d.NB; e += blockDim.x) shape[(size_t)b * d.NB + e] = e < 10 ? sx[9 + e] : 0.f;
    if (tid < 3) transl[(size_t)b * 3 + tid] = sx[tid];
    if (tid == 0) {
        float r = 0.f, zz = 0.f;
        for (int e = 0; e < xdim; ++e) r += fabsf(x0[(size_t)b * xdim + e] - sx[e]);
        for (int i = 0; i < Lz; ++i) zz = fmaf(sx[zoff + i], sx[zoff + i], zz);
        losses[(size_t)b * 4 + 0] = w_rec * (r / (float)xdim);
        losses[(size_t)b * 4 + 1] = w_vp * (zz / (float)Lz);
    }
}

// dL/dverts (scene frame) of the contact and collision terms of body b
__global__ void __launch_bounds__(256)
fit_vertex_grad_kernel(int V, int nu, int np_sdf, int num_contact, const float *__restrict__ verts,
                       const float *__restrict__ scene, const float *__restrict__ sdfv,
                       const float *__restrict__ sdfg, const float *__restrict__ partial,
                       const float *__restrict__ nnd, const int *__restrict__ nni,
                       const int *__restrict__ cslot, const float *__restrict__ cweight, float w_contact,
                       float w_coll, float robust_c, float *__restrict__ gverts,
                       float *__restrict__ cpart) {
    __shared__ float s_cnt;
    __shared__ float red[8];
    const int b = blockIdx.y, v = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) {
        float c = 0.f;
        for (int i = 0; i < np_sdf; ++i) c += partial[((size_t)b * np_sdf + i) * 2 + 1];
        s_cnt = c;
    }
    __syncthreads();
    float closs = 0.f;
    if (v < V) {
        const size_t o = (size_t)b * V + v;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (sdfv[o] < 0.f) {   // d/dv [ w * sum(-sdf)/cnt ]
            const float k = -w_coll / s_cnt;
            gx = k * sdfg[o * 3];
            gy = k * sdfg[o * 3 + 1];
            gz = k * sdfg[o * 3 + 2];
        }
        const int slot = cslot[v];
        if (slot >= 0) {
            const float dd = nnd[(size_t)b * nu + slot];
            const float s = sqrtf(dd + 1e-4f);
            const float den = s + robust_c;
            const float wgt = cweight[v];
            closs = wgt * (s / den);
            // d/dd [ s/(s+c) ] = c/(s+c)^2 * 1/(2s);   d dd/dp = 2 (p - q)   (chamfer.cu:165-168)
            const float gd = (w_contact / (float)num_contact) * wgt * (robust_c / (den * den)) * (0.5f / s);
            const float g2 = gd * 2.0f;
            const float *q = scene + (size_t)nni[(size_t)b * nu + slot] * 3;
            gx += g2 * (verts[o * 3] - q[0]);
            gy += g2 * (verts[o * 3 + 1] - q[1]);
            gz += g2 * (verts[o * 3 + 2] - q[2]);
        }
        gverts[o * 3] = gx;
        gverts[o * 3 + 1] = gy;
        gverts[o * 3 + 2] = gz;
    }
    closs = warp_sum(closs);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = closs;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        cpart[(size_t)b * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256)
fit_epilogue_kernel(FitDims d, psi_fit_config cfg, int np_sdf, int nchunk, int num_contact,
                    const float *__restrict__ x0, float *__restrict__ x, float *__restrict__ am,
                    float *__restrict__ av, int *__restrict__ step, const float *__restrict__ W1,
                    const float *__restrict__ W2, const float *__restrict__ W3,
                    const float *__restrict__ hand_l, const float *__restrict__ hand_r,
                    const float *__restrict__ h1pre, const float *__restrict__ h2pre,
                    const float *__restrict__ o6, const float *__restrict__ grot,
                    const float *__restrict__ gpose, const float *__restrict__ gshape,
                    const float *__restrict__ gtransl, const float *__restrict__ partial,
                    const float *__restrict__ cpart, float *__restrict__ losses) {
    __shared__ float sx[80], g[80], dso[128], dh2[512], dh1[512], pz[8][32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int H = d.hidden, Lz = d.latent, NO = d.nbody * 6;
    const int xdim = 9 + 10 + Lz + 2 * d.ncomp;
    const int zoff = 19, lhoff = 19 + Lz, rhoff = lhoff + d.ncomp;
    for (int e = tid; e < xdim; e += blockDim.x) { sx[e] = x[(size_t)b * xdim + e]; g[e] = 0.f; }
    __syncthreads();
    // Gram-Schmidt backward: joint 0 -> x[3:9], joints 1..nbody -> decoder outputs
    if (tid <= d.nbody) {
        float dx6[6];
        gs_bwd(tid == 0 ? sx + 3 : o6 + (size_t)b * NO + (tid - 1) * 6,
               grot + ((size_t)b * d.num_rot + tid) * 9, dx6);
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            if (tid == 0) g[3 + e] = dx6[e];
            else dso[(tid - 1) * 6 + e] = dx6[e];
        }
    }
    if (tid < 3) g[tid] = gtransl[(size_t)b * 3 + tid];
    if (tid < 10) g[9 + tid] = gshape[(size_t)b * d.NB + tid];
    // hand PCA backward
    if (tid < 2 * d.ncomp) {
        const int c = tid % d.ncomp;
        const bool right = tid >= d.ncomp;
        const float *comp = (right ? hand_r : hand_l) + c * 45;
        const float *gp = gpose + (size_t)b * d.J * 3 + (right ? (d.J - 15) * 3 : (d.J - 30) * 3);
        float a = 0.f;
        for (int k = 0; k < 45; ++k) a = fmaf(comp[k], gp[k], a);
        g[(right ? rhoff : lhoff) + c] = a;
    }
    __syncthreads();
    // MLP backward (W in the reference [out][in] layout -> coalesced over the input index)
    for (int i = tid; i < H; i += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < NO; ++k) a = fmaf(W3[(size_t)k * H + i], dso[k], a);
        dh2[i] = a * lrelu_grad(h2pre[(size_t)b * H + i]);
    }
    __syncthreads();
    for (int i = tid; i < H; i += blockDim.x) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int o = 0; o < H; o += 4) {
            a0 = fmaf(W2[(size_t)o * H + i], dh2[o], a0);
            a1 = fmaf(W2[(size_t)(o + 1) * H + i], dh2[o + 1], a1);
            a2 = fmaf(W2[(size_t)(o + 2) * H + i], dh2[o + 2], a2);
            a3 = fmaf(W2[(size_t)(o + 3) * H + i], dh2[o + 3], a3);
        }
        dh1[i] = ((a0 + a1) + (a2 + a3)) * lrelu_grad(h1pre[(size_t)b * H + i]);
    }
    __syncthreads();
    {   // dz[i] = sum_o W1[o][i] dh1[o] : 8 slices of the o range, reduced in fixed order
        const int i = tid & 31, sl = tid >> 5;
        float a = 0.f;
        if (i < Lz)
            for (int o = sl * (H / 8); o < (sl + 1) * (H / 8); ++o) a = fmaf(W1[(size_t)o * Lz + i], dh1[o], a);
        pz[sl][i] = a;
    }
    __syncthreads();
    if (tid < Lz) {
        float a = 0.f;
        for (int sl = 0; sl < 8; ++sl) a += pz[sl][tid];
        g[zoff + tid] = a + cfg.w_vposer * (2.0f * sx[zoff + tid] / (float)Lz);
    }
    __syncthreads();
    // + d L_rec, then Adam (torch.optim.Adam defaults; bias corrections in double)
    const int t = step[b] + 1;
    if (tid < xdim) {
        const float xe = sx[tid], diff = xe - x0[(size_t)b * xdim + tid];
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float ge = g[tid] + cfg.w_rec * sgn / (float)xdim;
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
        step[b] = t;
        float sn = 0.f, cn = 0.f, cs = 0.f;
        for (int i = 0; i < np_sdf; ++i) {
            sn += partial[((size_t)b * np_sdf + i) * 2];
            cn += partial[((size_t)b * np_sdf + i) * 2 + 1];
        }
        for (int i = 0; i < nchunk; ++i) cs += cpart[(size_t)b * nchunk + i];
        losses[(size_t)b * 4 + 2] = cfg.w_contact * (cs / (float)num_contact);
        losses[(size_t)b * 4 + 3] = cfg.w_collision * (cn > 0.f ? sn / cn : 0.f);
    }
}

__global__ void fit_reset_kernel(float *am, float *av, int *step, long n, int B) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { am[i] = 0.f; av[i] = 0.f; }
    if (i < B) step[i] = 0;
}

__global__ void fit_cam_kernel(const float *cam, long stride, int B, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 12) out[i] = cam[(size_t)(i / 12) * stride + (i % 12)];
}

static int enqueue_iteration(psi_fit_ctx *c, cudaStream_t st) {
    const FitDims d = {c->B, c->V, c->J, c->NB, c->latent, c->hidden, c->nbody, c->ncomp, c->num_rot};
    fit_prologue2_kernel<<<c->B, kMlpThreads, sizeof(MlpSmem), st>>>(d, c->x0, c->x, c->W1T, c->b1, c->W2T, c->b2, c->W3Tp, c->b3,
                                              c->hand_l, c->hand_r, c->pose_mean, c->cfg.w_rec,
                                              c->cfg.w_vposer, c->rot, c->pose, c->shape, c->transl,
                                              c->h1pre, c->h2pre, c->o6, c->losses);
    PSI_LAUNCHED();
    int rc = psi_lbs_fwd(c->model, c->B, c->shape, c->pose, c->transl, c->cam, 12, c->rot, c->num_rot,
                         c->verts, nullptr, c->saved, st);
    if (rc) return rc;
    // contact ids are ordered along a Morton curve of the template (fused.py): 32 consecutive
    // queries are neighbours on the body -> the group schedule walks the index once per warp
    rc = psi_nn_index_query_mode(c->index, c->verts, (long)c->V * 3, c->B, c->nu, c->csel, c->nnd, c->nni,
                                 c->nnhint, c->cfg.nn_mode > 0 ? c->cfg.nn_mode : 3, st);
    if (rc) return rc;
    rc = psi_sdf_fwd(c->sdf, 1, c->D, c->gmin, c->gmax, c->verts, c->B, c->V, nullptr, c->sdfv, c->sdfg,
                     c->partial, st);
    if (rc) return rc;
    {
        dim3 grid((unsigned)c->nchunk, (unsigned)c->B);
        fit_vertex_grad_kernel<<<grid, 256, 0, st>>>(c->V, c->nu, c->np_sdf, c->num_contact, c->verts,
                                                     c->scene_pts, c->sdfv, c->sdfg, c->partial, c->nnd,
                                                     c->nni, c->cslot, c->cweight, c->cfg.w_contact,
                                                     c->cfg.w_collision, c->cfg.robust_c, c->gverts, c->cpart);
        PSI_LAUNCHED();
    }
    rc = psi_lbs_bwd(c->model, c->B, c->shape, c->pose, c->cam, 12, c->saved, c->gverts, nullptr, c->gshape,
                     c->gpose, c->gtransl, c->grot, c->num_rot, c->lbs_ws, c->lbs_ws_bytes, st);
    if (rc) return rc;
    fit_epilogue2_kernel<<<c->B, kMlpThreads, sizeof(MlpSmem), st>>>(d, c->cfg, c->np_sdf, c->nchunk, c->num_contact, c->x0, c->x,
                                              c->am, c->av, c->step, c->W1, c->W2, c->W3, c->hand_l,
                                              c->hand_r, c->h1pre, c->h2pre, c->o6, c->grot, c->gpose,
                                              c->gshape, c->gtransl, c->partial, c->cpart, c->losses);
    PSI_LAUNCHED();
    return PSI_OK;
}

}  // namespace psi

extern "C" {

void psi_fit_destroy(psi_fit_ctx *c) {
    if (!c) return;
    if (c->exec) cudaGraphExecDestroy(c->exec);
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_out) cudaEventDestroy(c->ev_out);
    for (void *p : c->owned) cudaFree(p);
    delete c;
}

int psi_fit_create(psi_fit_ctx **out, const psi_lbs_model *model, int V, int J, int NB,
                   const psi_nn_index *index, const float *scene_points, const float *sdf, int D,
                   const float *h_grid_min, const float *h_grid_max, const float *h_W1,
                   const float *h_b1, const float *h_W2, const float *h_b2, const float *h_W3,
                   const float *h_b3, int latent, int hidden, int nbody, const float *h_hand_l,
                   const float *h_hand_r, const float *h_pose_mean, int ncomp,
                   const int *h_contact_ids, int num_contact, const psi_fit_config *cfg,
                   psi_stream_t stream) {
    using namespace psi;
    if (!out || !model || !index || !scene_points || !sdf || !h_grid_min || !h_grid_max || !h_W1 || !h_b1 ||
        !h_W2 || !h_b2 || !h_W3 || !h_b3 || !h_hand_l || !h_hand_r || !h_pose_mean || !h_contact_ids || !cfg)
        return PSI_ERR_BAD_ARG;
    if (cfg->B < 1 || num_contact < 1 || D < 1 || V < 1) return PSI_ERR_BAD_ARG;
    if (hidden != 512 || latent != 32 || nbody * 6 > 128 || nbody + 1 > J || ncomp > 16 ||
        J < 31 || NB < 10 || hidden % 32)
        return PSI_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    psi_fit_ctx *c = new (std::nothrow) psi_fit_ctx();
    if (!c) return PSI_ERR_ALLOC;
    c->model = model; c->index = index; c->cfg = *cfg;
    c->B = cfg->B; c->V = V; c->J = J; c->NB = NB; c->latent = latent; c->hidden = hidden; c->nbody = nbody;
    c->ncomp = ncomp; c->num_rot = nbody + 1; c->D = D; c->sdf = sdf; c->scene_pts = scene_points;
    c->num_contact = num_contact; c->exec = nullptr; c->pending_join = 0; c->gstream = nullptr; c->ev_in = c->ev_out = nullptr;
    if (cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming) != cudaSuccess) {
        psi_fit_destroy(c);
        return PSI_ERR_ALLOC;
    }
    for (int a = 0; a < 3; ++a) { c->gmin[a] = h_grid_min[a]; c->gmax[a] = h_grid_max[a]; }
    c->np_sdf = psi_sdf_num_partials(V);
    c->nchunk = (V + 255) / 256;
    int rc = PSI_OK;
    auto dev_alloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (rc != PSI_OK) return nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) { rc = PSI_ERR_ALLOC; return nullptr; }
        c->owned.push_back(p);
        return p;
    };
    auto upload_f = [&](const std::vector<float> &h) -> float * {
        float *p = (float *)dev_alloc(h.size() * sizeof(float));
        if (p && cudaMemcpyAsync(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PSI_ERR_ALLOC;
        return p;
    };
    auto upload_i = [&](const std::vector<int> &h) -> int * {
        int *p = (int *)dev_alloc(h.size() * sizeof(int));
        if (p && cudaMemcpyAsync(p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PSI_ERR_ALLOC;
        return p;
    };
    const int H = hidden, NO = nbody * 6;
    std::vector<float> W1(h_W1, h_W1 + (size_t)H * latent), W2(h_W2, h_W2 + (size_t)H * H), W3(h_W3, h_W3 + (size_t)NO * H);
    std::vector<float> W1T((size_t)latent * H), W2T((size_t)H * H), W3T((size_t)H * NO);
    for (int o = 0; o < H; ++o) for (int i = 0; i < latent; ++i) W1T[(size_t)i * H + o] = W1[(size_t)o * latent + i];
    for (int o = 0; o < H; ++o) for (int i = 0; i < H; ++i) W2T[(size_t)i * H + o] = W2[(size_t)o * H + i];
    for (int o = 0; o < NO; ++o) for (int i = 0; i < H; ++i) W3T[(size_t)i * NO + o] = W3[(size_t)o * H + i];
    // contact vertices: unique ids + multiplicity (duplicates across parts are kept by the reference, cvae.py:105-112)
    std::vector<int> cslot((size_t)V, -1), csel;
    std::vector<float> cweight((size_t)V, 0.f);
    for (int i = 0; i < num_contact; ++i) {
        const int v = h_contact_ids[i];
        if (v < 0 || v >= V) { psi_fit_destroy(c); return PSI_ERR_BAD_ARG; }
        if (cslot[v] < 0) { cslot[v] = (int)csel.size(); csel.push_back(v); }
        cweight[v] += 1.0f;
    }
    c->nu = (int)csel.size();
    std::vector<float> keep_alive[8] = {std::vector<float>(h_b1, h_b1 + H), std::vector<float>(h_b2, h_b2 + H),
                                        std::vector<float>(h_b3, h_b3 + NO),
                                        std::vector<float>(h_hand_l, h_hand_l + (size_t)ncomp * 45),
                                        std::vector<float>(h_hand_r, h_hand_r + (size_t)ncomp * 45),
                                        std::vector<float>(h_pose_mean, h_pose_mean + (size_t)J * 3)};
    c->W1 = upload_f(W1); c->W2 = upload_f(W2); c->W3 = upload_f(W3);
    c->W1T = upload_f(W1T); c->W2T = upload_f(W2T); c->W3T = upload_f(W3T);
    {   // output layer transposed and padded to 128 columns (16-byte aligned rows for the bulk copies)
        std::vector<float> W3Tp((size_t)H * 128, 0.f);
        for (int o = 0; o < NO; ++o) for (int i = 0; i < H; ++i) W3Tp[(size_t)i * 128 + o] = W3[(size_t)o * H + i];
        c->W3Tp = upload_f(W3Tp);
        if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;   // W3Tp dies here
    }
    if (cudaFuncSetAttribute(fit_prologue2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(fit_epilogue2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpSmem)) != cudaSuccess)
        rc = PSI_ERR_UNSUPPORTED;
    c->b1 = upload_f(keep_alive[0]); c->b2 = upload_f(keep_alive[1]); c->b3 = upload_f(keep_alive[2]);
    c->hand_l = upload_f(keep_alive[3]); c->hand_r = upload_f(keep_alive[4]); c->pose_mean = upload_f(keep_alive[5]);
    c->cweight = upload_f(cweight); c->csel = upload_i(csel); c->cslot = upload_i(cslot);
    const size_t B = (size_t)c->B, xd = 9 + 10 + (size_t)latent + 2 * (size_t)ncomp;
    auto fbuf = [&](size_t n) { return (float *)dev_alloc(n * sizeof(float)); };
    c->x0 = fbuf(B * xd); c->x = fbuf(B * xd); c->am = fbuf(B * xd); c->av = fbuf(B * xd);
    c->cam = fbuf(B * 12); c->rot = fbuf(B * c->num_rot * 9); c->pose = fbuf(B * J * 3); c->shape = fbuf(B * NB);
    c->transl = fbuf(B * 3); c->h1pre = fbuf(B * H); c->h2pre = fbuf(B * H); c->o6 = fbuf(B * NO);
    c->verts = fbuf(B * V * 3); c->saved = fbuf(psi_lbs_saved_floats(model, c->B) + 64);
    c->sdfv = fbuf(B * V); c->sdfg = fbuf(B * V * 3); c->partial = fbuf(B * c->np_sdf * 2);
    c->nnd = fbuf(B * c->nu); c->nni = (int *)dev_alloc(B * c->nu * sizeof(int));
    c->nnhint = (int *)dev_alloc(B * c->nu * sizeof(int)); c->gverts = fbuf(B * V * 3);
    c->cpart = fbuf(B * c->nchunk); c->gshape = fbuf(B * NB); c->gpose = fbuf(B * J * 3);
    c->grot = fbuf(B * c->num_rot * 9); c->gtransl = fbuf(B * 3); c->losses = fbuf(B * 4);
    c->step = (int *)dev_alloc(B * sizeof(int));
    c->lbs_ws_bytes = psi_lbs_bwd_workspace_bytes(model, c->B) + 64;
    c->lbs_ws = (float *)dev_alloc(c->lbs_ws_bytes);
    if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;
    if (rc != PSI_OK) { psi_fit_destroy(c); return rc; }
    *out = c;
    return PSI_OK;
}

int psi_fit_begin(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride, int num_iter,
                  psi_stream_t stream) {
    using namespace psi;
    if (!c || !xhr_init || !cam || num_iter < 0) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t xd = 9 + 10 + (size_t)c->latent + 2 * (size_t)c->ncomp, n = (size_t)c->B * xd;
    cudaError_t e = cudaMemcpyAsync(c->x0, xhr_init, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->x, xhr_init, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(c->nnhint, 0xff, (size_t)c->B * c->nu * sizeof(int), st);   // -1: no hint yet
    if (e != cudaSuccess) return (int)e;
    // one [3x4] transform per body (stride 0 broadcasts a shared one)
    fit_cam_kernel<<<(unsigned)((c->B * 12 + 255) / 256), 256, 0, st>>>(cam, cam_bstride, c->B, c->cam);
    PSI_LAUNCHED();
    fit_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->am, c->av, c->step, (long)n, c->B);
    PSI_LAUNCHED();
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (c->cfg.use_graph && cs == cudaStreamCaptureStatusNone && num_iter > 1) {
        // the loop runs on the context's own stream, ordered after / before the caller's stream
        cudaStream_t gs = c->gstream;
        e = cudaEventRecord(c->ev_in, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(gs, c->ev_in, 0);
        if (e != cudaSuccess) return (int)e;
        if (!c->exec) {
            // warm the lazily-initialised pieces outside the capture (state is reset below)
            int rc = enqueue_iteration(c, gs);
            if (rc) return rc;
            cudaGraph_t graph = nullptr;
            e = cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) return (int)e;
            rc = enqueue_iteration(c, gs);
            e = cudaStreamEndCapture(gs, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return (int)e;
            e = cudaGraphInstantiate(&c->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return (int)e;
            e = cudaMemcpyAsync(c->x, xhr_init, n * sizeof(float), cudaMemcpyDeviceToDevice, gs);
            if (e != cudaSuccess) return (int)e;
            fit_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, gs>>>(c->am, c->av, c->step, (long)n, c->B);
            PSI_LAUNCHED();
        }
        for (int it = 0; it < num_iter; ++it) {
            e = cudaGraphLaunch(c->exec, gs);
            if (e != cudaSuccess) return (int)e;
        }
        e = cudaEventRecord(c->ev_out, gs);
        if (e != cudaSuccess) return (int)e;
        c->pending_join = 1;
    } else {
        for (int it = 0; it < num_iter; ++it) {
            const int rc = enqueue_iteration(c, st);
            if (rc) return rc;
        }
        c->pending_join = 0;
    }
    return PSI_OK;
}

int psi_fit_end(psi_fit_ctx *c, float *xhr_out, float *losses_out, psi_stream_t stream) {
    if (!c || !xhr_out) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (c->pending_join) {
        e = cudaStreamWaitEvent(st, c->ev_out, 0);
        if (e != cudaSuccess) return (int)e;
        c->pending_join = 0;
    }
    const size_t n = (size_t)c->B * (9 + 10 + (size_t)c->latent + 2 * (size_t)c->ncomp);
    e = cudaMemcpyAsync(xhr_out, c->x, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && losses_out)
        e = cudaMemcpyAsync(losses_out, c->losses, (size_t)c->B * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? PSI_OK : (int)e;
}

int psi_fit_run(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride, int num_iter,
                float *xhr_out, float *losses_out, psi_stream_t stream) {
    if (!xhr_out) return PSI_ERR_BAD_ARG;
    const int rc = psi_fit_begin(c, xhr_init, cam, cam_bstride, num_iter, stream);
    if (rc) return rc;
    return psi_fit_end(c, xhr_out, losses_out, stream);
}

int psi_fit_launches_per_iteration(void) { return 13; }

}  // extern "C"
