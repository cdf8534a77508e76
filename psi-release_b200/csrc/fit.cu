// Fused scene-fitting loop: the whole iteration of FittingOP.cal_loss + backward + Adam
// (source/fitting_habitat.py:103-164,177-191) as 13 kernel launches, captured once in a CUDA graph.
//
//   fit_linear x2    VPoser decode, layers 2 and 3, over the whole batch (MLP 32->512->512->126, leaky 0.2,
//                    vposer_smpl.py:107-121) as tensor-core GEMMs; the 21 x 6D outputs land behind the
//                    root's 6D vector
//   lbs_fwd_impl     (lbs.cu) pose kernel: 6D -> R by Gram-Schmidt (cvae.py:46-55) for joints 0..21 --
//                    the reference's R -> axis-angle (torchgeometry) -> Rodrigues round trip is the
//                    identity on SO(3) up to rounding, and so is its Jacobian on the tangent directions
//                    that reach it; blend GEMM; skinning, whose epilogue samples the
//                    scene SDF at every fresh vertex (value + gradient + collision partial sums)
//   psi_nn_index_query_mode  contact NN distance (group schedule, hints from the last iteration)
//   lbs_bwd_impl     vertex kernel forms dL/dverts of the contact robustifier (fitting_habitat.py:141)
//                    and the collision mean (:155-160) in place; -> d betas, d hand axis-angles,
//                    d translation, d 6D (root) and the decoder-output gradient as a GEMM operand
//   fit_linear x2    decoder backward, layers 3 and 2
//   fit_step         per body: decoder layer 1 backward (d z = W1^T d h1), loss-term gradients, Adam step
//                    (torch.optim.Adam defaults, fitting_habitat.py:76), loss values, the inputs of the next
//                    iteration and its decoder layer 1 (h1 = lrelu(W1 z + b1)); W1 sits in shared memory
// Loss semantics: loss_mode 0 = the SUM over bodies of the reference's B=1 loss ('independent') -- every
// body is optimised exactly as the shipped batch_size-1 scripts do, bodies never interact, results are
// independent of B; loss_mode 1 = the reference's batch-coupled means (fitting_proxe.py:105,110,139,155-160
// with B > 1, demo.ipynb cell 16): the terms are means over the whole batch and the collision term divides by
// the batch-wide number of penetrating vertices (an integer counted with atomics, so still deterministic).
// All reductions have a fixed order: results are bit-reproducible.
//
// The loop itself is ONE graph launch: a conditional WHILE node whose body is the captured iteration; the
// last kernel of the body counts the remaining iterations down on the device (cudaGraphSetConditional).
// loop_mode 1 keeps the older form, one graph launch per iteration.
#include "common.cuh"
#include "rot6d.cuh"
#include "fit_fuse.cuh"
#include "fit_lbfgs.cuh"
#include <math.h>
#include <string.h>
#include <new>
#include <vector>

struct psi_fit_ctx {
    const psi_lbs_model *model;
    const psi_nn_index *index;
    psi_fit_config cfg;
    int B, V, J, NB, latent, hidden, nbody, ncomp, num_rot, nu, num_contact, D, np_sdf, nchunk;
    // constants: decoder weights as GEMM tiles (forward: W[out][in]; backward: W^T), biases, hands
    float *W1p, *Wf2, *Wf3, *Wb3, *Wb2, *b1, *b2, *b3, *hand_l, *hand_r, *pose_mean, *cweight;
    int no_pad;                    // decoder outputs (nbody*6) padded to the GEMM's N / K granularity
    int *csel, *cslot;
    const float *sdf, *scene_pts;
    float gmin[3], gmax[3];
    // state + scratch.  *A buffers are GEMM A operands ([body group][K/32][64][32], rows >= B zero)
    float *x0, *x, *am, *av, *cam, *rot6d, *pose, *shape, *transl, *h1pre, *h1A, *h2pre, *h2A,
        *g6_root, *g6A, *dh2A, *dh1, *verts, *saved, *sdfv, *sdfg, *partial, *nnd, *cpart,
        *gshape, *gpose, *gtransl, *lbs_ws, *losses, *gx, *xeval;
    int *nni, *step, *nnhint;
    psi::LbfgsState lb;            // optimizer 1: per-body L-BFGS state
    psi::LbfgsParams lbp;
    int *lb_info;                  // trace staging [B,4]
    int *neg_cnt;                  // loss_mode 1: this shard's count of penetrating vertices, [2] by iteration parity
    int *body_cnt;                 // loss_mode 0: per-body counts [2][B] by iteration parity
    // loss_mode 1 sharded over several contexts (psi_fit_set_peers): exchange buffer [2][16] of (tag << 32 | count)
    // written by the peers, the peers' buffers, the batch-wide count [2], a fit sequence number, an error flag
    unsigned long long *xbuf;
    unsigned long long *peer_buf[PSI_FIT_MAX_SHARDS];
    void *ipc_opened[PSI_FIT_MAX_SHARDS];
    int *neg_tot, *xseq, *xerr;
    int rank, world, batch_total;
    int *loop_left;                // iterations the WHILE node still has to run
    size_t lbs_ws_bytes;
    cudaGraphExec_t exec;          // one iteration (loop_mode 1)
    cudaGraphExec_t loop_exec;     // the whole loop (loop_mode 0)
    cudaGraphConditionalHandle loop_cond;
    int capturing_loop;            // enqueue_iteration is recording the WHILE body: fit_step gets the handle
    int unroll;                    // iterations per pass of the WHILE body
    int pending_join;
    int ev_out_recorded;           // ev_out marks the end of the last loop run on gstream
    cudaStream_t gstream;          // graphs cannot be captured on the legacy default stream
    cudaEvent_t ev_in, ev_out;
    std::vector<void *> owned;
};

namespace psi {

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.2f * v; }
__device__ __forceinline__ float lrelu_grad(float pre) { return pre > 0.f ? 1.0f : 0.2f; }

constexpr int kW1Row = 33;           // decoder layer 1 in shared memory: rows of 32 weights + 1 pad (conflict free both ways)
constexpr int kDefaultUnroll = 15;   // iterations per pass of the WHILE body when psi_fit_config.loop_unroll == 0

struct FitDims {
    int B, V, J, NB, latent, hidden, nbody, ncomp, num_rot;
};

}  // namespace psi
#include "fit_linear.cuh"
namespace psi {

// x layout [75]: t3 | 6D | betas10 | z32 | lh12 | rh12   (cvae.py:28-33 after convert_to_6D_rot)
//
// One per-body kernel closes an iteration and opens the next:
//   do_post: assemble dL/dx (translation, 6D global orientation, betas, latent through the decoder's
//            backward GEMMs, hand PCA), add d L_rec and d L_vposer, Adam step (torch.optim.Adam
//            defaults, fitting_habitat.py:76; bias corrections in double), the four loss values of
//            the iteration just evaluated;
//   then, for the (updated) x: the latent z as the decoder's first A operand, the root's 6D vector,
//            betas, translation, hand PCA + pose_mean (smplx) for the LBS kernels.
// psi_fit_begin launches it once with do_post = 0.
template <int OPT>      // 0 = Adam, 1 = per-body L-BFGS (fit_lbfgs.cuh)
__global__ void __launch_bounds__(128)
fit_step_kernel(FitDims d, psi_fit_config cfg, int do_post, int np_sdf, int nchunk, int num_contact,
                const float *__restrict__ x0, float *__restrict__ x, float *__restrict__ am,
                float *__restrict__ av, int *__restrict__ step, const float *__restrict__ hand_l,
                const float *__restrict__ hand_r, const float *__restrict__ pose_mean,
                const float *__restrict__ g6_root,
                const float *__restrict__ gpose, const float *__restrict__ gshape,
                const float *__restrict__ gtransl, const float *__restrict__ partial,
                const float *__restrict__ cpart, float *__restrict__ losses,
                float *__restrict__ rot6d, float *__restrict__ pose, float *__restrict__ shape,
                float *__restrict__ transl, float *__restrict__ gx_out, float *__restrict__ xeval_out,
                int *__restrict__ neg_cnt, const int *__restrict__ neg_tot, int batch_total, int *__restrict__ body_cnt,
                int *__restrict__ loop_left,
                cudaGraphConditionalHandle loop_cond, const LbfgsParams lbp, const LbfgsState lbs,
                const float *__restrict__ W1p, const float *__restrict__ b1, const float *__restrict__ dh1,
                float *__restrict__ h1pre, float *__restrict__ h1A) {
    extern __shared__ __align__(128) float sW[];      // decoder layer 1, W1[n][i] in rows of kW1Row floats
    __shared__ __align__(8) uint64_t wbar;
    __shared__ float sx[96], g[96], sx0[96];
    __shared__ float s_dh[512], s_dzp[4][32];
    __shared__ float s_loss, s_bc[2];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int Lz = d.latent, H = d.hidden;
    if (tid == 0) {
        mbar_init(&wbar, 1);
        mbar_fence_init();
        // a constant: requested before the dependency wait, lands while the state below is read
        mbar_arrive_expect_tx(&wbar, (uint32_t)(H * kW1Row * 4));
        tma_load_1d(sW, W1p, (uint32_t)(H * kW1Row * 4), &wbar);
    }
    float b1r[4];              // layer 1's bias for this thread's four hidden units (H = 512): a constant too
#pragma unroll
    for (int i = 0; i < 4; ++i) b1r[i] = b1[tid + i * 128];
    const int xdim = 9 + 10 + Lz + 2 * d.ncomp;   // 75
    const int zoff = 19, lhoff = 19 + Lz, rhoff = lhoff + d.ncomp;
    // ... and so are the hand PCA rows of the pose entry this thread writes at the end (entries tid and tid + 128 of
    // [J*3]: at most one of the two is a hand entry)
    const int hl0 = (d.J - 30) * 3, hr0 = (d.J - 15) * 3;
    const int eh = tid >= hl0 ? tid : tid + 128;            // this thread's hand entry, if eh < J*3
    float hc[16], pm2[2];
#pragma unroll
    for (int c = 0; c < 16; ++c)
        hc[c] = (eh < d.J * 3 && c < d.ncomp) ? (eh >= hr0 ? hand_r[c * 45 + (eh - hr0)] : hand_l[c * 45 + (eh - hl0)]) : 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) pm2[i] = tid + i * 128 < d.J * 3 ? pose_mean[tid + i * 128] : 0.f;
    pdl_wait();
    for (int e = tid; e < xdim; e += blockDim.x) sx[e] = x[(size_t)b * xdim + e];
    // Everything the post-processing half reads from global memory is requested HERE, in one round trip: the Adam
    // state of this thread's component, its raw gradient component (hand PCA backward included), the loss partial sums
    // (warp 3) and d h1.  (Loads left inside run-time loops further down cost a dependent round trip each.)
    float x0e = 0.f, ame = 0.f, ave = 0.f, gin = 0.f;
    float pf[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int t_prev = 0;
    if (do_post) {
        t_prev = step[b];
        if (OPT == 0 && tid == 96) {
            // Adam's bias corrections (double, as torch computes them on the host): one thread, while the loads fly
            const double bc1 = 1.0 - pow((double)cfg.beta1, (double)(t_prev + 1));
            const double bc2 = 1.0 - pow((double)cfg.beta2, (double)(t_prev + 1));
            s_bc[0] = (float)((double)cfg.lr / bc1);
            s_bc[1] = (float)sqrt(bc2);
        }
        if (tid < xdim) { x0e = x0[(size_t)b * xdim + tid]; ame = am[(size_t)b * xdim + tid]; ave = av[(size_t)b * xdim + tid]; }
        float t4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t4[i] = dh1[(size_t)b * H + tid + i * 128];
        if (tid < 3) gin = gtransl[(size_t)b * 3 + tid];
        else if (tid < 9) gin = g6_root[(size_t)b * 6 + tid - 3];
        else if (tid < 19) gin = gshape[(size_t)b * d.NB + tid - 9];
        else if (tid >= lhoff && tid < xdim) {       // hand PCA backward
            const int u = tid - lhoff, c = u % d.ncomp;
            const bool right = u >= d.ncomp;
            const float *comp = (right ? hand_r : hand_l) + c * 45;
            const float *gp = gpose + (size_t)b * d.J * 3 + (right ? (d.J - 15) * 3 : (d.J - 30) * 3);
            float a = 0.f;
            for (int k = 0; k < 45; ++k) a = fmaf(comp[k], gp[k], a);
            gin = a;
        }
        if (tid >= 96) {
            const int ln = tid - 96;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = ln + q * 32;
                if (i < np_sdf) {
                    pf[q * 2] = partial[((size_t)b * np_sdf + i) * 2];
                    pf[q * 2 + 1] = partial[((size_t)b * np_sdf + i) * 2 + 1];
                }
                if (i < nchunk) pf[4 + q] = cpart[(size_t)b * nchunk + i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s_dh[tid + i * 128] = t4[i];
        if (tid < xdim) sx0[tid] = x0e;
    }
    __syncthreads();
    mbar_wait(&wbar, 0);
    if (do_post) {
        // d z = W1^T d h1 (the decoder's last backward layer, 32 x 512 per body): lane = latent component, warp = a
        // quarter of the hidden units, four interleaved chains each; fixed order
        const int i = tid & 31, part = tid >> 5, k0 = part * (H / 4);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int k = k0; k < k0 + H / 4; k += 4) {
            a0 = fmaf(s_dh[k], sW[k * kW1Row + i], a0);
            a1 = fmaf(s_dh[k + 1], sW[(k + 1) * kW1Row + i], a1);
            a2 = fmaf(s_dh[k + 2], sW[(k + 2) * kW1Row + i], a2);
            a3 = fmaf(s_dh[k + 3], sW[(k + 3) * kW1Row + i], a3);
        }
        s_dzp[part][i] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    // loss_mode 1: the L1 / L2 terms are means over the whole batch (the contact and collision terms are scaled
    // where their gradients are formed, lbs_vertex_bwd<FIT>)
    const float bdiv = cfg.loss_mode == 1 ? (float)(neg_tot ? batch_total : d.B) : 1.0f;   // sharded: bodies of the WHOLE batch
    if (do_post) {
        if (tid >= zoff && tid < zoff + Lz) {
            const int i = tid - zoff;
            const float dzi = (s_dzp[0][i] + s_dzp[1][i]) + (s_dzp[2][i] + s_dzp[3][i]);
            g[tid] = dzi + cfg.w_vposer * (2.0f * sx[tid] / ((float)Lz * bdiv));
        } else if (tid < xdim) {
            g[tid] = gin;
        }
        if (tid >= 96) {        // warp 3: loss values of the iteration just evaluated (x before the update)
            const int ln = tid - 96;
            float r = 0.f, zz = 0.f, sn = 0.f, cn = 0.f, cs = 0.f;
            for (int e = ln; e < xdim; e += 32) r += fabsf(sx0[e] - sx[e]);
            for (int i = ln; i < Lz; i += 32) zz = fmaf(sx[zoff + i], sx[zoff + i], zz);
            sn += pf[0]; cn += pf[1];               // rows ln, ln + 32 (requested at the top), then the rest in order
            sn += pf[2]; cn += pf[3];
            for (int i = ln + 64; i < np_sdf; i += 32) {
                sn += partial[((size_t)b * np_sdf + i) * 2];
                cn += partial[((size_t)b * np_sdf + i) * 2 + 1];
            }
            cs += pf[4];
            cs += pf[5];
            for (int i = ln + 64; i < nchunk; i += 32) cs += cpart[(size_t)b * nchunk + i];
            r = warp_sum(r); zz = warp_sum(zz); sn = warp_sum(sn); cn = warp_sum(cn); cs = warp_sum(cs);
            // loss_mode 1: this body's share of the batch means (the rows add up to the reference's scalars)
            if (cfg.loss_mode == 1) cn = (float)(neg_tot ? neg_tot : neg_cnt)[t_prev & 1];
            if (ln == 0) {
            losses[(size_t)b * 4 + 0] = cfg.w_rec * (r / ((float)xdim * bdiv));
            losses[(size_t)b * 4 + 1] = cfg.w_vposer * (zz / ((float)Lz * bdiv));
            losses[(size_t)b * 4 + 2] = cfg.w_contact * (cs / ((float)num_contact * bdiv));
            losses[(size_t)b * 4 + 3] = cfg.w_collision * (cn > 0.f ? sn / cn : 0.f);
            s_loss = ((cfg.w_rec * (r / ((float)xdim * bdiv)) + cfg.w_vposer * (zz / ((float)Lz * bdiv))) +
                      cfg.w_contact * (cs / ((float)num_contact * bdiv))) + cfg.w_collision * (cn > 0.f ? sn / cn : 0.f);
            }
        }
        __syncthreads();
        const int t = t_prev + 1;
        float xn = 0.f;
        float ge = 0.f;
        if (tid < xdim) {
            const float xe = sx[tid], diff = xe - x0e;
            const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            ge = g[tid] + cfg.w_rec * sgn / ((float)xdim * bdiv);
            const size_t o = (size_t)b * xdim + tid;
            gx_out[o] = ge;          // trace: dL/dx of the iteration just evaluated, and where it was evaluated
            xeval_out[o] = xe;
        }
        if (OPT == 1) {
            // L-BFGS: one warp advances this body's line search / direction by one evaluation
            __syncthreads();                       // every thread has read g[tid]
            if (tid < xdim) g[tid] = ge;
            __syncthreads();
            if (tid < 32) lbfgs_feed_body(lbp, lbs, b, xdim, (double)s_loss, g, sx, x, tid);
        } else if (tid < xdim) {
            const float xe = sx[tid];
            const size_t o = (size_t)b * xdim + tid;
            const float m = cfg.beta1 * ame + (1.0f - cfg.beta1) * ge;
            const float v = cfg.beta2 * ave + (1.0f - cfg.beta2) * ge * ge;
            am[o] = m;
            av[o] = v;
            const float step_size = s_bc[0];
            const float denom = sqrtf(v) / s_bc[1] + cfg.eps;
            xn = xe - (m / denom) * step_size;
            x[o] = xn;
        }
        __syncthreads();           // everyone has read sx / step[b]
        if (OPT == 0 && tid < xdim) sx[tid] = xn;
        if (tid == 0) {
            step[b] = t;
            if (body_cnt) body_cnt[(size_t)(t & 1) * d.B + b] = 0;     // the next iteration's counter of this body
        }
        if (b == 0 && tid == 0) {
            // the next iteration's penetration counter (its parity is t & 1; nobody reads it during this kernel)
            if (neg_cnt) neg_cnt[t & 1] = 0;
            if (loop_left) {           // body of the WHILE node: one iteration done
                const int left = *loop_left - 1;
                *loop_left = left;
                cudaGraphSetConditional(loop_cond, left > 0 ? 1u : 0u);
            }
        }
        __syncthreads();
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
    // inputs of the next evaluation
    // decoder layer 1 for the new latent: h1 = lrelu(W1 z + b1), as the pre-activation (kept for the backward) and as
    // the A operand of layer 2's GEMM
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int n = tid + q * 128;
        float v0 = b1r[q], v1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            v0 = fmaf(sW[n * kW1Row + i], sx[zoff + i], v0);
            v1 = fmaf(sW[n * kW1Row + i + 1], sx[zoff + i + 1], v1);
        }
        const float v = v0 + v1;
        h1pre[(size_t)b * H + n] = v;
        h1A[a_index(b, n, H)] = lrelu(v);
    }
    if (tid < 6) rot6d[(size_t)b * d.num_rot * 6 + tid] = sx[3 + tid];
    // axis-angle pose vector [J*3]: only joints >= num_rot are read by the LBS kernels
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int e = tid + i * 128;
        if (e >= d.J * 3) break;
        float v = pm2[i];
        if (e >= hl0) {        // e == eh
            const int off = e >= hr0 ? rhoff : lhoff;
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < d.ncomp) v = fmaf(sx[off + c], hc[c], v);
        } else if (e < d.num_rot * 3) {
            v = 0.f;
        }
        pose[(size_t)b * d.J * 3 + e] = v;
    }
    for (int e = tid; e < d.NB; e += blockDim.x) shape[(size_t)b * d.NB + e] = e < 10 ? sx[9 + e] : 0.f;
    if (tid < 3) transl[(size_t)b * 3 + tid] = sx[tid];
}

__global__ void fit_reset_kernel(float *am, float *av, int *step, long n, int B, int *neg_cnt, int *body_cnt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { am[i] = 0.f; av[i] = 0.f; }
    if (i < B) step[i] = 0;
    if (i < 2) neg_cnt[i] = 0;
    if (i < 2 * B) body_cnt[i] = 0;
}

__global__ void fit_lbfgs_reset_kernel(LbfgsScalars *sc, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        LbfgsScalars z = {};
        z.phase = LB_START;
        z.H = 1.0;
        sc[i] = z;
    }
}
__global__ void fit_lbfgs_info_kernel(const LbfgsScalars *sc, int B, int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        out[i * 4 + 0] = sc[i].phase; out[i * 4 + 1] = sc[i].n_iter;
        out[i * 4 + 2] = sc[i].evals; out[i * 4 + 3] = sc[i].hist_count;
    }
}

// ---- batch-coupled loss sharded over several contexts: 8 bytes per peer per iteration through peer memory --------
struct PeerBufs {
    unsigned long long *p[PSI_FIT_MAX_SHARDS];
};
// after the SDF pass: this shard's penetration count of iteration `it` into every peer's slot [it & 1][rank]
__global__ void fit_publish_kernel(PeerBufs peers, int rank, int world, const int *__restrict__ neg_cnt,
                                   const int *__restrict__ step, const int *__restrict__ xseq) {
    pdl_wait();
    const int q = threadIdx.x;
    if (q >= world || q == rank) return;
    const int it = step[0], par = it & 1;
    const unsigned long long tag = ((unsigned long long)(unsigned)xseq[0] << 16) | (unsigned long long)((it + 1) & 0xffff);
    const unsigned long long val = (tag << 32) | (unsigned long long)(unsigned)neg_cnt[par];
    unsigned long long *dst = peers.p[q] + par * PSI_FIT_MAX_SHARDS + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(val) : "memory");
}
// before the vertex backward kernel: wait for every peer's count of this iteration, add them up
__global__ void fit_collect_kernel(const unsigned long long *__restrict__ xbuf, int rank, int world,
                                   const int *__restrict__ neg_cnt, int *__restrict__ neg_tot,
                                   const int *__restrict__ step, const int *__restrict__ xseq, int *__restrict__ xerr) {
    pdl_wait();
    const int q = threadIdx.x;
    const int it = step[0], par = it & 1;
    const unsigned long long tag = ((unsigned long long)(unsigned)xseq[0] << 16) | (unsigned long long)((it + 1) & 0xffff);
    int mine = 0;
    if (q < world && q != rank) {
        const unsigned long long *src = xbuf + par * PSI_FIT_MAX_SHARDS + q;
        unsigned long long v = 0;
        long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if ((v >> 32) == tag) break;
            if (++spins > (1L << 24)) {      // ~ a second: a peer is gone or out of step -- flag it, fall back to what we have
                atomicExch(xerr, 1);
                v = 0;
                break;
            }
            __nanosleep(64);
        }
        mine = (int)(unsigned)(v & 0xffffffffull);
    } else if (q == rank) {
        mine = neg_cnt[par];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);     // integers: any order
    if (q == 0) neg_tot[par] = mine;
}
__global__ void fit_seq_kernel(int *xseq) { xseq[0] = (xseq[0] + 1) & 0xffff; }

// head of the loop graph: arm the WHILE node with the iteration count psi_fit_begin stored
__global__ void fit_loop_head_kernel(cudaGraphConditionalHandle cond, const int *loop_left) {
    cudaGraphSetConditional(cond, *loop_left > 0 ? 1u : 0u);
}
__global__ void fit_loop_set_kernel(int *loop_left, int n) { *loop_left = n; }

__global__ void fit_cam_kernel(const float *cam, long stride, int B, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 12) out[i] = cam[(size_t)(i / 12) * stride + (i % 12)];
}

static int launch_linear(cudaStream_t st, int B, const float *A, const float *W, int K, int N, const float *bias,
                         const float *gate, float *pre_out, float *out_rm, int ld, int n_valid, float *outA,
                         int outA_kpad, int act) {
    LinearParams p;
    p.A = A; p.W = W; p.bias = bias; p.gate = gate; p.pre_out = pre_out; p.out_rm = out_rm; p.outA = outA;
    p.K = K; p.N = N; p.B = B; p.n_valid = n_valid; p.ld = ld; p.outA_kpad = outA_kpad; p.act = act;
    dim3 grid((unsigned)(N / kLT), (unsigned)(2 * ((B + kBG - 1) / kBG)));
    if (K % (2 * kKC) || K / kKC > kLMaxChunks || N % kLT) return PSI_ERR_UNSUPPORTED;
    (psi::skip_kernel("fit_linear") ? cudaSuccess : launch_pdl(fit_linear_kernel, dim3(grid), dim3(128), (size_t)(K / kKC) * kLChunkBytes, st, p));
    PSI_LAUNCHED_K("fit_linear");
    return PSI_OK;
}

static int launch_step(psi_fit_ctx *c, int do_post, cudaStream_t st) {
    const FitDims d = {c->B, c->V, c->J, c->NB, c->latent, c->hidden, c->nbody, c->ncomp, c->num_rot};
    auto go = [&](auto kernel) {
        return launch_pdl(kernel, dim3(c->B), dim3(128), (size_t)c->hidden * kW1Row * 4, st, d, c->cfg, do_post, c->np_sdf,
                          c->nchunk, c->num_contact,
                          c->x0, c->x, c->am, c->av, c->step, c->hand_l, c->hand_r, c->pose_mean, c->g6_root, c->gpose,
                          c->gshape, c->gtransl, c->partial, c->cpart, c->losses, c->rot6d, c->pose, c->shape, c->transl,
                          c->gx, c->xeval, c->cfg.loss_mode == 1 ? c->neg_cnt : nullptr,
                          c->cfg.loss_mode == 1 && c->world > 1 ? c->neg_tot : nullptr, c->batch_total,
                          c->cfg.loss_mode == 0 ? c->body_cnt : nullptr,
                          c->capturing_loop ? c->loop_left : nullptr, c->loop_cond, c->lbp, c->lb, c->W1p, c->b1, c->dh1,
                          c->h1pre, c->h1A);
    };
    if (!psi::skip_kernel("fit_step")) {
        if (c->cfg.optimizer == 1) go(fit_step_kernel<1>);
        else go(fit_step_kernel<0>);
    }
    PSI_LAUNCHED_K("fit_step");
    return PSI_OK;
}

// one iteration = 13 launches; expects the per-body inputs of fit_step_kernel's second half
static int enqueue_iteration(psi_fit_ctx *c, cudaStream_t st) {
    const int H = c->hidden, NO = c->nbody * 6, NOp = c->no_pad, B = c->B;
    // VPoser decode: 32 -> 512 -> 512 -> nbody*6 (the 6D vectors land behind the root's in rot6d)
    // (layer 1, 32 -> 512, is the tail of fit_step: per body, W1 in shared memory)
    int rc = launch_linear(st, B, c->h1A, c->Wf2, H, H, c->b2, nullptr, c->h2pre, nullptr, 0, H, c->h2A, H, 1);
    if (rc) return rc;
    rc = launch_linear(st, B, c->h2A, c->Wf3, H, NOp, c->b3, nullptr, nullptr, c->rot6d + 6, c->num_rot * 6, NO,
                       nullptr, 0, 0);
    if (rc) return rc;
    // LBS forward; the skinning epilogue also samples the scene SDF at every fresh vertex
    SdfFuse sf;
    sf.g.grid = c->sdf;
    sf.g.D = c->D;
    for (int a = 0; a < 3; ++a) {
        const float ext = c->gmax[a] - c->gmin[a];
        sf.g.gmin[a] = c->gmin[a];
        sf.g.inv_extent[a] = 1.0f / ext;
        sf.g.gscale[a] = (float)(c->D - 1) / 2.0f * (2.0f / ext);
    }
    sf.sdfv = c->sdfv; sf.sdfg = c->sdfg; sf.partial = c->partial;
    sf.neg_cnt = c->cfg.loss_mode == 1 ? c->neg_cnt : nullptr; sf.step = c->step;
    sf.body_cnt = c->cfg.loss_mode == 0 ? c->body_cnt : nullptr; sf.nbodies = B;
    rc = lbs_fwd_impl(c->model, B, c->shape, c->pose, c->transl, c->cam, 12, nullptr, c->rot6d, c->num_rot,
                      c->verts, nullptr, c->saved, &sf, st);
    if (rc) return rc;
    const bool sharded = c->cfg.loss_mode == 1 && c->world > 1;
    if (sharded) {       // this shard's penetration count -> the peers (8 bytes each, plain stores through the peer mapping)
        PeerBufs pb;
        for (int q = 0; q < PSI_FIT_MAX_SHARDS; ++q) pb.p[q] = c->peer_buf[q];
        launch_pdl(fit_publish_kernel, dim3(1), dim3(32), 0, st, pb, c->rank, c->world, c->neg_cnt, c->step, c->xseq);
        PSI_LAUNCHED_K("fit_publish");
    }
    // contact ids are ordered by dominant joint + kd cells of the template (fused.py): 32 consecutive
    // queries are neighbours on the posed body -> the group schedule walks the index once per warp
    rc = psi_nn_index_query_mode(c->index, c->verts, (long)c->V * 3, B, c->nu, c->csel, c->nnd, c->nni,
                                 c->nnhint, c->cfg.nn_mode > 0 ? c->cfg.nn_mode : 3, st);
    if (rc) return rc;
    if (sharded) {       // ... and theirs, by now long arrived (the NN walk ran in between)
        launch_pdl(fit_collect_kernel, dim3(1), dim3(32), 0, st, c->xbuf, c->rank, c->world, c->neg_cnt, c->neg_tot, c->step,
                   c->xseq, c->xerr);
        PSI_LAUNCHED_K("fit_collect");
    }
    // LBS backward; dL/dverts of the contact + collision terms is formed at the head of its vertex kernel
    VGradFuse vg;
    vg.verts = c->verts; vg.scene = c->scene_pts; vg.sdfv = c->sdfv; vg.sdfg = c->sdfg; vg.partial = c->partial;
    vg.nnd = c->nnd; vg.nni = c->nni; vg.cslot = c->cslot; vg.cweight = c->cweight;
    vg.w_contact = c->cfg.w_contact; vg.w_coll = c->cfg.w_collision; vg.robust_c = c->cfg.robust_c;
    vg.nu = c->nu; vg.np_sdf = c->np_sdf; vg.num_contact = c->num_contact; vg.cpart = c->cpart;
    vg.neg_cnt = c->cfg.loss_mode == 1 ? (sharded ? c->neg_tot : c->neg_cnt) : nullptr; vg.step = c->step;
    vg.body_cnt = c->cfg.loss_mode == 0 ? c->body_cnt : nullptr;
    vg.bdiv = c->cfg.loss_mode == 1 ? (float)(sharded ? c->batch_total : B) : 1.0f;
    rc = lbs_bwd_impl(c->model, B, c->pose, c->cam, 12, c->saved, nullptr, &vg, nullptr, c->gshape, c->gpose,
                      c->gtransl, nullptr, c->num_rot, c->rot6d, c->g6_root, c->g6A, NOp, c->lbs_ws,
                      c->lbs_ws_bytes, st);
    if (rc) return rc;
    // decoder backward: d h2 = (d o . W3) * lrelu', d h1 = (d h2 . W2) * lrelu', d z = d h1 . W1
    rc = launch_linear(st, B, c->g6A, c->Wb3, NOp, H, nullptr, c->h2pre, nullptr, nullptr, 0, H, c->dh2A, H, 0);
    if (rc) return rc;
    rc = launch_linear(st, B, c->dh2A, c->Wb2, H, H, nullptr, c->h1pre, nullptr, c->dh1, H, H, nullptr, 0, 0);
    if (rc) return rc;
    return launch_step(c, 1, st);       // d z, Adam, losses, and the next iteration's per-body inputs + decoder layer 1
}

}  // namespace psi

extern "C" {

void psi_fit_destroy(psi_fit_ctx *c) {
    if (!c) return;
    if (c->exec) cudaGraphExecDestroy(c->exec);
    if (c->loop_exec) cudaGraphExecDestroy(c->loop_exec);
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_out) cudaEventDestroy(c->ev_out);
    for (int q = 0; q < PSI_FIT_MAX_SHARDS; ++q)
        if (c->ipc_opened[q]) cudaIpcCloseMemHandle(c->ipc_opened[q]);
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
    psi::Range nvtx_range("psi_fit_create");
    using namespace psi;
    if (!out || !model || !index || !scene_points || !sdf || !h_grid_min || !h_grid_max || !h_W1 || !h_b1 ||
        !h_W2 || !h_b2 || !h_W3 || !h_b3 || !h_hand_l || !h_hand_r || !h_pose_mean || !h_contact_ids || !cfg)
        return PSI_ERR_BAD_ARG;
    if (cfg->B < 1 || num_contact < 1 || D < 1 || V < 1) return PSI_ERR_BAD_ARG;
    if (cfg->loss_mode < 0 || cfg->loss_mode > 1 || cfg->loop_mode < 0 || cfg->loop_mode > 1 || cfg->loop_unroll < 0 ||
        cfg->loop_unroll > 64 || cfg->optimizer < 0 || cfg->optimizer > 1 || cfg->lbfgs_history < 0 || cfg->lbfgs_history > 1024)
        return PSI_ERR_BAD_ARG;
    if (cfg->optimizer == 1 && cfg->loss_mode != 0) return PSI_ERR_UNSUPPORTED;   // per-body line searches need a per-body loss
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
    c->loop_exec = nullptr; c->capturing_loop = 0; c->loop_cond = 0; c->ev_out_recorded = 0;
    c->unroll = cfg->loop_unroll > 0 ? cfg->loop_unroll : kDefaultUnroll;
    if (cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming) != cudaSuccess) {
        psi_fit_destroy(c);
        return PSI_ERR_ALLOC;
    }
    for (int a = 0; a < 3; ++a) { c->gmin[a] = h_grid_min[a]; c->gmax[a] = h_grid_max[a]; }
    if (ensure_max_dyn_smem(fit_linear_kernel, (size_t)kLMaxChunks * kLChunkBytes) != PSI_OK ||
        ensure_max_dyn_smem(fit_step_kernel<0>, (size_t)hidden * kW1Row * 4) != PSI_OK ||
        ensure_max_dyn_smem(fit_step_kernel<1>, (size_t)hidden * kW1Row * 4) != PSI_OK) {
        psi_fit_destroy(c);
        return PSI_ERR_UNSUPPORTED;
    }
    c->np_sdf = (V + 255) / 256;               // rows of the skinning kernel's SDF partial sums
    c->nchunk = lbs_vertex_chunks(model);      // rows of the vertex backward kernel's contact partial sums
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
    {   // decoder weights as GEMM tiles: forward layers use W[out][in] as given, backward layers W^T
        auto transpose = [](const float *M, int R, int C) {
            std::vector<float> T((size_t)R * C);
            for (int r = 0; r < R; ++r) for (int q = 0; q < C; ++q) T[(size_t)q * R + r] = M[(size_t)r * C + q];
            return T;
        };
        int np = 0, kp = 0;
        std::vector<float> w1p((size_t)H * kW1Row, 0.f);            // layer 1 as fit_step stages it: W1[n][i], padded rows
        for (int n = 0; n < H; ++n)
            for (int i = 0; i < latent; ++i) w1p[(size_t)n * kW1Row + i] = h_W1[(size_t)n * latent + i];
        c->W1p = upload_f(w1p);
        std::vector<float> t2 = linear_weight_tiles(h_W2, H, H, &np, &kp);
        c->Wf2 = upload_f(t2);
        std::vector<float> t3 = linear_weight_tiles(h_W3, NO, H, &np, &kp);
        c->Wf3 = upload_f(t3);
        c->no_pad = (np + kKC - 1) / kKC * kKC;                     // N of layer 3 = K of its backward: 126 -> 128
        if (c->no_pad != np) rc = PSI_ERR_UNSUPPORTED;              // (nbody*6 rounded to 16 must also be a multiple of 32)
        const std::vector<float> tr3 = transpose(h_W3, NO, H);      // [H][NO]
        std::vector<float> tb3 = linear_weight_tiles(tr3.data(), H, NO, &np, &kp);
        c->Wb3 = upload_f(tb3);
        const std::vector<float> tr2 = transpose(h_W2, H, H);
        std::vector<float> tb2 = linear_weight_tiles(tr2.data(), H, H, &np, &kp);
        c->Wb2 = upload_f(tb2);
        std::vector<float> v1(h_b1, h_b1 + H), v2(h_b2, h_b2 + H), v3(h_b3, h_b3 + NO),
            hl(h_hand_l, h_hand_l + (size_t)ncomp * 45), hr(h_hand_r, h_hand_r + (size_t)ncomp * 45),
            pm(h_pose_mean, h_pose_mean + (size_t)J * 3);
        c->b1 = upload_f(v1); c->b2 = upload_f(v2); c->b3 = upload_f(v3);
        c->hand_l = upload_f(hl); c->hand_r = upload_f(hr); c->pose_mean = upload_f(pm);
        c->cweight = upload_f(cweight); c->csel = upload_i(csel); c->cslot = upload_i(cslot);
        if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;   // host vectors die here
    }
    const size_t B = (size_t)c->B, xd = 9 + 10 + (size_t)latent + 2 * (size_t)ncomp;
    auto fbuf = [&](size_t n) { return (float *)dev_alloc(n * sizeof(float)); };
    // GEMM A operands: zero-initialised once (rows of bodies >= B and padded k stay zero)
    const size_t Bp = (B + kBG - 1) / kBG * kBG;
    auto zbuf = [&](size_t n) {
        float *p = fbuf(n);
        if (p && cudaMemsetAsync(p, 0, n * sizeof(float), st) != cudaSuccess) rc = PSI_ERR_ALLOC;
        return p;
    };
    c->x0 = fbuf(B * xd); c->x = fbuf(B * xd); c->am = fbuf(B * xd); c->av = fbuf(B * xd);
    c->cam = fbuf(B * 12); c->rot6d = fbuf(B * c->num_rot * 6); c->pose = fbuf(B * J * 3); c->shape = fbuf(B * NB);
    c->transl = fbuf(B * 3); c->h1pre = fbuf(B * H); c->h2pre = fbuf(B * H);
    c->h1A = zbuf(Bp * H); c->h2A = zbuf(Bp * H);
    c->g6A = zbuf(Bp * c->no_pad); c->dh2A = zbuf(Bp * H); c->dh1 = fbuf(B * H);
    c->g6_root = fbuf(B * 6);
    c->verts = fbuf(B * V * 3); c->saved = fbuf(psi_lbs_saved_floats(model, c->B) + 64);
    c->sdfv = fbuf(B * V); c->sdfg = fbuf(B * V * 3); c->partial = fbuf(B * c->np_sdf * 2);
    c->nnd = fbuf(B * c->nu); c->nni = (int *)dev_alloc(B * c->nu * sizeof(int));
    c->nnhint = (int *)dev_alloc(B * c->nu * sizeof(int));
    c->cpart = fbuf(B * c->nchunk); c->gshape = fbuf(B * NB); c->gpose = fbuf(B * J * 3);
    c->gtransl = fbuf(B * 3); c->losses = zbuf(B * 4);
    c->gx = zbuf(B * xd); c->xeval = zbuf(B * xd);
    c->step = (int *)dev_alloc(B * sizeof(int));
    c->lb = psi::LbfgsState();
    c->lbp = psi::LbfgsParams();
    c->lb_info = nullptr;
    if (cfg->optimizer == 1) {
        const int hist = cfg->lbfgs_history > 0 ? cfg->lbfgs_history : 100;
        c->lbp.lr = cfg->lbfgs_lr > 0.f ? (double)cfg->lbfgs_lr : 1.0;
        c->lbp.tol_grad = cfg->lbfgs_tolerance_grad > 0.f ? (double)cfg->lbfgs_tolerance_grad : 1e-5;
        c->lbp.tol_change = cfg->lbfgs_tolerance_change > 0.f ? (double)cfg->lbfgs_tolerance_change : 1e-9;
        c->lbp.c1 = 1e-4; c->lbp.c2 = 0.9;                      // lbfgs_ls.py:54
        c->lbp.history = hist;
        c->lbp.max_iter = 1 << 30;                              // the budget is counted in closure evaluations (num_iter)
        c->lbp.max_ls = 25;
        c->lbp.zoom_max_iter = cfg->lbfgs_zoom_max > 0 ? cfg->lbfgs_zoom_max : 300;
        c->lbp.reset_lr = 0;
        c->lb.sc = (psi::LbfgsScalars *)dev_alloc(B * sizeof(psi::LbfgsScalars));
        c->lb.x_init = zbuf(B * psi::kLbStride); c->lb.d = zbuf(B * psi::kLbStride); c->lb.g0 = zbuf(B * psi::kLbStride);
        c->lb.g_prev = zbuf(B * psi::kLbStride); c->lb.bg = zbuf(B * 2 * psi::kLbStride);
        c->lb.Y = zbuf(B * hist * psi::kLbStride); c->lb.S = zbuf(B * hist * psi::kLbStride);
        c->lb.ro = (double *)dev_alloc(B * 2 * hist * sizeof(double));
        c->lb.xbest = zbuf(B * xd);
        c->lb_info = (int *)dev_alloc(B * 4 * sizeof(int));
    }
    c->neg_cnt = (int *)dev_alloc(2 * sizeof(int));
    c->body_cnt = (int *)dev_alloc(2 * B * sizeof(int));
    c->rank = 0; c->world = 1; c->batch_total = c->B;
    for (int q = 0; q < PSI_FIT_MAX_SHARDS; ++q) { c->peer_buf[q] = nullptr; c->ipc_opened[q] = nullptr; }
    c->neg_tot = (int *)dev_alloc(2 * sizeof(int));
    {   // exchange buffer + sequence number + error flag: one allocation of its own (cudaIpcGetMemHandle needs cudaMalloc memory)
        void *xp = dev_alloc((2 * PSI_FIT_MAX_SHARDS + 2) * sizeof(unsigned long long));
        c->xbuf = (unsigned long long *)xp;
        c->xseq = xp ? (int *)(c->xbuf + 2 * PSI_FIT_MAX_SHARDS) : nullptr;
        c->xerr = xp ? c->xseq + 1 : nullptr;
        if (xp && cudaMemsetAsync(xp, 0, (2 * PSI_FIT_MAX_SHARDS + 2) * sizeof(unsigned long long), st) != cudaSuccess) rc = PSI_ERR_ALLOC;
    }
    c->loop_left = (int *)dev_alloc(sizeof(int));
    c->lbs_ws_bytes = psi_lbs_bwd_workspace_bytes(model, c->B) + 64;
    c->lbs_ws = (float *)dev_alloc(c->lbs_ws_bytes);
    if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;
    if (rc != PSI_OK) { psi_fit_destroy(c); return rc; }
    *out = c;
    return PSI_OK;
}

// loop_mode 1: one iteration as an executable graph, replayed num_iter times by the host
static int build_iteration_graph(psi_fit_ctx *c) {
    cudaStream_t gs = c->gstream;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return (int)e;
    const int rc = psi::enqueue_iteration(c, gs);
    e = cudaStreamEndCapture(gs, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return (int)e;
    e = cudaGraphInstantiate(&c->exec, graph, 0);
    cudaGraphDestroy(graph);
    return e == cudaSuccess ? PSI_OK : (int)e;
}

// loop_mode 0: head kernel -> conditional WHILE node whose body graph is the captured iteration.  The body's
// last kernel (fit_step) decrements *loop_left and re-arms the condition, so the host launches the whole
// num_iter-iteration loop once (300 cudaGraphLaunch calls per fit before).
static int build_loop_graph(psi_fit_ctx *c) {
    cudaStream_t gs = c->gstream;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaGraphCreate(&g, 0);
    if (e != cudaSuccess) return (int)e;
    int rc = PSI_OK;
    do {
        e = cudaGraphConditionalHandleCreate(&c->loop_cond, g, 0, 0);
        if (e != cudaSuccess) break;
        cudaGraphNode_t head = nullptr, loop = nullptr;
        cudaKernelNodeParams kp = {};
        void *args[] = {&c->loop_cond, &c->loop_left};
        kp.func = (void *)psi::fit_loop_head_kernel;
        kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
        e = cudaGraphAddKernelNode(&head, g, nullptr, 0, &kp);
        if (e != cudaSuccess) break;
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = c->loop_cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        e = cudaGraphAddNode(&loop, g, &head, 1, &np);
        if (e != cudaSuccess) break;
        cudaGraph_t body = np.conditional.phGraph_out[0];
        e = cudaStreamBeginCaptureToGraph(gs, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) break;
        // `unroll` iterations per pass of the WHILE body: the conditional node costs a few microseconds per pass
        for (int u = 0; u < c->unroll && rc == PSI_OK; ++u) {
            c->capturing_loop = u == c->unroll - 1;      // the last fit_step of the body counts the pass
            rc = psi::enqueue_iteration(c, gs);
        }
        c->capturing_loop = 0;
        cudaGraph_t captured = nullptr;
        e = cudaStreamEndCapture(gs, &captured);     // == body
        if (rc != PSI_OK || e != cudaSuccess) break;
        e = cudaGraphInstantiate(&c->loop_exec, g, 0);
    } while (0);
    cudaGraphDestroy(g);
    if (rc != PSI_OK) return rc;
    return e == cudaSuccess ? PSI_OK : (int)e;
}

// the loop's starting state on stream `st`: x = xhr_init, no NN hints, optimiser state cleared
static int reset_state(psi_fit_ctx *c, const float *xhr_init, cudaStream_t st) {
    using namespace psi;
    const size_t xd = 9 + 10 + (size_t)c->latent + 2 * (size_t)c->ncomp, n = (size_t)c->B * xd;
    cudaError_t e = cudaMemcpyAsync(c->x, xhr_init, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->nnhint, 0xff, (size_t)c->B * c->nu * sizeof(int), st);   // -1: no hint yet
    if (e != cudaSuccess) return (int)e;
    fit_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->am, c->av, c->step, (long)n, c->B, c->neg_cnt, c->body_cnt);
    PSI_LAUNCHED();
    if (c->world > 1) {     // a new fit: the peers' stale slots of the last one must not match
        fit_seq_kernel<<<1, 1, 0, st>>>(c->xseq);
        PSI_LAUNCHED();
    }
    if (c->cfg.optimizer == 1) {
        e = cudaMemcpyAsync(c->lb.xbest, xhr_init, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
        fit_lbfgs_reset_kernel<<<(unsigned)((c->B + 127) / 128), 128, 0, st>>>(c->lb.sc, c->B);
        PSI_LAUNCHED();
    }
    return PSI_OK;
}

int psi_fit_begin(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride, int num_iter,
                  psi_stream_t stream) {
    using namespace psi;
    Range nvtx_range("psi_fit_begin");
    if (!c || !xhr_init || !cam || num_iter < 0) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t xd = 9 + 10 + (size_t)c->latent + 2 * (size_t)c->ncomp, n = (size_t)c->B * xd;
    cudaError_t e = cudaMemcpyAsync(c->x0, xhr_init, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    if (const int rrc = reset_state(c, xhr_init, st)) return rrc;
    // one [3x4] transform per body (stride 0 broadcasts a shared one)
    fit_cam_kernel<<<(unsigned)((c->B * 12 + 255) / 256), 256, 0, st>>>(cam, cam_bstride, c->B, c->cam);
    PSI_LAUNCHED();
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    const bool graph_path = c->cfg.use_graph && cs == cudaStreamCaptureStatusNone && num_iter > 1;
    // a sharded batch waits for its peers inside the iteration: the loop must run on the context's own stream
    if (c->cfg.loss_mode == 1 && c->world > 1 && !graph_path && num_iter > 0) return PSI_ERR_UNSUPPORTED;
    if (graph_path) {
        // the loop runs on the context's own stream, ordered after / before the caller's stream
        cudaStream_t gs = c->gstream;
        const bool whole = c->cfg.loop_mode == 0;
        e = cudaEventRecord(c->ev_in, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(gs, c->ev_in, 0);
        if (e != cudaSuccess) return (int)e;
        if (whole ? !c->loop_exec : !c->exec) {
            // warm the lazily-initialised pieces outside the capture (state is reset below)
            int rc = launch_step(c, 0, gs);
            if (rc == PSI_OK) rc = enqueue_iteration(c, gs);
            if (rc == PSI_OK && whole) rc = build_loop_graph(c);
            if (rc == PSI_OK && !c->exec && (!whole || c->unroll > 1)) rc = build_iteration_graph(c);
            if (rc) return rc;
            if (const int rrc = reset_state(c, xhr_init, gs)) return rrc;
        }
        const int rc = launch_step(c, 0, gs);
        if (rc) return rc;
        if (whole) {
            for (int it = 0; it < num_iter % c->unroll; ++it) {      // the remainder, one launch per iteration
                e = cudaGraphLaunch(c->exec, gs);
                if (e != cudaSuccess) return (int)e;
            }
            if (num_iter / c->unroll > 0) {
                fit_loop_set_kernel<<<1, 1, 0, gs>>>(c->loop_left, num_iter / c->unroll);
                PSI_LAUNCHED();
                e = cudaGraphLaunch(c->loop_exec, gs);
                if (e != cudaSuccess) return (int)e;
            }
        } else {
            for (int it = 0; it < num_iter; ++it) {
                e = cudaGraphLaunch(c->exec, gs);
                if (e != cudaSuccess) return (int)e;
            }
        }
        e = cudaEventRecord(c->ev_out, gs);
        if (e != cudaSuccess) return (int)e;
        c->pending_join = 1;
        c->ev_out_recorded = 1;
    } else {
        int rc = launch_step(c, 0, st);
        for (int it = 0; it < num_iter && rc == PSI_OK; ++it) rc = enqueue_iteration(c, st);
        if (rc) return rc;
        c->pending_join = 0;
        c->ev_out_recorded = 0;
    }
    return PSI_OK;
}

int psi_fit_end(psi_fit_ctx *c, float *xhr_out, float *losses_out, psi_stream_t stream) {
    psi::Range nvtx_range("psi_fit_end");
    if (!c || !xhr_out) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (c->pending_join) {
        e = cudaStreamWaitEvent(st, c->ev_out, 0);
        if (e != cudaSuccess) return (int)e;
        c->pending_join = 0;
    }
    const size_t n = (size_t)c->B * (9 + 10 + (size_t)c->latent + 2 * (size_t)c->ncomp);
    // Adam: the vector after the last step.  L-BFGS: the last accepted point (c->x is a trial point of a line search)
    e = cudaMemcpyAsync(xhr_out, c->cfg.optimizer == 1 ? c->lb.xbest : c->x, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
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

int psi_fit_exchange_handle(psi_fit_ctx *c, void *h_handle64) {
    if (!c || !h_handle64) return PSI_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t *)h_handle64, c->xbuf);
    return e == cudaSuccess ? PSI_OK : (int)e;
}

void *psi_fit_exchange_ptr(psi_fit_ctx *c) { return c ? (void *)c->xbuf : nullptr; }

int psi_fit_set_peers(psi_fit_ctx *c, int rank, int world, int batch_total, const void *const *h_handles, void *const *ptrs) {
    if (!c || rank < 0 || world < 1 || rank >= world || world > PSI_FIT_MAX_SHARDS || batch_total < c->B) return PSI_ERR_BAD_ARG;
    if (c->cfg.loss_mode != 1) return PSI_ERR_UNSUPPORTED;
    if (world > 1 && !h_handles && !ptrs) return PSI_ERR_BAD_ARG;
    // graphs captured for another peer set hold the old pointers
    if (c->exec) { cudaGraphExecDestroy(c->exec); c->exec = nullptr; }
    if (c->loop_exec) { cudaGraphExecDestroy(c->loop_exec); c->loop_exec = nullptr; }
    for (int q = 0; q < PSI_FIT_MAX_SHARDS; ++q) {
        if (c->ipc_opened[q]) { cudaIpcCloseMemHandle(c->ipc_opened[q]); c->ipc_opened[q] = nullptr; }
        c->peer_buf[q] = nullptr;
    }
    for (int q = 0; q < world; ++q) {
        if (q == rank) { c->peer_buf[q] = c->xbuf; continue; }
        if (ptrs && ptrs[q]) {
            // another context of this process; on another device the peer mapping has to be enabled once
            cudaPointerAttributes pa;
            int dev = 0;
            if (cudaPointerGetAttributes(&pa, ptrs[q]) == cudaSuccess && cudaGetDevice(&dev) == cudaSuccess && pa.device != dev) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(pa.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int)e;
                (void)cudaGetLastError();
            }
            c->peer_buf[q] = (unsigned long long *)ptrs[q];
        } else if (h_handles && h_handles[q]) {
            void *p = nullptr;
            cudaIpcMemHandle_t h;
            memcpy(&h, h_handles[q], sizeof(h));
            const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return (int)e;
            c->ipc_opened[q] = p;
            c->peer_buf[q] = (unsigned long long *)p;
        } else {
            return PSI_ERR_BAD_ARG;
        }
    }
    c->rank = rank; c->world = world; c->batch_total = batch_total;
    return PSI_OK;
}

// (buffer, bytes) of a traceable object
static const void *trace_buffer(const psi_fit_ctx *c, int what, size_t *bytes) {
    const size_t B = (size_t)c->B, xd = 9 + 10 + (size_t)c->latent + 2 * (size_t)c->ncomp, f = sizeof(float);
    switch (what) {
        case PSI_FIT_TRACE_X_EVAL: *bytes = B * xd * f; return c->xeval;
        case PSI_FIT_TRACE_GRAD_X: *bytes = B * xd * f; return c->gx;
        case PSI_FIT_TRACE_VERTS: *bytes = B * c->V * 3 * f; return c->verts;
        case PSI_FIT_TRACE_SDF: *bytes = B * c->V * f; return c->sdfv;
        case PSI_FIT_TRACE_SDF_GRAD: *bytes = B * c->V * 3 * f; return c->sdfg;
        case PSI_FIT_TRACE_NN_DIST: *bytes = B * c->nu * f; return c->nnd;
        case PSI_FIT_TRACE_NN_IDX: *bytes = B * c->nu * sizeof(int); return c->nni;
        case PSI_FIT_TRACE_QUERY_IDS: *bytes = (size_t)c->nu * sizeof(int); return c->csel;
        case PSI_FIT_TRACE_LOSSES: *bytes = B * 4 * f; return c->losses;
        case PSI_FIT_TRACE_X: *bytes = B * xd * f; return c->x;
        case PSI_FIT_TRACE_ADAM_M: *bytes = B * xd * f; return c->am;
        case PSI_FIT_TRACE_ADAM_V: *bytes = B * xd * f; return c->av;
        case PSI_FIT_TRACE_POSE6D: *bytes = B * c->num_rot * 6 * f; return c->rot6d;
        case PSI_FIT_TRACE_LBFGS_STATE: *bytes = c->cfg.optimizer == 1 ? B * 4 * sizeof(int) : 0; return c->lb_info;
        case PSI_FIT_TRACE_LBFGS_BEST: *bytes = c->cfg.optimizer == 1 ? B * xd * f : 0; return c->lb.xbest;
        case PSI_FIT_TRACE_EXCHANGE: *bytes = 4 * sizeof(int); return c->xseq;
        default: *bytes = 0; return nullptr;
    }
}

size_t psi_fit_trace_bytes(const psi_fit_ctx *c, int what) {
    size_t bytes = 0;
    if (c) trace_buffer(c, what, &bytes);
    return bytes;
}

int psi_fit_trace(psi_fit_ctx *c, int what, void *dst, size_t dst_bytes, psi_stream_t stream) {
    if (!c || !dst) return PSI_ERR_BAD_ARG;
    size_t bytes = 0;
    const void *src = trace_buffer(c, what, &bytes);
    if (!src || bytes == 0) return PSI_ERR_BAD_ARG;
    if (dst_bytes < bytes) return PSI_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->ev_out_recorded) {   // the loop ran (or still runs) on the context's own stream
        const cudaError_t e = cudaStreamWaitEvent(st, c->ev_out, 0);
        if (e != cudaSuccess) return (int)e;
    }
    if (what == PSI_FIT_TRACE_LBFGS_STATE) {
        psi::fit_lbfgs_info_kernel<<<(unsigned)((c->B + 127) / 128), 128, 0, st>>>(c->lb.sc, c->B, c->lb_info);
        PSI_LAUNCHED();
    }
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? PSI_OK : (int)e;
}

int psi_fit_profile(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride, int warm_iters,
                    int timed_iters, float *h_ms, const char **h_names, int max_launches, psi_stream_t stream) {
    psi::Range nvtx_range("psi_fit_profile");
    using namespace psi;
    if (!c || !xhr_init || !cam || !h_ms || !h_names || warm_iters < 0 || timed_iters < 1 || max_launches < 1)
        return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int graph = c->cfg.use_graph;
    c->cfg.use_graph = 0;                                   // eager iterations on the caller's stream
    int rc = psi_fit_begin(c, xhr_init, cam, cam_bstride, warm_iters, stream);
    c->cfg.use_graph = graph;
    if (rc) return rc;
    LaunchRecorder rec;
    rec.st = st;
    for (int i = 0; i < 65; ++i)
        if (cudaEventCreate(&rec.ev[i]) != cudaSuccess) {
            for (int k = 0; k < i; ++k) cudaEventDestroy(rec.ev[k]);
            return PSI_ERR_ALLOC;
        }
    int n = 0;
    for (int i = 0; i < max_launches; ++i) h_ms[i] = 0.f;
    for (int it = 0; it < timed_iters && rc == PSI_OK; ++it) {
        rec.n = 0;
        cudaEventRecord(rec.ev[0], st);
        recorder_set(&rec);
        rc = enqueue_iteration(c, st);
        recorder_set(nullptr);
        if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;
        n = rec.n < max_launches ? rec.n : max_launches;
        for (int i = 0; i < n && rc == PSI_OK; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, rec.ev[i], rec.ev[i + 1]);
            h_ms[i] += ms / (float)timed_iters;
            h_names[i] = rec.names[i];
        }
    }
    for (int i = 0; i < 65; ++i) cudaEventDestroy(rec.ev[i]);
    return rc == PSI_OK ? n : -rc - 1000;                   // >= 0: launches per iteration
}

}  // extern "C"
