/*
 * psi_b200.h -- C ABI of libpsi_b200.so: the PSI fitting hot path on B200 (sm_100a).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_`;
 *   - FP32 values, int32 indices, row-major, innermost dimension contiguous;
 *   - batch strides are in ELEMENTS (floats); a batch stride of 0 shares one cloud /
 *     one scene across the whole batch (the reference replicates them B times:
 *     source/fitting_proxe.py:90,96);
 *   - outputs are caller-owned and OVERWRITTEN (no pre-zero contract, unlike
 *     chamfer_pytorch/chamfer.cu:177-178 whose memsets are commented out);
 *   - work is enqueued on `stream` (a cudaStream_t) and the call returns without
 *     synchronising; never prints; handles (model / index / fit context) belong to the device
 *     that was current when they were created and calls on them must be made with that device
 *     current; different handles may be used from different threads and on different devices of
 *     one process (process-wide state: a launch counter and a mutex-guarded cache of the
 *     per-device kernel attributes);
 *   - return value: 0 = ok, >0 = a cudaError_t from the launch, <0 = PSI_ERR_*.
 *
 * Each entry point names the reference interface it replaces.
 */
#ifndef PSI_B200_H
#define PSI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *psi_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define PSI_API __attribute__((visibility("default")))
#else
#define PSI_API
#endif

#define PSI_OK 0
#define PSI_ERR_BAD_ARG (-1)      /* null pointer, negative size, bad alignment */
#define PSI_ERR_WORKSPACE (-2)    /* workspace too small */
#define PSI_ERR_UNSUPPORTED (-3)  /* shape outside what the kernels were built for */
#define PSI_ERR_ALLOC (-4)        /* device allocation failed (model upload only) */

#define PSI_ABI_VERSION 3   /* 3: psi_fit_config.{loop_mode,loop_unroll,loss_mode,optimizer,lbfgs_*}, psi_fit_trace*, psi_fit_exchange_*, psi_fit_set_peers;
                               2: psi_fit_config.nn_mode, psi_fit_profile, psi_nn_index_query_mode mode 3 */

/* ABI version of the loaded library (PSI_ABI_VERSION it was built with). */
PSI_API int psi_abi_version(void);
/* Number of kernel launches this library has enqueued since load (statistics only: a relaxed
 * atomic counter, the one piece of process-wide state; launches captured into a CUDA graph
 * count once, at capture). */
PSI_API unsigned long long psi_launch_count(void);
/* Static string for a PSI_ERR_* or cudaError_t value. */
PSI_API const char *psi_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Chamfer / nearest neighbour.
 * Replaces: chamfer.forward / chamfer.backward (chamfer_pytorch/chamfer_cuda.cpp:17-33,
 * kernels chamfer_pytorch/chamfer.cu:12-195), reached from
 * chamfer_pytorch/dist_chamfer.py:15-46 and dist_chamfer_idx.py:11-46.
 *
 * Distance definition (bit-exact with the reference build):
 *   dx = s.x - q.x ...;  d = fma(dz,dz, fma(dx,dx, rn(dy*dy)));  first minimum wins
 *   (lowest index on ties).  Scene points must be finite.  A query whose distances are all NaN
 *   or all +inf (a diverged body) gets index 0 and its distance to point 0, which is what the
 *   reference's `k == 0 ||` clauses return (chamfer.cu:36,121,126).
 * ---------------------------------------------------------------------------------------- */

/* Bytes of scratch psi_nn_fwd / psi_chamfer_fwd need for these sizes (may be 0). */
PSI_API size_t psi_nn_workspace_bytes(int B, int n, int m);

/* One direction: for every query q[b,j] (j<n) the nearest of s[b,0..m): squared distance
 * and index.  q_bstride / s_bstride in floats (0 = shared).  dist,idx: [B,n].
 * idx may be NULL.  m == 0 leaves outputs untouched (as the reference's loops do). */
PSI_API int psi_nn_fwd(const float *q, long q_bstride, int B, int n,
               const float *s, long s_bstride, int m,
               float *dist, int *idx,
               void *workspace, size_t workspace_bytes, psi_stream_t stream);

/* Both directions, the reference signature: xyz1 [B,n,3], xyz2 [B,m,3] contiguous.
 * dist2/idx2 may both be NULL to skip the xyz2->xyz1 direction (the fitting loss
 * discards it: source/fitting_habitat.py:138). */
PSI_API int psi_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int n, int m,
                    float *dist1, float *dist2, int *idx1, int *idx2,
                    void *workspace, size_t workspace_bytes, psi_stream_t stream);

/* Gradient of sum_j graddist[b,j]*dist[b,j] w.r.t. the queries only (a deterministic
 * gather): grad_q[b,j] = 2*graddist[b,j]*(q[b,j]-s[b,idx[b,j]]).  grad_q: [B,n,3]. */
PSI_API int psi_nn_bwd(const float *q, long q_bstride, int B, int n,
               const float *s, long s_bstride, int m,
               const float *graddist, const int *idx, float *grad_q, psi_stream_t stream);

/* The reference backward: both clouds receive gradient from both directions
 * (chamfer.cu:155-195).  gradxyz1 [B,n,3], gradxyz2 [B,m,3] are OVERWRITTEN (zeroed
 * here, then accumulated; the scatter into the other cloud uses float atomics exactly as
 * the reference does, so its rounding is order-dependent).  graddist2/idx2 may be NULL. */
PSI_API int psi_chamfer_bwd(const float *xyz1, const float *xyz2, int B, int n, int m,
                    const float *graddist1, const float *graddist2,
                    const int *idx1, const int *idx2,
                    float *gradxyz1, float *gradxyz2, psi_stream_t stream);

/* Exact NN against a STATIC scene through a cluster index (same outputs as psi_nn_fwd with
 * s_bstride == 0, bit for bit: the index only skips clusters whose floating-point lower bound
 * exceeds a distance already seen).  The fitting loop queries the same scene points
 * (self.s_verts_batch, source/fitting_habitat.py:93-96) 300 x B times; the index is built once.
 * h_points: HOST array [m,3], finite values.  Returns after the upload completed. */
typedef struct psi_nn_index psi_nn_index;
PSI_API int psi_nn_index_create(psi_nn_index **out, const float *h_points, int m, psi_stream_t stream);
PSI_API void psi_nn_index_destroy(psi_nn_index *ix);
PSI_API size_t psi_nn_index_bytes(const psi_nn_index *ix);
/* Query j of body b is q[b*q_bstride + 3*(qsel ? qsel[j] : j)]: qsel (device int[n], or NULL)
 * selects rows of a larger per-body array, e.g. the contact vertices of the body mesh
 * (body_verts_batch[:, vid, :], source/fitting_habitat.py:135) without a gather pass.
 * -> dist, idx [B,n]; idx (original point order) may be NULL. */
PSI_API int psi_nn_index_query(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                       const int *qsel, float *dist, int *idx, psi_stream_t stream);
/* Same, with a per-query HINT buffer (device int[B*n], in/out): on entry the cluster that held the
 * query's nearest neighbour last time (or -1), on exit this call's.  It only seeds the pruning
 * bound (that cluster is visited first) -- results are identical with any hint contents. */
PSI_API int psi_nn_index_query_hint(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                            const int *qsel, float *dist, int *idx, int *hint, psi_stream_t stream);
/* Same, choosing the schedule: mode 0 = by query count (the two above), 1 = one warp per query,
 * 2 = one thread per query, 3 = one warp per GROUP of 32 consecutive queries of a body sharing one
 * tree walk (2 and 3 are fast when consecutive queries are spatial neighbours; 3 is what the
 * fitting loop uses).  All modes return identical bits. */
PSI_API int psi_nn_index_query_mode(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                            const int *qsel, float *dist, int *idx, int *hint, int mode,
                            psi_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Scene SDF lookup.
 * Replaces: the normalise + F.grid_sample(sdf, grid[...,[2,1,0]], padding_mode='border')
 * block at source/fitting_habitat.py:145-152, fitting_proxe.py:144-151, train_s2.py:182-189,
 * utils/utils_eval_collision_habitat.py:121-128, with torch-1.2 semantics
 * (align_corners=True).
 *
 * sdf: S grids [S,D,D,D], index [x][y][z], z contiguous.  h_grid_min/h_grid_max: HOST
 * arrays [S,3].  verts [B,V,3] in the scene frame.  body_scene: device int[B] giving the
 * grid of each body, or NULL (all bodies use grid 0).  out [B,V]; grad [B,V,3] =
 * d out / d verts (may be NULL).  partial: NULL or float[B, psi_sdf_num_partials(V), 2]
 * receiving per-chunk (sum of -sdf over sdf<0, count of sdf<0) in a fixed order
 * (the collision reduction of fitting_habitat.py:155-160, reduced deterministically).
 * ---------------------------------------------------------------------------------------- */
PSI_API int psi_sdf_num_partials(int V);
PSI_API int psi_sdf_fwd(const float *sdf, int S, int D, const float *h_grid_min, const float *h_grid_max,
                const float *verts, int B, int V, const int *body_scene,
                float *out, float *grad, float *partial, psi_stream_t stream);

/* grad_verts[b,v,:] = grad_out[b,v] * grad[b,v,:]  (the backward of psi_sdf_fwd). */
PSI_API int psi_sdf_bwd(const float *grad_out, const float *grad, long count, float *grad_verts,
                psi_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SMPL-X linear blend skinning.
 * Replaces: smplx.SMPLX.forward -> smplx.lbs.lbs (pip smplx==0.1.13; math vendored at
 * human_body_prior/body_model/lbs.py:34-262; call sites source/fitting_habitat.py:126-129)
 * fused with the translation (body_model.py:246-247) and with
 * GeometryTransformer.verts_transform (source/cvae.py:141-149).
 *
 * The model handle owns device copies of the constants in the kernels' own layout.
 * ---------------------------------------------------------------------------------------- */
typedef struct psi_lbs_model psi_lbs_model;

/* Host arrays in the reference's layout: v_template [V,3], shapedirs [V,3,NB],
 * posedirs [P,V*3] with P=(J-1)*9 (body_model.py:123-125), J_regressor [J,V],
 * weights [V,J], parents int[J] (parents[0] ignored).  Uploads on `stream`, returns after
 * the upload completed. */
PSI_API int psi_lbs_model_create(psi_lbs_model **out, int V, int J, int NB,
                         const float *h_v_template, const float *h_shapedirs,
                         const float *h_posedirs, const float *h_J_regressor,
                         const float *h_weights, const int *h_parents, psi_stream_t stream);
PSI_API void psi_lbs_model_destroy(psi_lbs_model *m);
/* Bytes of constants resident in HBM for this model (for roofline accounting). */
PSI_API size_t psi_lbs_model_bytes(const psi_lbs_model *m);

/* Floats of per-call state psi_lbs_fwd saves for psi_lbs_bwd. */
PSI_API size_t psi_lbs_saved_floats(const psi_lbs_model *m, int B);

/* betas [B,NB] (shape+expression), pose [B,J*3] axis-angle (after hand PCA + pose_mean),
 * transl [B,3] or NULL, cam [B,12] rows of a 3x4 rigid transform applied last
 * (cam_bstride floats between bodies, 0 = shared) or NULL.  rot_in [B,num_rot,9] (or NULL with
 * num_rot 0): row-major rotation matrices used for joints 0..num_rot-1 INSTEAD of
 * Rodrigues(pose) -- the fused fitting loop hands VPoser's / the 6D global rotation over
 * without the matrix -> axis-angle -> matrix round trip (cvae.py:72-80 + lbs.py:165-192).
 * verts [B,V,3]; joints [B,J,3] (posed joints + transl, camera frame) or NULL;
 * saved: psi_lbs_saved_floats floats. */
PSI_API int psi_lbs_fwd(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                const float *transl, const float *cam, long cam_bstride,
                const float *rot_in, int num_rot,
                float *verts, float *joints, float *saved, psi_stream_t stream);

/* Given grad_verts [B,V,3] (and grad_joints [B,J,3] or NULL): grad_betas [B,NB],
 * grad_pose [B,J*3] (zero for joints < num_rot), grad_transl [B,3] (may be NULL),
 * grad_rot [B,num_rot,9] = d loss / d rot_in (NULL with num_rot 0).  workspace:
 * psi_lbs_bwd_workspace_bytes bytes. */
PSI_API size_t psi_lbs_bwd_workspace_bytes(const psi_lbs_model *m, int B);
PSI_API int psi_lbs_bwd(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                const float *cam, long cam_bstride, const float *saved,
                const float *grad_verts, const float *grad_joints,
                float *grad_betas, float *grad_pose, float *grad_transl,
                float *grad_rot, int num_rot,
                void *workspace, size_t workspace_bytes, psi_stream_t stream);

/* psi_lbs_bwd with an optional side stream (cudaStream_t) and two cudaEvent_t (fork, join): the
 * per-joint gather-reduce runs on the side stream next to the d-coefficient GEMM.  NULL = serial. */
PSI_API int psi_lbs_bwd2(const psi_lbs_model *m, int B, const float *betas, const float *pose,
                 const float *cam, long cam_bstride, const float *saved,
                 const float *grad_verts, const float *grad_joints,
                 float *grad_betas, float *grad_pose, float *grad_transl,
                 float *grad_rot, int num_rot,
                 void *workspace, size_t workspace_bytes, psi_stream_t stream,
                 psi_stream_t side_stream, void *ev_fork, void *ev_join);

/* ------------------------------------------------------------------------------------------
 * The fused fitting loop.
 * Replaces: FittingOP.cal_loss + loss.backward + optimizer.step, i.e. the body of the loop at
 * source/fitting_habitat.py:177-191 (cal_loss :103-164; Adam :76), for a batch of bodies in one
 * scene, as 13 kernel launches per iteration replayed from a CUDA graph.  Loss = the SUM over
 * bodies of the reference's B=1 loss (bodies never interact).
 * ---------------------------------------------------------------------------------------- */
typedef struct psi_fit_config {
    int B;                /* bodies in the batch */
    int use_graph;        /* capture one iteration into a CUDA graph and replay it */
    float w_rec, w_vposer, w_contact, w_collision;   /* lossconfig, fitting_habitat.py:261-266 */
    float robust_c;       /* 1.0 (fitting_habitat.py:141) or 0.01 (fitting_proxe.py:139) */
    float lr, beta1, beta2, eps;                     /* torch.optim.Adam: init_lr_h, .9, .999, 1e-8 */
    int nn_mode;          /* schedule of the in-loop NN query (psi_nn_index_query_mode); 0 = default (3) */
    int loop_mode;        /* with use_graph: 0 = the WHOLE loop as one graph launch (a conditional WHILE node whose
                             body is the iteration, counted down on the device), 1 = one graph launch per iteration */
    int loop_unroll;      /* loop_mode 0: iterations captured per pass of the WHILE body (0 = default); a remainder
                             of num_iter runs as single-iteration launches first */
    int loss_mode;        /* 0 = independent: sum over bodies of the reference's B=1 loss (the shipped scripts,
                             fitting_habitat.py:254); 1 = batch: the reference's batch-coupled means
                             (fitting_proxe.py:105,110,139,155-160 with B > 1; demo.ipynb cell 16) */
    int optimizer;        /* 0 = Adam (the fitting scripts', fitting_habitat.py:76); 1 = L-BFGS with strong-Wolfe line search
                             (human_body_prior/optimizers/lbfgs_ls.py:54-183,275-463), one independent optimiser per
                             body, num_iter = closure evaluations; loss_mode 0 only */
    float lbfgs_lr;             /* initial step of every line search (lbfgs_ls.py lr; 0 = 1.0) */
    float lbfgs_tolerance_grad;     /* stop when |g|_1 <= this   (0 = 1e-5, lbfgs_ls.py:214-216) */
    float lbfgs_tolerance_change;   /* stop on a step / loss change below this (0 = 1e-9) */
    int lbfgs_history;          /* curvature pairs kept (0 = 100) */
    int lbfgs_zoom_max;         /* zoom-phase iteration bound (0 = 300: the reference passes the optimiser's max_iter) */
} psi_fit_config;

typedef struct psi_fit_ctx psi_fit_ctx;

/* model/index/scene_points [m,3]/sdf [D,D,D] are borrowed (device objects that must outlive the
 * context; V,J,NB = the model's sizes).  VPoser decoder weights in the reference's nn.Linear
 * layout: h_W1 [hidden,latent], h_W2 [hidden,hidden], h_W3 [nbody*6,hidden] (+ biases);
 * hand PCA components h_hand_l/r [ncomp,45] and pose_mean [J*3] as smplx holds them;
 * h_contact_ids: the concatenated contact vertex ids (duplicates allowed, cvae.py:105-112). */
PSI_API int psi_fit_create(psi_fit_ctx **out, const psi_lbs_model *model, int V, int J, int NB,
                   const psi_nn_index *index, const float *scene_points, const float *sdf, int D,
                   const float *h_grid_min, const float *h_grid_max,
                   const float *h_W1, const float *h_b1, const float *h_W2, const float *h_b2,
                   const float *h_W3, const float *h_b3, int latent, int hidden, int nbody,
                   const float *h_hand_l, const float *h_hand_r, const float *h_pose_mean, int ncomp,
                   const int *h_contact_ids, int num_contact, const psi_fit_config *cfg,
                   psi_stream_t stream);
PSI_API void psi_fit_destroy(psi_fit_ctx *c);
/* xhr_init [B,75] = [transl 3 | rot6d 6 | betas 10 | vposer z 32 | lhand 12 | rhand 12]
 * (GeometryTransformer.convert_to_6D_rot of the generated body, fitting_habitat.py:174-175);
 * cam [B,12] rows of the 3x4 camera->scene transform (cam_bstride 0 = one for all bodies).
 * Runs num_iter iterations; xhr_out [B,75] = fitted vector; losses_out [B,4] (may be NULL) =
 * the four weighted terms of the last evaluated iteration.  Device pointers. */
PSI_API int psi_fit_run(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride,
                int num_iter, float *xhr_out, float *losses_out, psi_stream_t stream);
/* The same call split in two, so that several contexts (e.g. two halves of a batch) can run
 * their loops CONCURRENTLY on their own streams: psi_fit_begin enqueues everything and returns;
 * psi_fit_end makes `stream` wait for the loop and copies the results out. */
PSI_API int psi_fit_begin(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride,
                  int num_iter, psi_stream_t stream);
PSI_API int psi_fit_end(psi_fit_ctx *c, float *xhr_out, float *losses_out, psi_stream_t stream);
PSI_API int psi_fit_launches_per_iteration(void);

/* loss_mode 1 with the batch SHARDED over several contexts -- GPUs of one node, one process per GPU, or several
 * contexts of one process (SURVEY.md 8(e) option 2).  The batch-coupled loss needs, every iteration, the number of
 * penetrating vertices of the WHOLE batch (fitting_proxe.py:155-158) and the total body count for its means.  Each
 * context owns a small exchange buffer; once the shards know each other's buffers, every iteration PUBLISHES its
 * count into the peers' buffers with plain stores through the peer mapping (NVLink) right after the SDF pass and
 * COLLECTS theirs right before the vertex backward kernel -- 8 bytes per peer, hidden behind the NN walk; no NCCL
 * call, no host involvement, inside the same CUDA graph.  Results are bit-identical to one context fitting the
 * union batch.
 *   psi_fit_exchange_handle: the buffer's cudaIpcMemHandle_t (64 bytes, host) for ANOTHER PROCESS;
 *   psi_fit_exchange_ptr:    its device pointer for another context of the SAME process;
 *   psi_fit_set_peers:       rank / world (<= 16) / bodies of the whole batch, and per shard EITHER its IPC handle
 *                            (h_handles[q], 64 bytes) OR its device pointer (ptrs[q]); entry `rank` is ignored.
 * Every shard must run the same psi_fit_run / psi_fit_begin calls (same num_iter) with use_graph = 1. */
#define PSI_FIT_MAX_SHARDS 16
PSI_API int psi_fit_exchange_handle(psi_fit_ctx *c, void *h_handle64);
PSI_API void *psi_fit_exchange_ptr(psi_fit_ctx *c);
PSI_API int psi_fit_set_peers(psi_fit_ctx *c, int rank, int world, int batch_total, const void *const *h_handles,
                      void *const *ptrs);

/* Trace of the most recent iteration the context evaluated (parity tests, debugging): copies one of the
 * loop's own buffers to `dst` (device memory, psi_fit_trace_bytes(what) bytes) on `stream`, ordered after
 * the loop.  After psi_fit_run(num_iter = k) the "last iteration" is iteration k-1: X_EVAL is the vector it
 * was evaluated at, GRAD_X = dL/dx there (all four loss terms, before Adam), VERTS / SDF / NN are what its
 * kernels produced, X is the vector after its Adam step.  NN arrays are in QUERY order: slot s is body
 * vertex QUERY_IDS[s] (the unique contact ids in the order the loop queries them). */
#define PSI_FIT_TRACE_X_EVAL 0     /* float [B,75] */
#define PSI_FIT_TRACE_GRAD_X 1     /* float [B,75] */
#define PSI_FIT_TRACE_VERTS 2      /* float [B,V,3] scene frame */
#define PSI_FIT_TRACE_SDF 3        /* float [B,V] */
#define PSI_FIT_TRACE_SDF_GRAD 4   /* float [B,V,3] */
#define PSI_FIT_TRACE_NN_DIST 5    /* float [B,nu] */
#define PSI_FIT_TRACE_NN_IDX 6     /* int   [B,nu] original scene point index */
#define PSI_FIT_TRACE_QUERY_IDS 7  /* int   [nu] */
#define PSI_FIT_TRACE_LOSSES 8     /* float [B,4] */
#define PSI_FIT_TRACE_X 9          /* float [B,75] */
#define PSI_FIT_TRACE_ADAM_M 10    /* float [B,75] */
#define PSI_FIT_TRACE_ADAM_V 11    /* float [B,75] */
#define PSI_FIT_TRACE_POSE6D 12    /* float [B,nbody+1,6]: root 6D + VPoser decoder output of the last iteration */
#define PSI_FIT_TRACE_LBFGS_STATE 13   /* int [B,4]: (phase 0 start / 1 bracket / 2 zoom / 3 done, outer iterations, closure
                                          evaluations, curvature pairs held); optimizer 1 only */
#define PSI_FIT_TRACE_EXCHANGE 15      /* int [4]: (fit sequence number, 1 if a peer's count did not arrive in time, -, -); sharded batch loss */
#define PSI_FIT_TRACE_LBFGS_BEST 14    /* float [B,75]: last accepted point (+ best sufficient-decrease step of a search in progress) */
PSI_API size_t psi_fit_trace_bytes(const psi_fit_ctx *c, int what);   /* 0 for an unknown `what` */
PSI_API int psi_fit_trace(psi_fit_ctx *c, int what, void *dst, size_t dst_bytes, psi_stream_t stream);

/* Measurement aid (bench.py `roofline`): runs warm_iters + timed_iters iterations EAGERLY on
 * `stream` with a CUDA event recorded behind every kernel launch and returns the launch count n
 * of one iteration (< 0 on error); h_ms[i] (host) = average duration of launch i over the timed
 * iterations, h_names[i] = its kernel's name (static strings).  Synchronises the stream. */
PSI_API int psi_fit_profile(psi_fit_ctx *c, const float *xhr_init, const float *cam, long cam_bstride,
                    int warm_iters, int timed_iters, float *h_ms, const char **h_names, int max_launches,
                    psi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PSI_B200_H */
