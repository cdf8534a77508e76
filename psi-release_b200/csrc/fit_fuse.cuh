// Pieces of the fitting loop's loss that are fused into LBS kernels (lbs.cu) when the loop runs
// through psi_fit_* (fit.cu).  The generic psi_lbs_fwd / psi_lbs_bwd entry points pass none.
#pragma once
#include "sdf_sample.cuh"
namespace psi {

// scene-SDF lookup at every vertex, in the skinning kernel's epilogue (fitting_habitat.py:145-152)
struct SdfFuse {
    SdfGrid g;
    float *sdfv, *sdfg;      // [B,V], [B,V,3]
    float *partial;          // [B, vertex chunks, 2]: (sum of -sdf, count) over sdf < 0, per 256-vertex chunk
    // batch-coupled loss (psi_fit_config.loss_mode 1): the batch-wide count of penetrating vertices is
    // accumulated with integer atomics into neg_cnt[iteration & 1] (iteration = step[0]); NULL otherwise
    int *neg_cnt;
    const int *step;
    // per-body count of penetrating vertices, [2][B] by iteration parity (integer atomics: order-independent): what
    // lbs_vertex_bwd<FIT> divides by in loss_mode 0 -- one load instead of a sum over the chunk partials; NULL: not kept
    int *body_cnt;
    int nbodies;
};

// dL/dverts of the contact robustifier (fitting_habitat.py:133-141) and the collision mean
// (:155-160), evaluated at the head of the vertex backward kernel instead of a pass of its own
struct VGradFuse {
    const float *verts, *scene, *sdfv, *sdfg, *partial, *nnd;
    const int *nni, *cslot;
    const float *cweight;
    float w_contact, w_coll, robust_c;
    int nu, np_sdf, num_contact;
    float *cpart;            // [B, vertex chunks]: contact-loss partial sums
    const int *neg_cnt;      // loss_mode 1: batch-wide penetration count [2] (see SdfFuse), else NULL
    const int *body_cnt;     // loss_mode 0: per-body penetration counts [2][B] (see SdfFuse), or NULL (sum the partials)
    const int *step;
    float bdiv;              // loss_mode 1: bodies in the batch (the contact mean runs over B x Nc), else 1
};

}  // namespace psi
