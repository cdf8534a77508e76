// 6D rotation representation <-> rotation matrix (Gram-Schmidt), forward and backward.
// Reference: GeometryTransformer.convert_to_3D_rot, source/cvae.py:46-55 (F.normalize eps 1e-12).
#pragma once
namespace psi {

// cvae.py:46-55: x6 viewed [3,2]; columns a1 = (x0,x2,x4), a2 = (x1,x3,x5)
__device__ __forceinline__ void gs_fwd(const float *x6, float *R) {
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

__device__ __forceinline__ void gs_bwd(const float *x6, const float *dR, float *dx6) {
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

}  // namespace psi
