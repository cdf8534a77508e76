/*
 * psi_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * Plain-C restatement of the two integer/index-exact operators of the PSI
 * fitting hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (psi-release_b200) never does and fails loudly without its CUDA library.
 *
 * Follows the reference at:
 *   NN / Chamfer forward   /root/reference/chamfer_pytorch/chamfer.cu:12-134
 *       distance  d = fma(dz,dz, fma(dx,dx, rn(dy*dy))), dx = s.x - q.x   (the
 *       contraction nvcc 12.9 emits for `x2*x2+y2*y2+z2*z2`);
 *       per-chunk strict '<' and cross-chunk strict '>' reduce to "global first
 *       minimum": the LOWEST scene index wins ties (chamfer.cu:36,46,56,66,121,126).
 *   Chamfer backward       /root/reference/chamfer_pytorch/chamfer.cu:155-174
 *       g = grad*2; grad_xyz1[j] += g*(p-q); grad_xyz2[idx] -= g*(p-q)
 *   SDF trilinear lookup   torch-1.2 F.grid_sample semantics used at
 *       /root/reference/source/fitting_habitat.py:145-152
 *       (align_corners=True, padding_mode='border')
 *
 * Parity pin: on a GPU box tests/test_gpu_parity.py
 * (test_chamfer_matches_reference_cuda_build_bit_exact) checks this file against the
 * reference's own chamfer.cu compiled unmodified into oracle/_ref/ (bit-exact
 * dist + idx), and against the reference test's matmul formula
 * (chamfer_pytorch/test_chamfer.py:35-54, sum sq err < 1e-8).
 *
 * Build: see oracle/Makefile (-ffp-contract=off so only the explicit fmaf fuse).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

int psi_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void psi_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* the reference distance, scalar form (chamfer.cu:31-35 under nvcc's contraction) */
static inline float ref_dist(float qx, float qy, float qz, float sx, float sy, float sz) {
    float dx = sx - qx, dy = sy - qy, dz = sz - qz;
    float yy = dy * dy;                 /* rn(dy*dy) */
    return fmaf(dz, dz, fmaf(dx, dx, yy));
}

/* ---- one query against an SoA scene: scalar ---- */
static void nn_one_scalar(float qx, float qy, float qz, const float *sx, const float *sy,
                          const float *sz, int m, float *out_d, int *out_i) {
    float best = 0.0f;
    int bi = 0;
    for (int k = 0; k < m; ++k) {
        float d = ref_dist(qx, qy, qz, sx[k], sy[k], sz[k]);
        if (k == 0 || d < best) { best = d; bi = k; }
    }
    *out_d = best;
    *out_i = bi;
}

#if defined(__x86_64__)
/* ---- AVX2+FMA: 8 scene points per step, per-lane first-minimum, lane merge at the end.
 *      Per-lane strict '<' keeps the earliest index in that lane; the final merge takes
 *      the smallest distance and, among equal distances, the smallest index => identical
 *      to the scalar first-minimum scan for finite inputs. */
__attribute__((target("avx2,fma")))
static void nn_one_avx2(float qx, float qy, float qz, const float *sx, const float *sy,
                        const float *sz, int m, float *out_d, int *out_i) {
    const __m256 vqx = _mm256_set1_ps(qx), vqy = _mm256_set1_ps(qy), vqz = _mm256_set1_ps(qz);
    __m256 vbest = _mm256_set1_ps(INFINITY);
    __m256i vidx = _mm256_set1_epi32(0);
    __m256i vk = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
    const __m256i v8 = _mm256_set1_epi32(8);
    int k = 0;
    for (; k + 8 <= m; k += 8) {
        __m256 dx = _mm256_sub_ps(_mm256_loadu_ps(sx + k), vqx);
        __m256 dy = _mm256_sub_ps(_mm256_loadu_ps(sy + k), vqy);
        __m256 dz = _mm256_sub_ps(_mm256_loadu_ps(sz + k), vqz);
        __m256 d = _mm256_fmadd_ps(dz, dz, _mm256_fmadd_ps(dx, dx, _mm256_mul_ps(dy, dy)));
        __m256 lt = _mm256_cmp_ps(d, vbest, _CMP_LT_OQ);
        vbest = _mm256_blendv_ps(vbest, d, lt);
        vidx = _mm256_castps_si256(_mm256_blendv_ps(_mm256_castsi256_ps(vidx),
                                                    _mm256_castsi256_ps(vk), lt));
        vk = _mm256_add_epi32(vk, v8);
    }
    float bl[8];
    int il[8];
    _mm256_storeu_ps(bl, vbest);
    _mm256_storeu_si256((__m256i *)il, vidx);
    float best = INFINITY;
    int bi = 0;
    for (int l = 0; l < 8; ++l)
        if (bl[l] < best || (bl[l] == best && il[l] < bi)) { best = bl[l]; bi = il[l]; }
    /* no distance compared below +inf (NaN or overflowing query): the reference's `k == 0 ||` clauses
     * (chamfer.cu:36,121,126) leave point 0 and its distance, as the scalar scan above does */
    if (k > 0 && !(best < INFINITY)) { best = ref_dist(qx, qy, qz, sx[0], sy[0], sz[0]); bi = 0; }
    if (k == 0) { best = 0.0f; bi = 0; }
    for (; k < m; ++k) {
        float d = ref_dist(qx, qy, qz, sx[k], sy[k], sz[k]);
        if (k == 0 || d < best) { best = d; bi = k; }
    }
    *out_d = best;
    *out_i = bi;
}

__attribute__((target("avx512f")))
static void nn_one_avx512(float qx, float qy, float qz, const float *sx, const float *sy,
                          const float *sz, int m, float *out_d, int *out_i) {
    const __m512 vqx = _mm512_set1_ps(qx), vqy = _mm512_set1_ps(qy), vqz = _mm512_set1_ps(qz);
    __m512 vbest = _mm512_set1_ps(INFINITY);
    __m512i vidx = _mm512_set1_epi32(0);
    __m512i vk = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m512i v16 = _mm512_set1_epi32(16);
    int k = 0;
    for (; k + 16 <= m; k += 16) {
        __m512 dx = _mm512_sub_ps(_mm512_loadu_ps(sx + k), vqx);
        __m512 dy = _mm512_sub_ps(_mm512_loadu_ps(sy + k), vqy);
        __m512 dz = _mm512_sub_ps(_mm512_loadu_ps(sz + k), vqz);
        __m512 d = _mm512_fmadd_ps(dz, dz, _mm512_fmadd_ps(dx, dx, _mm512_mul_ps(dy, dy)));
        __mmask16 lt = _mm512_cmp_ps_mask(d, vbest, _CMP_LT_OQ);
        vbest = _mm512_mask_mov_ps(vbest, lt, d);
        vidx = _mm512_mask_mov_epi32(vidx, lt, vk);
        vk = _mm512_add_epi32(vk, v16);
    }
    float bl[16];
    int il[16];
    _mm512_storeu_ps(bl, vbest);
    _mm512_storeu_si512((void *)il, vidx);
    float best = INFINITY;
    int bi = 0;
    for (int l = 0; l < 16; ++l)
        if (bl[l] < best || (bl[l] == best && il[l] < bi)) { best = bl[l]; bi = il[l]; }
    /* no distance compared below +inf (NaN or overflowing query): the reference's `k == 0 ||` clauses
     * (chamfer.cu:36,121,126) leave point 0 and its distance, as the scalar scan above does */
    if (k > 0 && !(best < INFINITY)) { best = ref_dist(qx, qy, qz, sx[0], sy[0], sz[0]); bi = 0; }
    if (k == 0) { best = 0.0f; bi = 0; }
    for (; k < m; ++k) {
        float d = ref_dist(qx, qy, qz, sx[k], sy[k], sz[k]);
        if (k == 0 || d < best) { best = d; bi = k; }
    }
    *out_d = best;
    *out_i = bi;
}
#endif

typedef void (*nn_one_fn)(float, float, float, const float *, const float *, const float *, int,
                          float *, int *);

static nn_one_fn pick_nn(int force_scalar) {
#if defined(__x86_64__)
    if (!force_scalar) {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f")) return nn_one_avx512;
        if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return nn_one_avx2;
    }
#endif
    (void)force_scalar;
    return nn_one_scalar;
}

/* simd level actually used: 0 scalar, 8 avx2, 16 avx512 */
int psi_oracle_simd_width(void) {
    nn_one_fn f = pick_nn(0);
#if defined(__x86_64__)
    if (f == nn_one_avx512) return 16;
    if (f == nn_one_avx2) return 8;
#endif
    (void)f;
    return 1;
}

/*
 * One direction of the reference forward: for every query q[b,j] the first-minimum
 * squared distance to s[b,:].  Batch strides are in floats; stride 0 shares a cloud
 * across the batch.  mode: 0 = vectorised (same results), 1 = force the scalar scan.
 */
int psi_oracle_nn_fwd(const float *q, long q_bstride, int B, int n, const float *s,
                      long s_bstride, int m, float *dist, int *idx, int mode) {
    if (B < 0 || n < 0 || m < 0) return 1;
    if (B == 0 || n == 0) return 0;
    if (m == 0) {                       /* no candidates: the kernel's loops never write */
        return 0;
    }
    nn_one_fn f = pick_nn(mode == 1);
    float *soa = (float *)malloc(sizeof(float) * 3 * (size_t)m);
    if (!soa) return 2;
    const float *prev = NULL;
    for (int b = 0; b < B; ++b) {
        const float *sb = s + (size_t)b * s_bstride;
        if (sb != prev) {
            for (int k = 0; k < m; ++k) {
                soa[k] = sb[3 * k];
                soa[m + k] = sb[3 * k + 1];
                soa[2 * (size_t)m + k] = sb[3 * k + 2];
            }
            prev = sb;
        }
        const float *qb = q + (size_t)b * q_bstride;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < n; ++j) {
            f(qb[3 * j], qb[3 * j + 1], qb[3 * j + 2], soa, soa + m, soa + 2 * (size_t)m, m,
              dist + (size_t)b * n + j, idx + (size_t)b * n + j);
        }
    }
    free(soa);
    return 0;
}

/* both directions, reference layout (chamfer.cu:136-154) */
int psi_oracle_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int n, int m,
                           float *dist1, int *idx1, float *dist2, int *idx2) {
    int rc = psi_oracle_nn_fwd(xyz1, (long)n * 3, B, n, xyz2, (long)m * 3, m, dist1, idx1, 0);
    if (rc) return rc;
    return psi_oracle_nn_fwd(xyz2, (long)m * 3, B, m, xyz1, (long)n * 3, n, dist2, idx2, 0);
}

/*
 * Backward (chamfer.cu:155-195).  ACCUMULATES into gxyz1/gxyz2 (the reference relies on
 * pre-zeroed buffers, chamfer.cu:177-178).  Serial accumulation order = ascending (b,j),
 * direction 1 then direction 2; the reference's atomics are order-free, so float sums
 * there are only reproducible up to rounding -- tests use a tolerance for gxyz2.
 */
int psi_oracle_chamfer_bwd(const float *xyz1, const float *xyz2, int B, int n, int m,
                           const float *gd1, const int *idx1, const float *gd2, const int *idx2,
                           float *gxyz1, float *gxyz2) {
    for (int dir = 0; dir < 2; ++dir) {
        const float *a = dir ? xyz2 : xyz1, *bb = dir ? xyz1 : xyz2;
        const float *gd = dir ? gd2 : gd1;
        const int *ix = dir ? idx2 : idx1;
        float *ga = dir ? gxyz2 : gxyz1, *gb = dir ? gxyz1 : gxyz2;
        int na = dir ? m : n, nb = dir ? n : m;
        if (!gd || !ix) continue;
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < na; ++j) {
                const float *p = a + ((size_t)b * na + j) * 3;
                int j2 = ix[(size_t)b * na + j];
                const float *qq = bb + ((size_t)b * nb + j2) * 3;
                float g = gd[(size_t)b * na + j] * 2.0f;
                for (int c = 0; c < 3; ++c) {
                    float t = g * (p[c] - qq[c]);
                    ga[((size_t)b * na + j) * 3 + c] += t;
                    gb[((size_t)b * nb + j2) * 3 + c] += -t;
                }
            }
    }
    return 0;
}

/*
 * SDF lookup with analytic gradient.  verts [n,3] in the scene frame, sdf grid [D,D,D]
 * indexed [x][y][z] (z contiguous).  Operation order of the reference:
 *   u = (v-min)/(max-min)*2-1      (fitting_habitat.py:148)
 *   f = ((u+1)/2)*(D-1)            (grid_sampler unnormalize, align_corners=True)
 *   f = clamp(f, 0, D-1)           (padding_mode='border'; gradient 0 outside)
 *   trilinear with weights (i1 - f) / (f - i0) as ATen's grid_sampler_3d does.
 * grad may be NULL.  out[n], grad[n,3] = d sdf / d v.
 */
int psi_oracle_sdf_fwd(const float *sdf, int D, const float *gmin, const float *gmax,
                       const float *verts, long n, float *out, float *grad) {
    if (D < 1) return 1;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        float f[3], gm[3];
        for (int a = 0; a < 3; ++a) {
            float v = verts[3 * i + a];
            float u = (v - gmin[a]) / (gmax[a] - gmin[a]) * 2.0f - 1.0f;
            float c = ((u + 1.0f) / 2.0f) * (float)(D - 1);
            float mult = (float)(D - 1) / 2.0f * (2.0f / (gmax[a] - gmin[a]));
            /* clip_coordinates_set_grad: x<=0 -> 0, grad 0 ; x>=D-1 -> D-1, grad 0 */
            if (c <= 0.0f) { c = 0.0f; mult = 0.0f; }
            else if (c >= (float)(D - 1)) { c = (float)(D - 1); mult = 0.0f; }
            f[a] = c;
            gm[a] = mult;
        }
        int i0[3];
        float w0[3], w1[3];
        for (int a = 0; a < 3; ++a) {
            float fl = floorf(f[a]);
            i0[a] = (int)fl;
            w1[a] = f[a] - fl;              /* weight of corner i0+1: (f - i0)   */
            w0[a] = (fl + 1.0f) - f[a];     /* weight of corner i0  : (i0+1 - f) */
        }
        float acc = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
        for (int cx = 0; cx < 2; ++cx)
            for (int cy = 0; cy < 2; ++cy)
                for (int cz = 0; cz < 2; ++cz) {
                    int ix = i0[0] + cx, iy = i0[1] + cy, iz = i0[2] + cz;
                    if (ix > D - 1 || iy > D - 1 || iz > D - 1) continue;   /* out of bounds: 0 */
                    float s = sdf[((size_t)ix * D + iy) * D + iz];
                    float wx = cx ? w1[0] : w0[0], wy = cy ? w1[1] : w0[1], wz = cz ? w1[2] : w0[2];
                    acc += s * ((wz * wy) * wx);      /* ATen order: (x_w*y_w)*z_w with its x = our z */
                    gx += (cx ? s : -s) * (wy * wz);
                    gy += (cy ? s : -s) * (wx * wz);
                    gz += (cz ? s : -s) * (wx * wy);
                }
        out[i] = acc;
        if (grad) {
            grad[3 * i + 0] = gx * gm[0];
            grad[3 * i + 1] = gy * gm[1];
            grad[3 * i + 2] = gz * gm[2];
        }
    }
    return 0;
}
