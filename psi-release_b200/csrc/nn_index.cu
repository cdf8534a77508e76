// Exact nearest neighbour against a STATIC scene through a two-level cluster index.
//
// Same contract as psi_nn_fwd with a shared scene (chamfer_pytorch/chamfer.cu:12-134 semantics:
// d = fma(dz,dz,fma(dx,dx,rn(dy*dy))), lowest ORIGINAL index wins ties) -- the outputs are
// bit-identical to the brute-force kernel; only provably irrelevant pair evaluations are skipped.
//
// Why it is exact.  Scene points are sorted along a Morton curve and cut into clusters of 32 with an
// axis-aligned box [lo,hi]; 32 clusters form a super-cluster with its own box.  For a query q and a
// box, g_a = max(fl(lo_a-q_a), fl(q_a-hi_a), 0) satisfies g_a <= |fl(s_a-q_a)| for every point s in the
// box (rounding is monotone), and fma(gz,gz,fma(gx,gx,rn(gy*gy))) is monotone in |g|, so
// lb(q,box) <= d(q,s) IN FLOATING POINT for every s in the box.  A cluster is skipped only if
// lb > ub where ub is a distance already seen; so every skipped point has d > final minimum and cannot
// win or tie.  Visited points are compared lexicographically on (d, original index).
//
// Execution: one warp per query.  Lanes evaluate 32 boxes (or the 32 points of a cluster, one
// coalesced 512-byte load) in parallel; the running upper bound is a warp-wide `redux.sync.min` on
// the distance bits (d >= 0, so the IEEE bit pattern is monotone); the best-first seed (nearest
// super-cluster, then its nearest cluster) makes the bound tight before the ordered sweep.
#include "common.cuh"
#include <algorithm>
#include <math.h>
#include <math_constants.h>
#include <new>
#include <vector>

struct psi_nn_index {
    int m, num_clusters, num_supers, spad;  // spad: supers rounded up to 32
    float4 *pts;    // [num_supers*32*32] (x,y,z,orig index bits); pads = +inf / INT_MAX
    float4 *cbox;   // [num_supers*32][2] lo,hi ; pad clusters lo=hi=+inf
    float4 *sbox;   // [spad][2]
    size_t bytes;
};

namespace psi {

__device__ __forceinline__ float box_lb(const float4 lo, const float4 hi, float qx, float qy, float qz) {
    const float gx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
    return __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
}

struct LaneBest {
    float d;
    int i;
};

// all lanes: evaluate the 32 points of cluster c, fold into the lane-local best, return new ub bits
__device__ __forceinline__ unsigned visit_cluster(const float4 *__restrict__ pts, int c, int lane,
                                                  float qx, float qy, float qz, LaneBest &lb,
                                                  unsigned ub_bits) {
    const float4 p = __ldg(pts + (size_t)c * 32 + lane);
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
    const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const int oi = __float_as_int(p.w);
    if (d < lb.d || (d == lb.d && oi < lb.i)) {
        lb.d = d;
        lb.i = oi;
    }
    const unsigned m = __reduce_min_sync(0xffffffffu, __float_as_uint(d));
    return min(ub_bits, m);
}

// all lanes: sweep the clusters of super-cluster s whose bound admits them; `first` (or -1) is
// visited before the ordered sweep (best-first seed)
__device__ __forceinline__ unsigned visit_super(const psi_nn_index ix, int s, int lane, float qx,
                                                float qy, float qz, LaneBest &lb, unsigned ub_bits,
                                                bool seed) {
    const int c = s * 32 + lane;
    const float4 lo = __ldg(ix.cbox + (size_t)c * 2), hi = __ldg(ix.cbox + (size_t)c * 2 + 1);
    const unsigned clb = __float_as_uint(box_lb(lo, hi, qx, qy, qz));
    int skip = -1;
    if (seed) {
        const unsigned mn = __reduce_min_sync(0xffffffffu, clb);
        skip = __ffs(__ballot_sync(0xffffffffu, clb == mn)) - 1;
        ub_bits = visit_cluster(ix.pts, s * 32 + skip, lane, qx, qy, qz, lb, ub_bits);
    }
    unsigned mask = __ballot_sync(0xffffffffu, clb <= ub_bits && lane != skip);
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        ub_bits = visit_cluster(ix.pts, s * 32 + k, lane, qx, qy, qz, lb, ub_bits);
        mask &= __ballot_sync(0xffffffffu, clb <= ub_bits);   // the bound may have tightened
    }
    return ub_bits;
}

__global__ void __launch_bounds__(128)
nn_index_query_kernel(const psi_nn_index ix, const float *__restrict__ q, long q_bstride, int n,
                      long total, float *__restrict__ dist, int *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const long warps = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < total; t += warps) {
        const long b = t / n;
        const float *qp = q + b * q_bstride + (t - b * n) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        LaneBest lb;
        lb.d = CUDART_INF_F;
        lb.i = 0x7fffffff;
        unsigned ub = 0x7f800000u;  // +inf
        // seed: nearest super-cluster by box bound (lowest index on ties)
        unsigned best_lb = 0xffffffffu;
        int s0 = 0;
        for (int r = 0; r < ix.spad; r += 32) {
            const float4 lo = __ldg(ix.sbox + (size_t)(r + lane) * 2), hi = __ldg(ix.sbox + (size_t)(r + lane) * 2 + 1);
            const unsigned slb = __float_as_uint(box_lb(lo, hi, qx, qy, qz));
            const unsigned mn = __reduce_min_sync(0xffffffffu, slb);
            if (mn < best_lb) {
                best_lb = mn;
                s0 = r + __ffs(__ballot_sync(0xffffffffu, slb == mn)) - 1;
            }
        }
        ub = visit_super(ix, s0, lane, qx, qy, qz, lb, ub, true);
        // ordered sweep over the remaining super-clusters
        for (int r = 0; r < ix.spad; r += 32) {
            const float4 lo = __ldg(ix.sbox + (size_t)(r + lane) * 2), hi = __ldg(ix.sbox + (size_t)(r + lane) * 2 + 1);
            const unsigned slb = __float_as_uint(box_lb(lo, hi, qx, qy, qz));
            unsigned mask = __ballot_sync(0xffffffffu, slb <= ub && (r + lane) != s0 && (r + lane) < ix.num_supers);
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1;
                ub = visit_super(ix, r + k, lane, qx, qy, qz, lb, ub, false);
                mask &= __ballot_sync(0xffffffffu, slb <= ub);
            }
        }
        // lexicographic (d, original index) minimum over the lanes
        const unsigned dmin = __reduce_min_sync(0xffffffffu, __float_as_uint(lb.d));
        const unsigned imin = __reduce_min_sync(
            0xffffffffu, (__float_as_uint(lb.d) == dmin) ? (unsigned)lb.i : 0x7fffffffu);
        if (lane == 0) {
            dist[t] = __uint_as_float(dmin);
            if (idx) idx[t] = (int)imin;
        }
    }
}

static inline unsigned spread10(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

}  // namespace psi

extern "C" {

void psi_nn_index_destroy(psi_nn_index *ix) {
    if (!ix) return;
    cudaFree(ix->pts);
    cudaFree(ix->cbox);
    cudaFree(ix->sbox);
    delete ix;
}

int psi_nn_index_create(psi_nn_index **out, const float *h_points, int m, psi_stream_t stream) {
    using namespace psi;
    if (!out || !h_points || m < 1) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    float lo[3] = {HUGE_VALF, HUGE_VALF, HUGE_VALF}, hi[3] = {-HUGE_VALF, -HUGE_VALF, -HUGE_VALF};
    for (int i = 0; i < m; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = h_points[(size_t)i * 3 + a];
            if (!(v == v) || v == HUGE_VALF || v == -HUGE_VALF) return PSI_ERR_BAD_ARG;  // finite only
            lo[a] = v < lo[a] ? v : lo[a];
            hi[a] = v > hi[a] ? v : hi[a];
        }
    std::vector<std::pair<unsigned, int>> order((size_t)m);
    for (int i = 0; i < m; ++i) {
        unsigned code = 0;
        for (int a = 0; a < 3; ++a) {
            const double ext = (double)hi[a] - (double)lo[a];
            double u = ext > 0 ? ((double)h_points[(size_t)i * 3 + a] - lo[a]) / ext : 0.0;
            unsigned cell = (unsigned)(u * 1023.0 + 0.5);
            code |= spread10(cell > 1023u ? 1023u : cell) << a;
        }
        order[i] = std::make_pair(code, i);
    }
    std::sort(order.begin(), order.end());
    psi_nn_index *ix = new (std::nothrow) psi_nn_index();
    if (!ix) return PSI_ERR_ALLOC;
    ix->m = m;
    ix->num_clusters = (m + 31) / 32;
    ix->num_supers = (ix->num_clusters + 31) / 32;
    ix->spad = ((ix->num_supers + 31) / 32) * 32;
    const size_t ncl = (size_t)ix->num_supers * 32;
    const float inf = HUGE_VALF;
    std::vector<float4> pts(ncl * 32), cbox(ncl * 2), sbox((size_t)ix->spad * 2);
    for (size_t i = 0; i < pts.size(); ++i) {
        if (i < (size_t)m) {
            const int oi = order[i].second;
            pts[i] = make_float4(h_points[(size_t)oi * 3], h_points[(size_t)oi * 3 + 1], h_points[(size_t)oi * 3 + 2], 0.f);
            reinterpret_cast<int &>(pts[i].w) = oi;
        } else {
            pts[i] = make_float4(inf, inf, inf, 0.f);
            reinterpret_cast<int &>(pts[i].w) = 0x7fffffff;
        }
    }
    auto grow = [](float4 &l, float4 &h, const float4 &a, const float4 &b) {
        l.x = a.x < l.x ? a.x : l.x; l.y = a.y < l.y ? a.y : l.y; l.z = a.z < l.z ? a.z : l.z;
        h.x = b.x > h.x ? b.x : h.x; h.y = b.y > h.y ? b.y : h.y; h.z = b.z > h.z ? b.z : h.z;
    };
    for (size_t c = 0; c < ncl; ++c) {
        float4 l = make_float4(inf, inf, inf, 0.f), h = make_float4(-inf, -inf, -inf, 0.f);
        bool any = false;
        for (int k = 0; k < 32; ++k) {
            const size_t i = c * 32 + k;
            if (i < (size_t)m) { grow(l, h, pts[i], pts[i]); any = true; }
        }
        if (!any) { l = make_float4(inf, inf, inf, 0.f); h = l; }
        cbox[c * 2] = l;
        cbox[c * 2 + 1] = h;
    }
    for (int s = 0; s < ix->spad; ++s) {
        float4 l = make_float4(inf, inf, inf, 0.f), h = make_float4(-inf, -inf, -inf, 0.f);
        bool any = false;
        if (s < ix->num_supers)
            for (int k = 0; k < 32; ++k) {
                const size_t c = (size_t)s * 32 + k;
                if (c < (size_t)ix->num_clusters) { grow(l, h, cbox[c * 2], cbox[c * 2 + 1]); any = true; }
            }
        if (!any) { l = make_float4(inf, inf, inf, 0.f); h = l; }
        sbox[(size_t)s * 2] = l;
        sbox[(size_t)s * 2 + 1] = h;
    }
    ix->bytes = 0;
    auto up = [&](float4 **dst, const std::vector<float4> &h) -> int {
        const size_t nb = h.size() * sizeof(float4);
        if (cudaMalloc((void **)dst, nb) != cudaSuccess) return PSI_ERR_ALLOC;
        cudaError_t e = cudaMemcpyAsync(*dst, h.data(), nb, cudaMemcpyHostToDevice, st);
        ix->bytes += nb;
        return e == cudaSuccess ? PSI_OK : (int)e;
    };
    int rc = up(&ix->pts, pts);
    if (rc == PSI_OK) rc = up(&ix->cbox, cbox);
    if (rc == PSI_OK) rc = up(&ix->sbox, sbox);
    if (rc == PSI_OK) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = (int)e;
    }
    if (rc != PSI_OK) {
        psi_nn_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return PSI_OK;
}

size_t psi_nn_index_bytes(const psi_nn_index *ix) { return ix ? ix->bytes : 0; }

int psi_nn_index_query(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                       float *dist, int *idx, psi_stream_t stream) {
    if (!ix || B < 0 || n < 0) return PSI_ERR_BAD_ARG;
    if (B == 0 || n == 0) return PSI_OK;
    if (!q || !dist) return PSI_ERR_BAD_ARG;
    const long total = (long)B * n;
    long blocks = (total + 3) / 4;
    const long cap = (long)PSI_NUM_SMS * 16;
    if (blocks > cap) blocks = cap;
    psi::nn_index_query_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*ix, q, q_bstride, n,
                                                                                  total, dist, idx);
    PSI_LAUNCHED();
    return PSI_OK;
}

}  // extern "C"
