// Exact nearest neighbour against a STATIC scene through a two-level box index.
//
// Same contract as psi_nn_fwd with a shared scene (chamfer_pytorch/chamfer.cu:12-134 semantics:
// d = fma(dz,dz,fma(dx,dx,rn(dy*dy))), lowest ORIGINAL index wins ties) -- the outputs are
// bit-identical to the brute-force kernel; only provably irrelevant pair evaluations are skipped.
//
// Index (built once per scene on the host): a balanced kd bisection (longest axis, sizes kept
// multiples of the leaf) orders the points so that every 32 consecutive points form a compact
// CLUSTER and every 8 consecutive clusters a compact SUPER-cluster, each with an axis-aligned box.
//
// Why it is exact.  For a query q and a box [lo,hi], g_a = max(fl(lo_a-q_a), fl(q_a-hi_a), 0)
// satisfies g_a <= |fl(s_a-q_a)| for every point s in the box (rounding is monotone), and
// fma(gz,gz,fma(gx,gx,rn(gy*gy))) is monotone in |g|, so lb(q,box) <= d(q,s) IN FLOATING POINT for
// every s in the box.  A box is skipped only if lb > ub where ub is a distance already seen, so every
// skipped point has d > final minimum and cannot win or tie.  Visited points are compared
// lexicographically on (d, original index).
//
// Execution: one warp per query, boxes resident in shared memory (6 kB + 50 kB at 50 000 points).
// Lanes evaluate 32 super boxes per round (bounds kept in registers), then for every admitted
// super its 8 cluster boxes, then the 32 points of an admitted cluster with ONE coalesced 512-byte
// load; the running upper bound is a warp-wide redux.sync.min on the distance bits (d >= 0, so the
// IEEE pattern is monotone).  A best-first seed (nearest super, its nearest cluster) tightens the
// bound before the ordered sweep.
#include "common.cuh"
#include <algorithm>
#include <math.h>
#include <math_constants.h>
#include <new>
#include <vector>

namespace psi {
constexpr int kLeaf = 32;        // points per cluster (one per lane)
constexpr int kFan = 8;          // clusters per super-cluster
constexpr int kMaxRounds = 16;   // super bounds kept in registers: up to 16*32 supers = 131 072 points
constexpr int kIdxThreads = 512;
}  // namespace psi

struct psi_nn_index {
    int m, num_clusters, num_supers, spad, rounds;
    float4 *pts;    // [num_supers*8*32] (x,y,z,orig index bits); pads = +inf / INT_MAX
    float4 *cbox;   // [2][num_supers*8]: every lo, then every hi ; pad clusters lo=hi=+inf
    float4 *sbox;   // [2][spad], spad = rounds*32
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

// all lanes: the 32 points of cluster c -> lane-local best; returns the tightened bound
__device__ __forceinline__ unsigned visit_cluster(const float4 *__restrict__ pts, int c, int lane,
                                                  float qx, float qy, float qz, LaneBest &lb,
                                                  unsigned ub_bits) {
    const float4 p = __ldg(pts + (size_t)c * kLeaf + lane);
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
    const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const int oi = __float_as_int(p.w);
    if (d < lb.d || (d == lb.d && oi < lb.i)) {
        lb.d = d;
        lb.i = oi;
    }
    return min(ub_bits, __reduce_min_sync(0xffffffffu, __float_as_uint(d)));
}

// all lanes: the clusters of super s that the bound admits (lanes 0..7 hold one cluster box each).
// seed: visit the nearest cluster first.  `done` = a cluster already visited (or -1).
template <bool SMEM>
__device__ __forceinline__ unsigned visit_super(const psi_nn_index &ix, const float4 *cbox, int s,
                                                int lane, float qx, float qy, float qz, LaneBest &lb,
                                                unsigned ub_bits, bool seed) {
    unsigned clb = 0x7f800000u;
    if (lane < kFan) {
        const int c = s * kFan + lane;
        const float4 lo = SMEM ? cbox[c] : __ldg(cbox + c);                       // SoA: all lo, then all hi
        const float4 hi = SMEM ? cbox[ix.num_clusters + c] : __ldg(cbox + ix.num_clusters + c);
        clb = __float_as_uint(box_lb(lo, hi, qx, qy, qz));
    }
    int skip = -1;
    if (seed) {
        const unsigned mn = __reduce_min_sync(0xffffffffu, clb);
        skip = __ffs(__ballot_sync(0xffffffffu, clb == mn)) - 1;
        ub_bits = visit_cluster(ix.pts, s * kFan + skip, lane, qx, qy, qz, lb, ub_bits);
    }
    unsigned mask = __ballot_sync(0xffffffffu, clb <= ub_bits && lane != skip && lane < kFan);
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        ub_bits = visit_cluster(ix.pts, s * kFan + k, lane, qx, qy, qz, lb, ub_bits);
        mask &= __ballot_sync(0xffffffffu, clb <= ub_bits);   // the bound may have tightened
    }
    return ub_bits;
}

template <bool SMEM>
__global__ void __launch_bounds__(kIdxThreads)
nn_index_query_kernel(const psi_nn_index ix, const float *__restrict__ q, long q_bstride, int n,
                      const int *__restrict__ qsel, long total, float *__restrict__ dist,
                      int *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4 *sbox = ix.sbox, *cbox = ix.cbox;
    if (SMEM) {
        float4 *s4 = reinterpret_cast<float4 *>(smem_raw);
        const int ns = ix.spad * 2, nc = ix.num_supers * kFan * 2;
        for (int i = threadIdx.x; i < ns; i += blockDim.x) s4[i] = __ldg(ix.sbox + i);
        for (int i = threadIdx.x; i < nc; i += blockDim.x) s4[ns + i] = __ldg(ix.cbox + i);
        __syncthreads();
        sbox = s4;
        cbox = s4 + ns;
    }
    const int lane = threadIdx.x & 31;
    const long warps = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < total; t += warps) {
        const long b = t / n;
        const long j = t - b * n;
        const float *qp = q + b * q_bstride + (qsel ? (long)__ldg(qsel + j) : j) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        LaneBest lb;
        lb.d = CUDART_INF_F;
        lb.i = 0x7fffffff;
        // level 1: every super-cluster bound, 32 per round, kept in registers
        unsigned slb[kMaxRounds];
        unsigned best_lb = 0xffffffffu;
        int s0 = 0;
#pragma unroll
        for (int r = 0; r < kMaxRounds; ++r) {
            slb[r] = 0x7f800000u;
            if (r < ix.rounds) {
                const int s = r * 32 + lane;
                const float4 lo = SMEM ? sbox[s] : __ldg(sbox + s);
                const float4 hi = SMEM ? sbox[ix.spad + s] : __ldg(sbox + ix.spad + s);
                slb[r] = __float_as_uint(box_lb(lo, hi, qx, qy, qz));
                const unsigned mn = __reduce_min_sync(0xffffffffu, slb[r]);
                if (mn < best_lb) {
                    best_lb = mn;
                    s0 = r * 32 + __ffs(__ballot_sync(0xffffffffu, slb[r] == mn)) - 1;
                }
            }
        }
        // seed, then the ordered sweep over the admitted supers
        unsigned ub = visit_super<SMEM>(ix, cbox, s0, lane, qx, qy, qz, lb, 0x7f800000u, true);
#pragma unroll
        for (int r = 0; r < kMaxRounds; ++r) {
            if (r < ix.rounds) {
                unsigned mask = __ballot_sync(0xffffffffu, slb[r] <= ub && (r * 32 + lane) != s0);
                while (mask) {
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    ub = visit_super<SMEM>(ix, cbox, r * 32 + k, lane, qx, qy, qz, lb, ub, false);
                    mask &= __ballot_sync(0xffffffffu, slb[r] <= ub);
                }
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

// balanced kd bisection: reorder ids[0..n) so that consecutive runs of `leaf` are compact
static void kd_order(const float *P, int *ids, int n, int leaf) {
    if (n <= leaf) return;
    float lo[3] = {HUGE_VALF, HUGE_VALF, HUGE_VALF}, hi[3] = {-HUGE_VALF, -HUGE_VALF, -HUGE_VALF};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = P[(size_t)ids[i] * 3 + a];
            lo[a] = v < lo[a] ? v : lo[a];
            hi[a] = v > hi[a] ? v : hi[a];
        }
    int ax = 0;
    if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1;
    if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
    const int nl = ((n / leaf + 1) / 2) * leaf;   // left part: a multiple of the leaf size
    std::nth_element(ids, ids + nl, ids + n, [&](int a, int b) {
        const float va = P[(size_t)a * 3 + ax], vb = P[(size_t)b * 3 + ax];
        return va < vb || (va == vb && a < b);
    });
    kd_order(P, ids, nl, leaf);
    kd_order(P, ids + nl, n - nl, leaf);
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
    for (size_t i = 0; i < (size_t)m * 3; ++i) {
        const float v = h_points[i];
        if (!(v == v) || v == HUGE_VALF || v == -HUGE_VALF) return PSI_ERR_BAD_ARG;  // finite only
    }
    const int num_supers = (m + kLeaf * kFan - 1) / (kLeaf * kFan);
    const int rounds = (num_supers + 31) / 32;
    if (rounds > kMaxRounds) return PSI_ERR_UNSUPPORTED;
    std::vector<int> ids((size_t)m);
    for (int i = 0; i < m; ++i) ids[i] = i;
    kd_order(h_points, ids.data(), m, kLeaf * kFan);                       // supers
    for (int s = 0; s < num_supers; ++s) {                                  // clusters inside a super
        const int b = s * kLeaf * kFan, e = std::min(m, b + kLeaf * kFan);
        kd_order(h_points, ids.data() + b, e - b, kLeaf);
    }
    psi_nn_index *ix = new (std::nothrow) psi_nn_index();
    if (!ix) return PSI_ERR_ALLOC;
    ix->m = m;
    ix->num_supers = num_supers;
    ix->num_clusters = num_supers * kFan;
    ix->rounds = rounds;
    ix->spad = rounds * 32;
    const size_t ncl = (size_t)ix->num_clusters;
    const float inf = HUGE_VALF;
    std::vector<float4> pts(ncl * kLeaf), cbox(ncl * 2), sbox((size_t)ix->spad * 2);
    for (size_t i = 0; i < pts.size(); ++i) {
        if (i < (size_t)m) {
            const int oi = ids[i];
            pts[i] = make_float4(h_points[(size_t)oi * 3], h_points[(size_t)oi * 3 + 1], h_points[(size_t)oi * 3 + 2], 0.f);
            reinterpret_cast<int &>(pts[i].w) = oi;
        } else {
            pts[i] = make_float4(inf, inf, inf, 0.f);
            reinterpret_cast<int &>(pts[i].w) = 0x7fffffff;
        }
    }
    auto box_of = [&](size_t first, size_t count, float4 &l, float4 &h) {
        l = make_float4(inf, inf, inf, 0.f);
        h = make_float4(-inf, -inf, -inf, 0.f);
        bool any = false;
        for (size_t i = first; i < first + count && i < (size_t)m; ++i) {
            const float4 &p = pts[i];
            l.x = p.x < l.x ? p.x : l.x; l.y = p.y < l.y ? p.y : l.y; l.z = p.z < l.z ? p.z : l.z;
            h.x = p.x > h.x ? p.x : h.x; h.y = p.y > h.y ? p.y : h.y; h.z = p.z > h.z ? p.z : h.z;
            any = true;
        }
        if (!any) { l = make_float4(inf, inf, inf, 0.f); h = l; }   // empty: bound = +inf, never admitted
    };
    for (size_t c = 0; c < ncl; ++c) box_of(c * kLeaf, kLeaf, cbox[c], cbox[ncl + c]);
    for (int s = 0; s < ix->spad; ++s)
        box_of((size_t)s * kLeaf * kFan, s < num_supers ? (size_t)kLeaf * kFan : 0, sbox[(size_t)s], sbox[(size_t)ix->spad + s]);
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

int psi_nn_index_query(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n, const int *qsel,
                       float *dist, int *idx, psi_stream_t stream) {
    using namespace psi;
    if (!ix || B < 0 || n < 0) return PSI_ERR_BAD_ARG;
    if (B == 0 || n == 0) return PSI_OK;
    if (!q || !dist) return PSI_ERR_BAD_ARG;
    const long total = (long)B * n;
    const int wpb = kIdxThreads / 32;
    const size_t box_bytes = ((size_t)ix->spad * 2 + (size_t)ix->num_clusters * 2) * sizeof(float4);
    const bool smem = box_bytes <= 56 * 1024;     // 4 CTAs of 16 warps per SM stay resident
    long blocks = (total + wpb - 1) / wpb;
    const long cap = (long)PSI_NUM_SMS * 4;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(nn_index_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
            attr = true;
        }
        nn_index_query_kernel<true><<<(unsigned)blocks, kIdxThreads, box_bytes, st>>>(*ix, q, q_bstride, n, qsel, total, dist, idx);
    } else {
        nn_index_query_kernel<false><<<(unsigned)blocks, kIdxThreads, 0, st>>>(*ix, q, q_bstride, n, qsel, total, dist, idx);
    }
    PSI_LAUNCHED();
    return PSI_OK;
}

}  // extern "C"
