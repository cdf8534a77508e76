// Exact nearest neighbour against a STATIC scene through an 8-ary box tree.
//
// Same contract as psi_nn_fwd with a shared scene (chamfer_pytorch/chamfer.cu:12-134 semantics:
// d = fma(dz,dz,fma(dx,dx,rn(dy*dy))), lowest ORIGINAL index wins ties) -- the outputs are
// bit-identical to the brute-force kernel; only provably irrelevant pair evaluations are skipped.
//
// Index (built once per scene on the host): a balanced kd bisection (longest axis, part sizes kept
// multiples of the level below) orders the points so that 32 consecutive points form a compact
// leaf CLUSTER and every 8 consecutive nodes of a level form their parent: level k node i covers
// points [i*32*8^(L-k), ...), L levels below the root (L = 4 for 50 000 points: 4 / 25 / 196 /
// 1563 nodes).  Every node has an axis-aligned box; boxes are stored SoA (all lo, then all hi)
// and live in shared memory (57 kB at 50 000 points).
//
// Why it is exact.  For a query q and a box [lo,hi], g_a = max(fl(lo_a-q_a), fl(q_a-hi_a), 0)
// satisfies g_a <= |fl(s_a-q_a)| for every point s in the box (rounding is monotone), and
// fma(gz,gz,fma(gx,gx,rn(gy*gy))) is monotone in |g|, so lb(q,box) <= d(q,s) IN FLOATING POINT for
// every s in the box.  A node is skipped only if lb > ub where ub is the distance of a point
// already evaluated, so every skipped point has d > final minimum and cannot win or tie.
// Visited points are compared lexicographically on (d, original index).
//
// Execution: EIGHT lanes per query (4 queries per warp) -- the tree's fan-out, so every lane
// evaluates a live child box at every expansion.  A group walks the tree depth-first with a small
// stack in shared memory (entries carry their bound and are re-checked when popped); a leaf is 4
// coalesced 128-byte loads per group, 4 points per lane; the running bound is a group-wide
// redux.sync.min on the distance bits (d >= 0: the IEEE pattern is monotone).  The bound is seeded
// either by a greedy descent (nearest child at every level) or from a HINT: the cluster that held
// the same query's nearest neighbour in the previous fitting iteration (bodies move little per
// Adam step) -- just another visited leaf, so exactness does not depend on it.
#include "common.cuh"
#include <algorithm>
#include <math.h>
#include <math_constants.h>
#include <new>
#include <vector>

namespace psi {
constexpr int kLeaf = 32;        // points per leaf cluster
constexpr int kFan = 8;          // children per node = lanes per query
constexpr int kMaxLev = 6;       // 32 * 8^6 = 8.4 M points
constexpr int kIdxThreads = 512;
constexpr int kGroupsPerCta = kIdxThreads / kFan;
constexpr int kStackDepth = 7 * kMaxLev + 2;
constexpr size_t kIdxSmemMax = 100 * 1024;
}  // namespace psi

struct psi_nn_index {
    int m, nlev, nbox;
    int count[psi::kMaxLev + 1];   // nodes at level k (1..nlev); level nlev = leaf clusters
    int off[psi::kMaxLev + 1];     // first box of level k
    float4 *pts;    // [count[nlev]*32] (x,y,z,orig index bits); pads = +inf / INT_MAX
    float4 *boxes;  // SoA [lo: nbox][hi: nbox]
    int *pos_of;    // sorted position of every original point (index -> leaf, for the hint)
    size_t bytes;
};

namespace psi {

__device__ __forceinline__ unsigned box_lb(const float4 lo, const float4 hi, float qx, float qy, float qz) {
    const float gx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
    return __float_as_uint(__fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy))));
}

struct Query {
    float x, y, z;
    float bd;        // lane-local best distance
    int bi;          // ... and its original index
    unsigned ub;     // group-uniform bound (bits of a distance already seen)
    unsigned gmask;  // the 8 lanes of this query
    int gl;          // lane within the group
};

// bound of child `gl` of node `parent` (a level k-1 node; parent 0 at k == 1 is the root)
template <bool SMEM>
__device__ __forceinline__ unsigned child_lb(const psi_nn_index &ix, const int *lev, const float4 *boxes,
                                             int k, int parent, const Query &q) {
    const int child = parent * kFan + q.gl;
    if (child >= lev[k]) return 0x7f800000u;                 // lev[k] = count, lev[8+k] = first box
    const int node = lev[8 + k] + child;
    const float4 lo = SMEM ? boxes[node] : __ldg(boxes + node);
    const float4 hi = SMEM ? boxes[ix.nbox + node] : __ldg(boxes + ix.nbox + node);
    return box_lb(lo, hi, q.x, q.y, q.z);
}

// the 8 lanes of the group: 4 points each of leaf c
__device__ __forceinline__ void visit_leaf(const float4 *__restrict__ pts, int c, Query &q) {
    const float4 *base = pts + (size_t)c * kLeaf + q.gl;
    float dmin = CUDART_INF_F;
#pragma unroll
    for (int p = 0; p < kLeaf / kFan; ++p) {
        const float4 pt = __ldg(base + p * kFan);
        const float dx = __fsub_rn(pt.x, q.x), dy = __fsub_rn(pt.y, q.y), dz = __fsub_rn(pt.z, q.z);
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        const int oi = __float_as_int(pt.w);
        if (d < q.bd || (d == q.bd && oi < q.bi)) {
            q.bd = d;
            q.bi = oi;
        }
        dmin = fminf(dmin, d);
    }
    q.ub = min(q.ub, __reduce_min_sync(q.gmask, __float_as_uint(dmin)));
}

template <bool SMEM>
__global__ void __launch_bounds__(kIdxThreads)
nn_tree_query_kernel(const psi_nn_index ix, const float *__restrict__ q_in, long q_bstride, int n,
                     const int *__restrict__ qsel, long total, float *__restrict__ dist,
                     int *__restrict__ idx, int *__restrict__ hint) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int lev[16];
    if (threadIdx.x < 16) lev[threadIdx.x] = threadIdx.x < 8 ? ix.count[threadIdx.x & 7] : ix.off[threadIdx.x & 7];
    __syncthreads();
    uint2 *stacks = reinterpret_cast<uint2 *>(smem_raw);                  // [groups][kStackDepth]
    const float4 *boxes = ix.boxes;
    if (SMEM) {
        float4 *s4 = reinterpret_cast<float4 *>(smem_raw + (size_t)kGroupsPerCta * kStackDepth * sizeof(uint2));
        for (int i = threadIdx.x; i < 2 * ix.nbox; i += blockDim.x) s4[i] = __ldg(ix.boxes + i);
        __syncthreads();
        boxes = s4;
    }
    const int lane = threadIdx.x & 31;
    Query q;
    q.gl = lane & 7;
    const int gshift = lane & 24;
    q.gmask = 0xffu << gshift;
    uint2 *stk = stacks + (size_t)(threadIdx.x >> 3) * kStackDepth;
    const long ngroups = (long)gridDim.x * kGroupsPerCta;
    const int L = ix.nlev;
    for (long t = (long)blockIdx.x * kGroupsPerCta + (threadIdx.x >> 3); t < total; t += ngroups) {
        const long b = t / n;
        const long j = t - b * n;
        const float *qp = q_in + b * q_bstride + (qsel ? (long)__ldg(qsel + j) : j) * 3;
        q.x = __ldg(qp); q.y = __ldg(qp + 1); q.z = __ldg(qp + 2);
        q.bd = CUDART_INF_F;
        q.bi = 0x7fffffff;
        q.ub = 0x7f800000u;
        // ---- seed the bound: hinted leaf, else greedy descent to the nearest leaf
        int seeded = hint ? hint[t] : -1;
        if (seeded < 0 || seeded >= lev[L]) {
            int node = 0;
            for (int k = 1; k <= L; ++k) {
                const unsigned lb = child_lb<SMEM>(ix, lev, boxes, k, node, q);
                const unsigned mn = __reduce_min_sync(q.gmask, lb);
                const unsigned eq = (__ballot_sync(q.gmask, lb == mn) >> gshift) & 0xffu;
                node = node * kFan + (__ffs(eq) - 1);
            }
            seeded = node;
        }
        visit_leaf(ix.pts, seeded, q);
        // ---- depth-first sweep; stack entry = (level << 28 | node, bound bits)
        int sp = 0;
        int level = 0, node = 0;          // the root (level 0) is expanded first
        bool expand = true;
        while (true) {
            if (expand) {
                const int k = level + 1;
                const unsigned lb = child_lb<SMEM>(ix, lev, boxes, k, node, q);
                const bool adm = lb <= q.ub;
                const unsigned m8 = (__ballot_sync(q.gmask, adm) >> gshift) & 0xffu;
                if (adm) {   // lowest child index ends on top of the stack
                    const int above = __popc(m8 & ((1u << q.gl) - 1u));
                    stk[sp + __popc(m8) - 1 - above] = make_uint2(((unsigned)k << 28) | (unsigned)(node * kFan + q.gl), lb);
                }
                sp += __popc(m8);
                __syncwarp(q.gmask);
            }
            if (sp == 0) break;
            --sp;
            const uint2 e = stk[sp];
            __syncwarp(q.gmask);          // everyone has read the entry before it can be overwritten
            expand = false;
            if (e.y > q.ub) continue;     // the bound tightened since this entry was pushed
            level = (int)(e.x >> 28);
            node = (int)(e.x & 0x0fffffffu);
            if (level == L) {
                if (node != seeded) visit_leaf(ix.pts, node, q);
            } else {
                expand = true;
            }
        }
        // ---- lexicographic (d, original index) minimum over the group
        const unsigned dmin = __reduce_min_sync(q.gmask, __float_as_uint(q.bd));
        const bool win = __float_as_uint(q.bd) == dmin;
        const unsigned imin = __reduce_min_sync(q.gmask, win ? (unsigned)q.bi : 0x7fffffffu);
        if (q.gl == 0) {
            dist[t] = __uint_as_float(dmin);
            if (idx) idx[t] = (int)imin;
            if (hint) hint[t] = (imin < (unsigned)ix.m) ? __ldg(ix.pos_of + imin) / kLeaf : -1;
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
    cudaFree(ix->boxes);
    cudaFree(ix->pos_of);
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
    int L = 1;
    long span = (long)kLeaf * kFan;           // points under the root with L levels
    while (span < m) { span *= kFan; ++L; }
    if (L > kMaxLev) return PSI_ERR_UNSUPPORTED;
    psi_nn_index *ix = new (std::nothrow) psi_nn_index();
    if (!ix) return PSI_ERR_ALLOC;
    ix->pts = nullptr; ix->boxes = nullptr; ix->pos_of = nullptr;
    ix->m = m;
    ix->nlev = L;
    // level k node size = 32 * 8^(L-k)
    std::vector<long> size(L + 1, 0);
    size[L] = kLeaf;
    for (int k = L - 1; k >= 1; --k) size[k] = size[k + 1] * kFan;
    int nbox = 0;
    for (int k = 0; k <= kMaxLev; ++k) { ix->count[k] = 0; ix->off[k] = 0; }
    for (int k = 1; k <= L; ++k) {
        ix->count[k] = (int)((m + size[k] - 1) / size[k]);
        ix->off[k] = nbox;
        nbox += ix->count[k];
    }
    ix->nbox = nbox;
    std::vector<int> ids((size_t)m);
    for (int i = 0; i < m; ++i) ids[i] = i;
    for (int k = 1; k <= L; ++k) {            // refine level by level inside every parent range
        const long parent = k == 1 ? (long)m : size[k - 1];
        for (long b = 0; b < m; b += parent)
            kd_order(h_points, ids.data() + b, (int)std::min<long>(parent, m - b), (int)size[k]);
    }
    const float inf = HUGE_VALF;
    std::vector<float4> pts((size_t)ix->count[L] * kLeaf), boxes((size_t)2 * nbox);
    std::vector<int> pos_of((size_t)m);
    for (size_t i = 0; i < pts.size(); ++i) {
        if (i < (size_t)m) {
            const int oi = ids[i];
            pos_of[oi] = (int)i;
            pts[i] = make_float4(h_points[(size_t)oi * 3], h_points[(size_t)oi * 3 + 1], h_points[(size_t)oi * 3 + 2], 0.f);
            reinterpret_cast<int &>(pts[i].w) = oi;
        } else {
            pts[i] = make_float4(inf, inf, inf, 0.f);
            reinterpret_cast<int &>(pts[i].w) = 0x7fffffff;
        }
    }
    for (int k = 1; k <= L; ++k)
        for (int i = 0; i < ix->count[k]; ++i) {
            float4 l = make_float4(inf, inf, inf, 0.f), h = make_float4(-inf, -inf, -inf, 0.f);
            const size_t first = (size_t)i * size[k], last = std::min<size_t>((size_t)m, first + size[k]);
            for (size_t p = first; p < last; ++p) {
                const float4 &pt = pts[p];
                l.x = pt.x < l.x ? pt.x : l.x; l.y = pt.y < l.y ? pt.y : l.y; l.z = pt.z < l.z ? pt.z : l.z;
                h.x = pt.x > h.x ? pt.x : h.x; h.y = pt.y > h.y ? pt.y : h.y; h.z = pt.z > h.z ? pt.z : h.z;
            }
            boxes[(size_t)ix->off[k] + i] = l;
            boxes[(size_t)nbox + ix->off[k] + i] = h;
        }
    ix->bytes = 0;
    int rc = PSI_OK;
    auto up = [&](void **dst, const void *src, size_t nb) {
        if (rc != PSI_OK) return;
        if (cudaMalloc(dst, nb) != cudaSuccess) { rc = PSI_ERR_ALLOC; return; }
        if (cudaMemcpyAsync(*dst, src, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PSI_ERR_ALLOC;
        ix->bytes += nb;
    };
    up((void **)&ix->pts, pts.data(), pts.size() * sizeof(float4));
    up((void **)&ix->boxes, boxes.data(), boxes.size() * sizeof(float4));
    up((void **)&ix->pos_of, pos_of.data(), pos_of.size() * sizeof(int));
    if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;
    if (rc != PSI_OK) {
        psi_nn_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return PSI_OK;
}

size_t psi_nn_index_bytes(const psi_nn_index *ix) { return ix ? ix->bytes : 0; }

int psi_nn_index_query_hint(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                            const int *qsel, float *dist, int *idx, int *hint, psi_stream_t stream) {
    using namespace psi;
    if (!ix || B < 0 || n < 0) return PSI_ERR_BAD_ARG;
    if (B == 0 || n == 0) return PSI_OK;
    if (!q || !dist) return PSI_ERR_BAD_ARG;
    const long total = (long)B * n;
    const size_t stack_bytes = (size_t)kGroupsPerCta * kStackDepth * sizeof(uint2);
    const size_t box_bytes = (size_t)2 * ix->nbox * sizeof(float4);
    const bool smem = stack_bytes + box_bytes <= kIdxSmemMax;
    long blocks = (total + kGroupsPerCta - 1) / kGroupsPerCta;
    const long cap = (long)PSI_NUM_SMS * (smem ? 2 : 4);
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(nn_tree_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIdxSmemMax);
            attr = true;
        }
        nn_tree_query_kernel<true><<<(unsigned)blocks, kIdxThreads, stack_bytes + box_bytes, st>>>(
            *ix, q, q_bstride, n, qsel, total, dist, idx, hint);
    } else {
        nn_tree_query_kernel<false><<<(unsigned)blocks, kIdxThreads, stack_bytes, st>>>(
            *ix, q, q_bstride, n, qsel, total, dist, idx, hint);
    }
    PSI_LAUNCHED();
    return PSI_OK;
}

int psi_nn_index_query(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                       const int *qsel, float *dist, int *idx, psi_stream_t stream) {
    return psi_nn_index_query_hint(ix, q, q_bstride, B, n, qsel, dist, idx, nullptr, stream);
}

}  // extern "C"
