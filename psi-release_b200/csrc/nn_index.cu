// Exact nearest neighbour against a STATIC scene through a three-level box index.
//
// Same contract as psi_nn_fwd with a shared scene (chamfer_pytorch/chamfer.cu:12-134 semantics:
// d = fma(dz,dz,fma(dx,dx,rn(dy*dy))), lowest ORIGINAL index wins ties) -- the outputs are
// bit-identical to the brute-force kernel; only provably irrelevant pair evaluations are skipped.
//
// Index (built once per scene on the host): a balanced kd bisection (longest axis, part sizes kept
// multiples of the level below) orders the points so that 32 consecutive points form a compact
// CLUSTER, 8 clusters a SUPER (256 points), 8 supers a MEGA (2048 points); every node has an
// axis-aligned box.  Boxes are stored SoA (all lo, then all hi) and live in shared memory.
//
// Why it is exact.  For a query q and a box [lo,hi], g_a = max(fl(lo_a-q_a), fl(q_a-hi_a), 0)
// satisfies g_a <= |fl(s_a-q_a)| for every point s in the box (rounding is monotone), and
// fma(gz,gz,fma(gx,gx,rn(gy*gy))) is monotone in |g|, so lb(q,box) <= d(q,s) IN FLOATING POINT for
// every s in the box.  A box is skipped only if lb > ub where ub is the distance of a point
// already evaluated, so every skipped point has d > final minimum and cannot win or tie.
// Visited points are compared lexicographically on (d, original index).
//
// Three schedules return the same bits (psi_nn_index_query_mode): one warp per query (below; generic,
// incoherent query sets), one thread per query, and one warp per GROUP of 32 consecutive queries sharing
// one tree walk (nn_index_group_kernel, what the fitting loop uses).
//
// Warp-per-query execution.  A 32-lane round evaluates 32 mega boxes, or the 8 children of
// up to 4 admitted parents at once (lane = parent slot*8 + child); an admitted cluster is ONE
// coalesced 512-byte load, one point per lane; the running bound is a warp-wide redux.sync.min on
// the distance bits (d >= 0: the IEEE pattern is monotone).  The bound is seeded either
// best-first (nearest mega -> super -> cluster) or from a HINT: the cluster that held the
// nearest neighbour of the same query in the previous fitting iteration (bodies move little per
// Adam step), which is just another visited cluster, so exactness does not depend on it.
#include "common.cuh"
#include <algorithm>
#include <math.h>
#include <math_constants.h>
#include <new>
#include <stdlib.h>
#include <vector>

namespace psi {
constexpr int kLeaf = 32;         // points per cluster (one per lane)
constexpr int kFan = 8;           // children per super / mega
constexpr int kMaxMegaRounds = 8; // mega bounds kept in registers: 8*32 megas = 524 288 points
constexpr int kIdxThreads = 512;
constexpr size_t kIdxSmemMax = 72 * 1024;
}  // namespace psi

struct psi_nn_index {
    int m, num_megas, num_supers, num_clusters, mpad, rounds;
    float4 *pts;    // [num_clusters*32] (x,y,z,orig index bits); pads = +inf / INT_MAX
    float4 *boxes;  // SoA: [lo: mega(mpad) | super | cluster][hi: same]; empty nodes lo=hi=+inf
    int *pos_of;    // sorted position of every original point (index -> cluster, for the hint)
    float4 *pts2;   // thread kernel: point PAIRS {x0,x1,y0,y1},{z0,z1,idx0,idx1}  [num_clusters*16][2]
    float4 *box2;   // thread kernel: node PAIRS {lo.x0,lo.x1,lo.y0,lo.y1},{lo.z0,lo.z1,-hi.x0,-hi.x1},
                    //                {-hi.y0,-hi.y1,-hi.z0,-hi.z1}   [nbox/2][3]  (levels start at even nodes)
    int nbox;       // mpad + num_supers + num_clusters
    float p0x, p0y, p0z;   // original point 0: the answer for a query no point compares below +inf for
    size_t bytes;
};

namespace psi {

__device__ __forceinline__ unsigned box_lb(const float4 lo, const float4 hi, float qx, float qy, float qz) {
    const float gx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
    return __float_as_uint(__fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy))));
}

// No point won (every distance NaN or +inf: a diverged body).  The reference's `k == 0 ||` clauses
// (chamfer.cu:36,121,126) then leave index 0 with its distance, and so does the brute-force kernel; an
// INT_MAX "no index" would be dereferenced by the gradient kernels.
__device__ __forceinline__ float dist_point0(const psi_nn_index &ix, float qx, float qy, float qz) {
    const float dx = __fsub_rn(ix.p0x, qx), dy = __fsub_rn(ix.p0y, qy), dz = __fsub_rn(ix.p0z, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

struct Query {
    float x, y, z;
    float bd;        // lane-local best distance
    int bi;          // ... and its original index
    unsigned ub;     // warp-uniform bound (bits of a distance already seen)
    int lane;
};

template <bool SMEM>
__device__ __forceinline__ unsigned node_lb(const psi_nn_index &ix, const float4 *boxes, int node,
                                            const Query &q) {
    const float4 lo = SMEM ? boxes[node] : __ldg(boxes + node);
    const float4 hi = SMEM ? boxes[ix.nbox + node] : __ldg(boxes + ix.nbox + node);
    return box_lb(lo, hi, q.x, q.y, q.z);
}

// all lanes: the 32 points of cluster c
__device__ __forceinline__ void visit_cluster(const float4 *__restrict__ pts, int c, Query &q) {
    const float4 p = __ldg(pts + (size_t)c * kLeaf + q.lane);
    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const int oi = __float_as_int(p.w);
    if (d < q.bd || (d == q.bd && oi < q.bi)) {
        q.bd = d;
        q.bi = oi;
    }
    q.ub = min(q.ub, __reduce_min_sync(0xffffffffu, __float_as_uint(d)));
}

// take up to 4 set bits out of a warp-uniform mask; lane group g = lane/8 gets the g-th one (or -1)
__device__ __forceinline__ int take4(unsigned &mask, int lane) {
    int pos[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        pos[k] = __ffs(mask) - 1;          // -1 when the mask is empty
        mask &= mask - (mask != 0);
    }
    const int g = lane >> 3;
    return g == 0 ? pos[0] : (g == 1 ? pos[1] : (g == 2 ? pos[2] : pos[3]));
}

// Expand admitted supers (bits of `smask` over lanes holding `my_super`) into clusters and visit.
template <bool SMEM>
__device__ __forceinline__ void sweep_supers(const psi_nn_index &ix, const float4 *boxes, unsigned smask,
                                             unsigned slb, int my_super, int skip_cluster, Query &q) {
    const int cbase = ix.mpad + ix.num_supers;
    while (smask) {
        const int src = take4(smask, q.lane);                       // lane that holds my parent super
        const int sid = __shfl_sync(0xffffffffu, my_super, src < 0 ? 0 : src);
        const int cid = sid * kFan + (q.lane & 7);
        unsigned clb = 0x7f800000u;
        if (src >= 0) clb = node_lb<SMEM>(ix, boxes, cbase + cid, q);
        unsigned cmask = __ballot_sync(0xffffffffu, clb <= q.ub && cid != skip_cluster && src >= 0);
        while (cmask) {
            const int cl = __ffs(cmask) - 1;
            cmask &= cmask - 1;
            visit_cluster(ix.pts, __shfl_sync(0xffffffffu, cid, cl), q);
            cmask &= __ballot_sync(0xffffffffu, clb <= q.ub);       // the bound may have tightened
        }
        smask &= __ballot_sync(0xffffffffu, slb <= q.ub);
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(kIdxThreads)
nn_index_query_kernel(const psi_nn_index ix, const float *__restrict__ q_in, long q_bstride, int n,
                      const int *__restrict__ qsel, long total, float *__restrict__ dist,
                      int *__restrict__ idx, int *__restrict__ hint) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float4 *boxes = ix.boxes;
    if (SMEM) {
        float4 *s4 = reinterpret_cast<float4 *>(smem_raw);
        for (int i = threadIdx.x; i < 2 * ix.nbox; i += blockDim.x) s4[i] = __ldg(ix.boxes + i);
        __syncthreads();
        boxes = s4;
    }
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long warps = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < total; t += warps) {
        const long b = t / n;
        const long j = t - b * n;
        const float *qp = q_in + b * q_bstride + (qsel ? (long)__ldg(qsel + j) : j) * 3;
        Query q;
        q.x = __ldg(qp); q.y = __ldg(qp + 1); q.z = __ldg(qp + 2);
        q.bd = CUDART_INF_F;
        q.bi = 0x7fffffff;
        q.ub = 0x7f800000u;
        q.lane = lane;
        // level 0: every mega bound (32 per round), kept in registers
        unsigned mlb[kMaxMegaRounds];
#pragma unroll
        for (int r = 0; r < kMaxMegaRounds; ++r) {
            mlb[r] = 0x7f800000u;
            if (r < ix.rounds) mlb[r] = node_lb<SMEM>(ix, boxes, r * 32 + lane, q);
        }
        // seed the bound
        int seeded;
        const int h = hint ? hint[t] : -1;
        if (h >= 0 && (long)h * kLeaf < ix.m) {
            seeded = h;                                             // last iteration's winning cluster
        } else {                                                    // best-first descent
            unsigned best = 0xffffffffu;
            int m0 = 0;
#pragma unroll
            for (int r = 0; r < kMaxMegaRounds; ++r)
                if (r < ix.rounds) {
                    const unsigned mn = __reduce_min_sync(0xffffffffu, mlb[r]);
                    if (mn < best) {
                        best = mn;
                        m0 = r * 32 + __ffs(__ballot_sync(0xffffffffu, mlb[r] == mn)) - 1;
                    }
                }
            unsigned lb8 = 0x7f800000u;
            if (lane < kFan) lb8 = node_lb<SMEM>(ix, boxes, ix.mpad + m0 * kFan + lane, q);
            unsigned mn = __reduce_min_sync(0xffffffffu, lb8);
            const int s0 = m0 * kFan + __ffs(__ballot_sync(0xffffffffu, lb8 == mn)) - 1;
            lb8 = 0x7f800000u;
            if (lane < kFan) lb8 = node_lb<SMEM>(ix, boxes, ix.mpad + ix.num_supers + s0 * kFan + lane, q);
            mn = __reduce_min_sync(0xffffffffu, lb8);
            seeded = s0 * kFan + __ffs(__ballot_sync(0xffffffffu, lb8 == mn)) - 1;
        }
        visit_cluster(ix.pts, seeded, q);
        // sweep: megas -> supers -> clusters, 4 parents per 32-lane round
#pragma unroll
        for (int r = 0; r < kMaxMegaRounds; ++r) {
            if (r < ix.rounds) {
                unsigned mmask = __ballot_sync(0xffffffffu, mlb[r] <= q.ub && r * 32 + lane < ix.num_megas);
                while (mmask) {
                    const int src = take4(mmask, lane);
                    const int my_super = (r * 32 + (src < 0 ? 0 : src)) * kFan + (lane & 7);
                    unsigned slb = 0x7f800000u;
                    if (src >= 0) slb = node_lb<SMEM>(ix, boxes, ix.mpad + my_super, q);
                    const unsigned smask = __ballot_sync(0xffffffffu, slb <= q.ub && src >= 0);
                    sweep_supers<SMEM>(ix, boxes, smask, slb, my_super, seeded, q);
                    mmask &= __ballot_sync(0xffffffffu, mlb[r] <= q.ub);
                }
            }
        }
        // lexicographic (d, original index) minimum over the lanes
        const unsigned dmin = __reduce_min_sync(0xffffffffu, __float_as_uint(q.bd));
        const bool win = __float_as_uint(q.bd) == dmin;
        const unsigned imin = __reduce_min_sync(0xffffffffu, win ? (unsigned)q.bi : 0x7fffffffu);
        if (lane == 0) {
            const bool none = imin >= (unsigned)ix.m;
            dist[t] = none ? dist_point0(ix, q.x, q.y, q.z) : __uint_as_float(dmin);
            if (idx) idx[t] = none ? 0 : (int)imin;
            if (hint) hint[t] = none ? -1 : __ldg(ix.pos_of + imin) / kLeaf;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Thread-per-query walk of the same index, for LARGE, SPATIALLY COHERENT query sets (the fitting
// loop orders the contact vertices along a kd curve of the template mesh, so the 32 queries of
// a warp are neighbours on the body and walk nearly the same nodes: loop control, box loads
// (shared-memory broadcasts) and leaf loads (same 512-byte leaf) are shared by the warp).
// Same pruning rule, same lexicographic (d, original index) minimum => same bits as the
// warp-per-query kernel and as the brute force; only the schedule differs.
template <bool SMEM>
__device__ __forceinline__ float node_lbf(const psi_nn_index &ix, const float4 *boxes, int node,
                                          float qx, float qy, float qz) {
    const float4 lo = SMEM ? boxes[node] : __ldg(boxes + node);
    const float4 hi = SMEM ? boxes[ix.nbox + node] : __ldg(boxes + ix.nbox + node);
    return __uint_as_float(box_lb(lo, hi, qx, qy, qz));
}

template <bool SMEM>
__global__ void __launch_bounds__(256, 4)
nn_index_thread_kernel(const psi_nn_index ix, const float *__restrict__ q_in, long q_bstride, int n,
                       const int *__restrict__ qsel, long total, float *__restrict__ dist,
                       int *__restrict__ idx, int *__restrict__ hint) {
    // Packed FP32 (FADD2 / FMUL2 / FFMA2): two boxes or two points per instruction, from the
    // pair-interleaved copies box2 / pts2.  Only the mega + super pairs (5.6 kB at 50 000 points)
    // are staged in shared memory so that 8 CTAs stay resident per SM (the walk is latency
    // bound); the cluster pairs come through L1, where the coherent warps hit.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ntop2 = (ix.mpad + ix.num_supers) / 2;           // node pairs of the two top levels
    float4 *s4 = reinterpret_cast<float4 *>(smem_raw);
    if (SMEM) {
        for (int i = threadIdx.x; i < ntop2 * 3; i += blockDim.x) s4[i] = __ldg(ix.box2 + i);
        __syncthreads();
    }
    pdl_wait();
    const int sbase = ix.mpad, cbase = ix.mpad + ix.num_supers;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long b = t / n;
        const long j = t - b * n;
        const float *qp = q_in + b * q_bstride + (qsel ? (long)__ldg(qsel + j) : j) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        const float2 q2x = make_float2(qx, qx), q2y = make_float2(qy, qy), q2z = make_float2(qz, qz);
        const float2 n2x = make_float2(-qx, -qx), n2y = make_float2(-qy, -qy), n2z = make_float2(-qz, -qz);
        float bd = CUDART_INF_F;
        int bi = 0x7fffffff;
        // bounds of the node pair (2p, 2p+1)
        auto lb2 = [&](int pair) -> float2 {
            float4 A, Bq, C;
            if (SMEM && pair < ntop2) {
                A = s4[pair * 3]; Bq = s4[pair * 3 + 1]; C = s4[pair * 3 + 2];
            } else {
                A = __ldg(ix.box2 + (size_t)pair * 3); Bq = __ldg(ix.box2 + (size_t)pair * 3 + 1);
                C = __ldg(ix.box2 + (size_t)pair * 3 + 2);
            }
            const float2 lx = __fadd2_rn(make_float2(A.x, A.y), n2x), hx = __fadd2_rn(make_float2(Bq.z, Bq.w), q2x);
            const float2 ly = __fadd2_rn(make_float2(A.z, A.w), n2y), hy = __fadd2_rn(make_float2(C.x, C.y), q2y);
            const float2 lz = __fadd2_rn(make_float2(Bq.x, Bq.y), n2z), hz = __fadd2_rn(make_float2(C.z, C.w), q2z);
            const float2 gx = make_float2(fmaxf(fmaxf(lx.x, hx.x), 0.f), fmaxf(fmaxf(lx.y, hx.y), 0.f));
            const float2 gy = make_float2(fmaxf(fmaxf(ly.x, hy.x), 0.f), fmaxf(fmaxf(ly.y, hy.y), 0.f));
            const float2 gz = make_float2(fmaxf(fmaxf(lz.x, hz.x), 0.f), fmaxf(fmaxf(lz.y, hz.y), 0.f));
            return __ffma2_rn(gz, gz, __ffma2_rn(gx, gx, __fmul2_rn(gy, gy)));
        };
        // leaf c: pass 1 keeps only the minimum distance (one FMNMX3 per point pair); the index is
        // recovered by a second pass only when the leaf improves or ties the best so far
        auto visit = [&](int c) {
            const float4 *p = ix.pts2 + (size_t)c * kLeaf;          // 16 pairs x 2 float4
            float lm = CUDART_INF_F;
#pragma unroll 8
            for (int i = 0; i < kLeaf / 2; ++i) {
                const float4 u = __ldg(p + 2 * i), w = __ldg(p + 2 * i + 1);
                const float2 dx = __fadd2_rn(make_float2(u.x, u.y), n2x);
                const float2 dy = __fadd2_rn(make_float2(u.z, u.w), n2y);
                const float2 dz = __fadd2_rn(make_float2(w.x, w.y), n2z);
                const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
                lm = fminf(lm, fminf(d.x, d.y));
            }
            if (lm <= bd) {
                int li = 0x7fffffff;
#pragma unroll 4
                for (int i = 0; i < kLeaf / 2; ++i) {
                    const float4 u = __ldg(p + 2 * i), w = __ldg(p + 2 * i + 1);
                    const float2 dx = __fadd2_rn(make_float2(u.x, u.y), n2x);
                    const float2 dy = __fadd2_rn(make_float2(u.z, u.w), n2y);
                    const float2 dz = __fadd2_rn(make_float2(w.x, w.y), n2z);
                    const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
                    if (d.x == lm) li = min(li, __float_as_int(w.z));
                    if (d.y == lm) li = min(li, __float_as_int(w.w));
                }
                if (lm < bd || li < bi) bi = li;       // lm == bd: lowest original index wins
                bd = lm;
            }
        };
        // seed: hinted leaf, else greedy descent (nearest mega -> super -> cluster)
        int seeded = hint ? hint[t] : -1;
        if (seeded < 0 || (long)seeded * kLeaf >= ix.m) {
            float best = CUDART_INF_F;
            int m0 = 0;
            for (int g = 0; g < ix.mpad / 2; ++g) {
                const float2 lb = lb2(g);
                if (lb.x < best) { best = lb.x; m0 = 2 * g; }
                if (lb.y < best) { best = lb.y; m0 = 2 * g + 1; }
            }
            best = CUDART_INF_F;
            int s0 = m0 * kFan;
            for (int pr = 0; pr < kFan / 2; ++pr) {
                const float2 lb = lb2((sbase + m0 * kFan) / 2 + pr);
                if (lb.x < best) { best = lb.x; s0 = m0 * kFan + 2 * pr; }
                if (lb.y < best) { best = lb.y; s0 = m0 * kFan + 2 * pr + 1; }
            }
            best = CUDART_INF_F;
            seeded = s0 * kFan;
            for (int pr = 0; pr < kFan / 2; ++pr) {
                const float2 lb = lb2((cbase + s0 * kFan) / 2 + pr);
                if (lb.x < best) { best = lb.x; seeded = s0 * kFan + 2 * pr; }
                if (lb.y < best) { best = lb.y; seeded = s0 * kFan + 2 * pr + 1; }
            }
        }
        visit(seeded);
        // ordered sweep with pruning against this query's best distance so far
#pragma unroll 1
        for (int g2 = 0; g2 < (ix.num_megas + 1) / 2; ++g2) {
            const float2 ml = lb2(g2);
#pragma unroll 1
            for (int gh = 0; gh < 2; ++gh) {
                const int g = 2 * g2 + gh;
                if ((gh ? ml.y : ml.x) > bd || g >= ix.num_megas) continue;
#pragma unroll 1
                for (int sp = 0; sp < kFan / 2; ++sp) {
                    const float2 sl = lb2((sbase + g * kFan) / 2 + sp);
#pragma unroll 1
                    for (int sh = 0; sh < 2; ++sh) {
                        const int sidx = g * kFan + 2 * sp + sh;
                        if ((sh ? sl.y : sl.x) > bd) continue;
#pragma unroll 1
                        for (int cp = 0; cp < kFan / 2; ++cp) {
                            const float2 cl = lb2((cbase + sidx * kFan) / 2 + cp);
                            const int c0 = sidx * kFan + 2 * cp;
                            if (cl.x <= bd && c0 != seeded) visit(c0);
                            if (cl.y <= bd && c0 + 1 != seeded) visit(c0 + 1);
                        }
                    }
                }
            }
        }
        const bool none = (unsigned)bi >= (unsigned)ix.m;
        dist[t] = none ? dist_point0(ix, qx, qy, qz) : bd;
        if (idx) idx[t] = none ? 0 : bi;
        if (hint) hint[t] = none ? -1 : __ldg(ix.pos_of + bi) / kLeaf;
    }
}

// ---------------------------------------------------------------------------------------------
// GROUP schedule: one warp owns 32 CONSECUTIVE queries of one body and walks the tree ONCE for
// all of them.  (The thread-per-query kernel pays for the union of its 32 lanes' walks anyway --
// divergent lanes wait -- and measured ~7000 warp instructions per 32 queries, most of them box
// tests and loop control.)  Here the walk is warp-uniform:
//   * nodes are tested lane-parallel (lane = node) against the group's query box [qlo,qhi] and the
//     group bound ub = max over lanes of the lane's best distance so far;
//   * an admitted leaf is tested once more against every lane's own query/bound, and if any lane
//     still needs it ALL lanes evaluate its 32 points (packed f32x2, broadcast loads);
//   * only the minimum distance and the leaf that produced it are tracked; the original index is
//     recovered at the end by one scan of the winning leaf per lane (ties between leaves are
//     resolved on the spot: lowest original index, as everywhere else).
// Exactness.  For a lane with qlo <= q <= qhi (per axis), fl(lo-qhi) <= fl(lo-q) and
// fl(qlo-hi) <= fl(q-hi) (rounding is monotone), hence lb_group(node) <= lb_lane(node) <= d(q,s) for
// every point s of the node, all in floating point.  A node is skipped only if lb_group > ub >=
// the lane's current best, a leaf only if lb_lane > the lane's best for EVERY lane; lanes that
// evaluate extra leaves only see more real points.  Same bits as every other schedule.
__device__ __forceinline__ unsigned fkey(float f) {          // order-preserving float -> uint
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ unsigned box_lb_group(const float4 lo, const float4 hi, float lx, float ly, float lz,
                                                 float hx, float hy, float hz) {
    const float gx = fmaxf(fmaxf(__fsub_rn(lo.x, hx), __fsub_rn(lx, hi.x)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(lo.y, hy), __fsub_rn(ly, hi.y)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(lo.z, hz), __fsub_rn(lz, hi.z)), 0.f);
    return __float_as_uint(__fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy))));
}

// minimum distance from (qx,qy,qz) (given negated, duplicated) to the 32 points of leaf c (warp-uniform).
// The leaf (1 kB) is fetched by the warp with two coalesced 512-byte loads into its shared-memory
// slot and then read back as broadcasts: one L1 round trip per leaf instead of 32 uniform loads.
__device__ __forceinline__ float leaf_min(const float4 *__restrict__ pts2, int c, float4 *wbuf, int lane,
                                          float2 n2x, float2 n2y, float2 n2z) {
    const float4 *p = pts2 + (size_t)c * kLeaf;                 // 16 pairs x 2 float4
    const float4 v0 = __ldg(p + lane), v1 = __ldg(p + 32 + lane);
    __syncwarp();                                               // the previous leaf has been consumed
    wbuf[lane] = v0;
    wbuf[32 + lane] = v1;
    __syncwarp();
    float lm = CUDART_INF_F;
#pragma unroll 8
    for (int i = 0; i < kLeaf / 2; ++i) {
        const float4 u = wbuf[2 * i], w = wbuf[2 * i + 1];
        const float2 dx = __fadd2_rn(make_float2(u.x, u.y), n2x);
        const float2 dy = __fadd2_rn(make_float2(u.z, u.w), n2y);
        const float2 dz = __fadd2_rn(make_float2(w.x, w.y), n2z);
        const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
        lm = fminf(lm, fminf(d.x, d.y));
    }
    return lm;
}
// lowest original index among the points of leaf c at distance exactly lm
__device__ __noinline__ int leaf_arg(const float4 *__restrict__ pts2, int c, float lm, float qx, float qy, float qz) {
    const float4 *p = pts2 + (size_t)c * kLeaf;
    const float2 n2x = make_float2(-qx, -qx), n2y = make_float2(-qy, -qy), n2z = make_float2(-qz, -qz);
    int li = 0x7fffffff;
#pragma unroll 4
    for (int i = 0; i < kLeaf / 2; ++i) {
        const float4 u = __ldg(p + 2 * i), w = __ldg(p + 2 * i + 1);
        const float2 dx = __fadd2_rn(make_float2(u.x, u.y), n2x);
        const float2 dy = __fadd2_rn(make_float2(u.z, u.w), n2y);
        const float2 dz = __fadd2_rn(make_float2(w.x, w.y), n2z);
        const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
        if (d.x == lm) li = min(li, __float_as_int(w.z));
        if (d.y == lm) li = min(li, __float_as_int(w.w));
    }
    return li;
}

#ifndef PSI_NN_GRP_THREADS
#define PSI_NN_GRP_THREADS 256
#endif
constexpr int kGrpThreads = PSI_NN_GRP_THREADS;
#ifndef PSI_NN_GRP_MINB
#define PSI_NN_GRP_MINB 4
#endif
#ifndef PSI_NN_SUBGROUPS
#define PSI_NN_SUBGROUPS 4
#endif
constexpr int kSubGroups = PSI_NN_SUBGROUPS;       // sub-groups of a 32-query group with their own box and bound (8 measured slower: 129 vs 118 us)
constexpr int kSubSize = 32 / kSubGroups;

// -DPSI_NN_STATS: event counters of the group walk (debug builds only; tools/nn_stats.py)
#ifdef PSI_NN_STATS
__device__ unsigned long long g_nn_stats[8];
#define NN_STAT(i) do { if (lane == 0) atomicAdd(&g_nn_stats[i], 1ull); } while (0)
#else
#define NN_STAT(i) do { } while (0)
#endif

template <bool SMEM>
__global__ void __launch_bounds__(kGrpThreads, PSI_NN_GRP_MINB)
nn_index_group_kernel(const psi_nn_index ix, const float *__restrict__ q_in, long q_bstride, int n,
                      const int *__restrict__ qsel, int B, float *__restrict__ dist,
                      int *__restrict__ idx, int *__restrict__ hint) {
    // mega + super boxes (lo | hi, SoA) in shared memory; cluster boxes come through L1
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float4 s_sub[kGrpThreads / 32][kSubGroups][2];   // per warp: sub-group query boxes {lo, ub bits | hi}
    __shared__ float4 s_leaf[kGrpThreads / 32][2 * kLeaf];      // per warp: the leaf being evaluated
    const int ntop = ix.mpad + ix.num_supers;
    const float4 *top_lo = ix.boxes, *top_hi = ix.boxes + ix.nbox;
    if (SMEM) {
        float4 *s4 = reinterpret_cast<float4 *>(smem_raw);
        for (int i = threadIdx.x; i < ntop; i += blockDim.x) {
            s4[i] = __ldg(ix.boxes + i);
            s4[ntop + i] = __ldg(ix.boxes + ix.nbox + i);
        }
        __syncthreads();
        top_lo = s4;
        top_hi = s4 + ntop;
    }
    pdl_wait();
    const float4 *__restrict__ c_lo = ix.boxes + ntop, *__restrict__ c_hi = ix.boxes + ix.nbox + ntop;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned submask = ((1u << kSubSize) - 1u) << (lane & ~(kSubSize - 1));   // my sub-group: kSubSize consecutive queries
    float4(*sub)[2] = s_sub[threadIdx.x >> 5];
    const int wpb = (n + 31) / 32;                               // warps (groups) per body
    const long ngroups = (long)B * wpb;
    const long nwarps = (long)gridDim.x * (kGrpThreads / 32);
    for (long gw = (long)blockIdx.x * (kGrpThreads / 32) + (threadIdx.x >> 5); gw < ngroups; gw += nwarps) {
        const long b = gw / wpb;
        const int j = (int)(gw - b * wpb) * 32 + lane;
        const int jj = min(j, n - 1);                            // lanes past the end duplicate the last query
        const long t = b * n + jj;
        const float *qp = q_in + b * q_bstride + (qsel ? (long)__ldg(qsel + jj) : (long)jj) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        const float2 n2x = make_float2(-qx, -qx), n2y = make_float2(-qy, -qy), n2z = make_float2(-qz, -qz);
        NN_STAT(0);
        // query boxes: the 4 sub-groups (shared memory) and the whole group (registers)
        float lx, ly, lz, hx, hy, hz;
        {
            const unsigned kx = fkey(qx), ky = fkey(qy), kz = fkey(qz);
            const unsigned ax = __reduce_min_sync(submask, kx), bx = __reduce_max_sync(submask, kx);
            const unsigned ay = __reduce_min_sync(submask, ky), by = __reduce_max_sync(submask, ky);
            const unsigned az = __reduce_min_sync(submask, kz), bz = __reduce_max_sync(submask, kz);
            __syncwarp();
            if ((lane & (kSubSize - 1)) == 0) {
                sub[lane / kSubSize][0] = make_float4(fkey_inv(ax), fkey_inv(ay), fkey_inv(az), CUDART_INF_F);
                sub[lane / kSubSize][1] = make_float4(fkey_inv(bx), fkey_inv(by), fkey_inv(bz), 0.f);
            }
            lx = fkey_inv(__reduce_min_sync(full, ax)); hx = fkey_inv(__reduce_max_sync(full, bx));
            ly = fkey_inv(__reduce_min_sync(full, ay)); hy = fkey_inv(__reduce_max_sync(full, by));
            lz = fkey_inv(__reduce_min_sync(full, az)); hz = fkey_inv(__reduce_max_sync(full, bz));
        }
        float bd = CUDART_INF_F;     // this lane's best distance so far ...
        int bleaf = -1;              // ... the leaf it was found in ...
        int bi = -1;                 // ... and its original index once known (-1: not recovered yet)
        auto visit = [&](int c) {    // all lanes, warp-uniform c
            const float lm = leaf_min(ix.pts2, c, s_leaf[threadIdx.x >> 5], lane, n2x, n2y, n2z);
            if (lm < bd) {
                bd = lm; bleaf = c; bi = -1;
            } else if (lm == bd && c != bleaf && bleaf >= 0) {   // exact tie between two leaves
                if (bi < 0) bi = leaf_arg(ix.pts2, bleaf, bd, qx, qy, qz);
                bi = min(bi, leaf_arg(ix.pts2, c, lm, qx, qy, qz));
            }
        };
        // bounds: ub = max over the group, sub[g][0].w = max over sub-group g (bits of distances already seen)
        unsigned ub;
        auto publish_bounds = [&]() {
            const unsigned mine = __reduce_max_sync(submask, __float_as_uint(bd));
            ub = __reduce_max_sync(full, mine);
            if ((lane & (kSubSize - 1)) == 0) sub[lane / kSubSize][0].w = __uint_as_float(mine);
            __syncwarp();
        };
        // ---- seeds: last iteration's winning leaves, or a greedy descent for the lanes without one
        int h = hint ? hint[t] : -1;
        const bool need = h < 0 || (long)h * kLeaf >= ix.m;         // no hint, or not a leaf that holds points
        const unsigned needm = __ballot_sync(full, need);
        if (needm) {
            const int src = __ffs(needm) - 1;
            const float cx = __shfl_sync(full, qx, src), cy = __shfl_sync(full, qy, src), cz = __shfl_sync(full, qz, src);
            unsigned best = 0xffffffffu;
            int m0 = 0;
            for (int r = 0; r < ix.rounds; ++r) {
                const unsigned lb = box_lb(top_lo[r * 32 + lane], top_hi[r * 32 + lane], cx, cy, cz);
                const unsigned mn = __reduce_min_sync(full, lb);
                if (mn < best) { best = mn; m0 = r * 32 + __ffs(__ballot_sync(full, lb == mn)) - 1; }
            }
            unsigned lb8 = 0x7f800000u;
            if (lane < kFan) lb8 = box_lb(top_lo[ix.mpad + m0 * kFan + lane], top_hi[ix.mpad + m0 * kFan + lane], cx, cy, cz);
            unsigned mn = __reduce_min_sync(full, lb8);
            const int s0 = m0 * kFan + __ffs(__ballot_sync(full, lb8 == mn)) - 1;
            lb8 = 0x7f800000u;
            if (lane < kFan) lb8 = box_lb(__ldg(c_lo + s0 * kFan + lane), __ldg(c_hi + s0 * kFan + lane), cx, cy, cz);
            mn = __reduce_min_sync(full, lb8);
            const int c0 = s0 * kFan + __ffs(__ballot_sync(full, lb8 == mn)) - 1;
            if (need) h = c0;
        }
        int myvis = -1, nvis = 0;    // lane k remembers the k-th seed leaf (at most 32 distinct seeds)
        {
            unsigned todo = full;
            while (todo) {
                const int c = __shfl_sync(full, h, __ffs(todo) - 1);
                NN_STAT(1);
                visit(c);
                if (lane == nvis) myvis = c;
                ++nvis;
                todo &= ~__ballot_sync(full, h == c);
            }
        }
        publish_bounds();
        // ---- sweep: megas -> supers -> clusters, 32 nodes per round, all warp-uniform
        for (int r = 0; r < ix.rounds; ++r) {
            const unsigned lbm = box_lb_group(top_lo[r * 32 + lane], top_hi[r * 32 + lane], lx, ly, lz, hx, hy, hz);
            // (pad megas never enter: with ub = +inf -- NaN queries -- their children would be out of range)
            unsigned mmask = __ballot_sync(full, lbm <= ub && r * 32 + lane < ix.num_megas);
            while (mmask) {
                NN_STAT(2);
                const int src = take4(mmask, lane);
                const int my_super = (r * 32 + (src < 0 ? 0 : src)) * kFan + (lane & 7);
                unsigned slb = 0x7f800000u;
                if (src >= 0) slb = box_lb_group(top_lo[ix.mpad + my_super], top_hi[ix.mpad + my_super], lx, ly, lz, hx, hy, hz);
                unsigned smask = __ballot_sync(full, src >= 0 && slb <= ub);
                while (smask) {
                    NN_STAT(3);
                    const int src2 = take4(smask, lane);
                    const int cid = __shfl_sync(full, my_super, src2 < 0 ? 0 : src2) * kFan + (lane & 7);
                    // lane = cluster: admitted if any of the 4 sub-groups (tighter boxes, tighter bounds) may need it
                    bool adm = false;
                    if (src2 >= 0) {
                        const float4 clo = __ldg(c_lo + cid), chi = __ldg(c_hi + cid);
#pragma unroll
                        for (int g = 0; g < kSubGroups; ++g) {
                            const float4 sl = sub[g][0], sh = sub[g][1];
                            adm |= box_lb_group(clo, chi, sl.x, sl.y, sl.z, sh.x, sh.y, sh.z) <= __float_as_uint(sl.w);
                        }
                    }
                    unsigned cmask = __ballot_sync(full, adm);
                    for (int k = 0; k < nvis && cmask; ++k)                     // the seeds were evaluated already
                        cmask &= ~__ballot_sync(full, cid == __shfl_sync(full, myvis, k));
                    while (cmask) {
                        const int cl = __ffs(cmask) - 1;
                        cmask &= cmask - 1;
                        const int c = __shfl_sync(full, cid, cl);
                        NN_STAT(4);
                        // lane = query: does any lane still need leaf c?
                        const float lbl = __uint_as_float(box_lb(__ldg(c_lo + c), __ldg(c_hi + c), qx, qy, qz));
                        if (__any_sync(full, lbl <= bd)) {
                            NN_STAT(5);
                            visit(c);
                            publish_bounds();
                        }
                    }
                    smask &= __ballot_sync(full, slb <= ub);
                }
                mmask &= __ballot_sync(full, lbm <= ub);
            }
        }
        if (bi < 0) {
            if (bleaf >= 0) bi = leaf_arg(ix.pts2, bleaf, bd, qx, qy, qz);
            else { bi = 0; bd = dist_point0(ix, qx, qy, qz); }
        }
        if (j < n) {
            dist[t] = bd;
            if (idx) idx[t] = bi;
            if (hint) hint[t] = bleaf;
        }
    }
    pdl_launch_dependents();   // the rest is this kernel's tail: let the next kernel's CTAs be scheduled
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
    cudaFree(ix->pts2);
    cudaFree(ix->box2);
    delete ix;
}

int psi_nn_index_create(psi_nn_index **out, const float *h_points, int m, psi_stream_t stream) {
    psi::Range nvtx_range("psi_nn_index_create");
    using namespace psi;
    if (!out || !h_points || m < 1) return PSI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    for (size_t i = 0; i < (size_t)m * 3; ++i) {
        const float v = h_points[i];
        if (!(v == v) || v == HUGE_VALF || v == -HUGE_VALF) return PSI_ERR_BAD_ARG;  // finite only
    }
    const int mega_pts = kLeaf * kFan * kFan, super_pts = kLeaf * kFan;
    const int num_megas = (m + mega_pts - 1) / mega_pts;
    const int rounds = (num_megas + 31) / 32;
    if (rounds > kMaxMegaRounds) return PSI_ERR_UNSUPPORTED;
    std::vector<int> ids((size_t)m);
    for (int i = 0; i < m; ++i) ids[i] = i;
    kd_order(h_points, ids.data(), m, mega_pts);
    for (int b = 0; b < m; b += mega_pts) kd_order(h_points, ids.data() + b, std::min(mega_pts, m - b), super_pts);
    for (int b = 0; b < m; b += super_pts) kd_order(h_points, ids.data() + b, std::min(super_pts, m - b), kLeaf);
    psi_nn_index *ix = new (std::nothrow) psi_nn_index();
    if (!ix) return PSI_ERR_ALLOC;
    ix->pts = nullptr; ix->boxes = nullptr; ix->pos_of = nullptr; ix->pts2 = nullptr; ix->box2 = nullptr;
    ix->m = m;
    ix->p0x = h_points[0]; ix->p0y = h_points[1]; ix->p0z = h_points[2];
    ix->num_megas = num_megas;
    ix->rounds = rounds;
    ix->mpad = rounds * 32;
    ix->num_supers = num_megas * kFan;
    ix->num_clusters = ix->num_supers * kFan;
    ix->nbox = ix->mpad + ix->num_supers + ix->num_clusters;
    const float inf = HUGE_VALF;
    std::vector<float4> pts((size_t)ix->num_clusters * kLeaf), boxes((size_t)2 * ix->nbox);
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
    auto box_of = [&](size_t first, size_t count, int node) {
        float4 l = make_float4(inf, inf, inf, 0.f), h = make_float4(-inf, -inf, -inf, 0.f);
        bool any = false;
        for (size_t i = first; i < first + count && i < (size_t)m; ++i) {
            const float4 &p = pts[i];
            l.x = p.x < l.x ? p.x : l.x; l.y = p.y < l.y ? p.y : l.y; l.z = p.z < l.z ? p.z : l.z;
            h.x = p.x > h.x ? p.x : h.x; h.y = p.y > h.y ? p.y : h.y; h.z = p.z > h.z ? p.z : h.z;
            any = true;
        }
        if (!any) { l = make_float4(inf, inf, inf, 0.f); h = l; }   // empty: bound = +inf, never admitted
        boxes[node] = l;
        boxes[(size_t)ix->nbox + node] = h;
    };
    for (int g = 0; g < ix->mpad; ++g) box_of((size_t)g * mega_pts, g < num_megas ? mega_pts : 0, g);
    for (int s = 0; s < ix->num_supers; ++s) box_of((size_t)s * super_pts, super_pts, ix->mpad + s);
    for (int c = 0; c < ix->num_clusters; ++c) box_of((size_t)c * kLeaf, kLeaf, ix->mpad + ix->num_supers + c);
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
    {   // pair-interleaved copies for the packed (f32x2) thread-per-query kernel
        std::vector<float4> pts2(pts.size()), box2((size_t)(ix->nbox / 2) * 3);
        for (size_t p = 0; p < pts.size() / 2; ++p) {
            const float4 a = pts[2 * p], b = pts[2 * p + 1];
            pts2[2 * p] = make_float4(a.x, b.x, a.y, b.y);
            pts2[2 * p + 1] = make_float4(a.z, b.z, a.w, b.w);
        }
        for (int p = 0; p < ix->nbox / 2; ++p) {
            const float4 l0 = boxes[2 * p], l1 = boxes[2 * p + 1];
            const float4 h0 = boxes[(size_t)ix->nbox + 2 * p], h1 = boxes[(size_t)ix->nbox + 2 * p + 1];
            box2[(size_t)p * 3] = make_float4(l0.x, l1.x, l0.y, l1.y);
            box2[(size_t)p * 3 + 1] = make_float4(l0.z, l1.z, -h0.x, -h1.x);
            box2[(size_t)p * 3 + 2] = make_float4(-h0.y, -h1.y, -h0.z, -h1.z);
        }
        up((void **)&ix->pts2, pts2.data(), pts2.size() * sizeof(float4));
        up((void **)&ix->box2, box2.data(), box2.size() * sizeof(float4));
        if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;   // vectors die here
    }
    if (rc == PSI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PSI_ERR_ALLOC;
    if (rc != PSI_OK) {
        psi_nn_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return PSI_OK;
}

size_t psi_nn_index_bytes(const psi_nn_index *ix) { return ix ? ix->bytes : 0; }

#ifdef PSI_NN_STATS
__attribute__((visibility("default"))) int psi_debug_nn_stats(unsigned long long *h_out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h_out, psi::g_nn_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(psi::g_nn_stats, z, sizeof(z)); }
    return 0;
}
#endif

// mode 0: pick by query count; 1: warp per query; 2: thread per query
static int nn_index_query_impl(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                               const int *qsel, float *dist, int *idx, int *hint, int mode,
                               psi_stream_t stream) {
    psi::Range nvtx_range("psi_nn_index_query");
    using namespace psi;
    if (!ix || B < 0 || n < 0) return PSI_ERR_BAD_ARG;
    if (B == 0 || n == 0) return PSI_OK;
    if (!q || !dist) return PSI_ERR_BAD_ARG;
    const long total = (long)B * n;
    const int wpb = kIdxThreads / 32;
    const size_t box_bytes = (size_t)2 * ix->nbox * sizeof(float4);
    const bool smem = box_bytes <= kIdxSmemMax;
    cudaStream_t st = (cudaStream_t)stream;
    if (const int arc = ensure_max_dyn_smem(nn_index_query_kernel<true>, kIdxSmemMax)) return arc;
    if (mode == 3) {
        // one warp per 32 consecutive queries of a body (coherent query order: the fitting loop)
        const long ngroups = (long)B * ((n + 31) / 32);
        // one group per warp: the hardware block scheduler balances the uneven walks
        long blocks = (ngroups + kGrpThreads / 32 - 1) / (kGrpThreads / 32);
        const long cap = 1L << 20;
        if (blocks > cap) blocks = cap;
        const size_t top_bytes = (size_t)2 * (ix->mpad + ix->num_supers) * sizeof(float4);
        if (top_bytes <= 24 * 1024)
            (psi::skip_kernel(mode == 3 ? "nn_index_group" : "nn_index_query") ? cudaSuccess : launch_pdl_sel(nn_index_group_kernel<true>, dim3((unsigned)blocks), dim3(kGrpThreads), top_bytes, st, *ix, q, q_bstride, n, qsel, B, dist, idx, hint));
        else
            (psi::skip_kernel(mode == 3 ? "nn_index_group" : "nn_index_query") ? cudaSuccess : launch_pdl_sel(nn_index_group_kernel<false>, dim3((unsigned)blocks), dim3(kGrpThreads), 0, st, *ix, q, q_bstride, n, qsel, B, dist, idx, hint));
    } else if (mode == 2 || (mode == 0 && total >= (long)PSI_NUM_SMS * 768)) {
        // one thread per query: enough queries to fill the machine with 256-thread CTAs
        long blocks = (total + 255) / 256;
        const long cap = (long)PSI_NUM_SMS * 8;
        if (blocks > cap) blocks = cap;
        const size_t top_bytes = (size_t)3 * ((ix->mpad + ix->num_supers) / 2) * sizeof(float4);
        // 64 registers (no spills with the packed pairs), 4 CTAs/SM; measured best of 3/4/5/6/8
        if (top_bytes <= 24 * 1024)
            (psi::skip_kernel(mode == 3 ? "nn_index_group" : "nn_index_query") ? cudaSuccess : launch_pdl(nn_index_thread_kernel<true>, dim3((unsigned)blocks), dim3(256), top_bytes, st, *ix, q, q_bstride, n, qsel, total, dist, idx, hint));
        else
            (psi::skip_kernel(mode == 3 ? "nn_index_group" : "nn_index_query") ? cudaSuccess : launch_pdl(nn_index_thread_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, st, *ix, q, q_bstride, n, qsel, total, dist, idx, hint));
    } else {
        long blocks = (total + wpb - 1) / wpb;
        const long cap = (long)PSI_NUM_SMS * (smem ? 3 : 4);
        if (blocks > cap) blocks = cap;
        if (smem)
            (psi::skip_kernel(mode == 3 ? "nn_index_group" : "nn_index_query") ? cudaSuccess : launch_pdl(nn_index_query_kernel<true>, dim3((unsigned)blocks), dim3(kIdxThreads), box_bytes, st, *ix, q, q_bstride, n, qsel, total, dist, idx, hint));
        else
            (psi::skip_kernel(mode == 3 ? "nn_index_group" : "nn_index_query") ? cudaSuccess : launch_pdl(nn_index_query_kernel<false>, dim3((unsigned)blocks), dim3(kIdxThreads), 0, st, *ix, q, q_bstride, n, qsel, total, dist, idx, hint));
    }
    PSI_LAUNCHED_K(mode == 3 ? "nn_index_group" : "nn_index_query");
    return PSI_OK;
}

int psi_nn_index_query_hint(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                            const int *qsel, float *dist, int *idx, int *hint, psi_stream_t stream) {
    return nn_index_query_impl(ix, q, q_bstride, B, n, qsel, dist, idx, hint, 0, stream);
}

int psi_nn_index_query_mode(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                            const int *qsel, float *dist, int *idx, int *hint, int mode,
                            psi_stream_t stream) {
    return nn_index_query_impl(ix, q, q_bstride, B, n, qsel, dist, idx, hint, mode, stream);
}

int psi_nn_index_query(const psi_nn_index *ix, const float *q, long q_bstride, int B, int n,
                       const int *qsel, float *dist, int *idx, psi_stream_t stream) {
    return nn_index_query_impl(ix, q, q_bstride, B, n, qsel, dist, idx, nullptr, 0, stream);
}

}  // extern "C"
