// L-BFGS with strong-Wolfe line search for the fused fitting loop, one independent optimiser PER BODY.
//
// Follows human_body_prior/optimizers/lbfgs_ls.py (a copy of pytorch PR #8824): _cubic_interpolate :24-51,
// _strong_Wolfe :54-183, LBFGS.step :275-463 -- the optimiser BASELINE.json's metric names ("300 L-BFGS iters")
// and SURVEY.md T7's optional second mode.  The reference drives it through a closure; here the loop evaluates ALL
// bodies once per graph iteration, so the algorithm is an explicit state machine that consumes one (loss,
// gradient) pair per body per iteration and answers with that body's next trial point -- every body sits at its
// own stage of its own line search.  `num_iter` of psi_fit_run counts closure evaluations.
//
// One warp per body runs the machine: lane l holds components l, l+32, l+64 of the 75-vector; dot products are
// FP32 products summed by warp shuffles (torch's dot is FP32 too), scalars (step length, losses, directional
// derivatives, curvature) are doubles as the reference's Python floats are.  State lives in global memory between
// iterations (psi_fit_ctx::lb_*).  oracle/lbfgs.py is the CPU statement of the same machine, itself pinned to
// torch.optim.LBFGS; tests compare the two evaluation by evaluation.
#pragma once

namespace psi {

constexpr int kLbStride = 96;                 // padded vector length (3 components per lane)
enum { LB_START = 0, LB_BRACKET = 1, LB_ZOOM = 2, LB_DONE = 3 };

struct LbfgsParams {
    double lr, tol_grad, tol_change, c1, c2;
    int history, max_iter, max_ls, zoom_max_iter, reset_lr;
};

// per-body scalar state
struct LbfgsScalars {
    double t, f0, gtd0, d_norm, t_prev, f_prev, gtd_prev, bt[2], bf[2], bgtd[2], H, t_acc, t_best, f_best;
    int phase, ls_iter, low, high, insuf, n_iter, evals, hist_count, hist_head, pad;
};

struct LbfgsState {           // device pointers, [B] rows each
    LbfgsScalars *sc;         // [B]
    float *x_init, *d, *g0, *g_prev, *bg;     // [B][96] (bg: [B][2][96])
    float *Y, *S;             // [B][history][96]
    double *ro;               // [B][history]
    float *xbest;             // [B][xdim] row-major: the last accepted point (+ the best Armijo step of a search in progress)
};

struct Vec3 {                 // this lane's three components
    float v[3];
};

__device__ __forceinline__ Vec3 lb_load(const float *p, int lane) {
    Vec3 r;
#pragma unroll
    for (int k = 0; k < 3; ++k) r.v[k] = p[lane + 32 * k];
    return r;
}
__device__ __forceinline__ void lb_store(float *p, int lane, const Vec3 &a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) p[lane + 32 * k] = a.v[k];
}
__device__ __forceinline__ double lb_dot(const Vec3 &a, const Vec3 &b) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) s = fmaf(a.v[k], b.v[k], s);
    return (double)warp_sum(s);
}
__device__ __forceinline__ double lb_abs_sum(const Vec3 &a) {
    return (double)warp_sum(fabsf(a.v[0]) + fabsf(a.v[1]) + fabsf(a.v[2]));
}
__device__ __forceinline__ double lb_abs_max(const Vec3 &a) {
    float m = fmaxf(fmaxf(fabsf(a.v[0]), fabsf(a.v[1])), fabsf(a.v[2]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    return (double)m;
}
__device__ __forceinline__ Vec3 lb_axpy(const Vec3 &x, double a, const Vec3 &y) {     // x + a*y
    Vec3 r;
    const float af = (float)a;
#pragma unroll
    for (int k = 0; k < 3; ++k) r.v[k] = fmaf(af, y.v[k], x.v[k]);
    return r;
}

// lbfgs_ls.py:24-51
__device__ inline double lb_cubic(double x1, double f1, double g1, double x2, double f2, double g2, bool bounded,
                                  double lo_b, double hi_b) {
    double xmin = lo_b, xmax = hi_b;
    if (!bounded) {
        xmin = x1 <= x2 ? x1 : x2;
        xmax = x1 <= x2 ? x2 : x1;
    }
    const double d1 = g1 + g2 - 3.0 * (f1 - f2) / (x1 - x2);
    const double d2s = d1 * d1 - g1 * g2;
    if (d2s >= 0.0) {
        const double d2 = sqrt(d2s);
        const double mp = x1 <= x2 ? x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2))
                                   : x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.0 * d2));
        return fmin(fmax(mp, xmin), xmax);
    }
    return (xmin + xmax) / 2.0;
}

// The machine of one body, executed by a full warp with warp-uniform control flow.  `g`, `f`: gradient and loss
// at the point handed out by the previous call; returns this lane's components of the next trial point in `x`.
struct LbfgsMachine {
    const LbfgsParams &P;
    LbfgsScalars s;            // registers; written back by lane 0 at the end
    float *x_init, *d, *g0, *g_prev, *bg, *Y, *S;
    double *ro;
    int lane;
    Vec3 vx_init, vd;          // cached

    __device__ Vec3 trial(double t) const { return lb_axpy(vx_init, t, vd); }

    // lbfgs_ls.py:332-389: curvature pair, two-loop recursion, first trial step of the next line search
    __device__ Vec3 new_direction(double f, const Vec3 &g, bool first) {
        s.n_iter += 1;
        Vec3 dn;
        if (first) {
#pragma unroll
            for (int k = 0; k < 3; ++k) dn.v[k] = -g.v[k];
            s.hist_count = 0; s.hist_head = 0; s.H = 1.0;
        } else {
            const Vec3 g0v = lb_load(g0, lane);
            Vec3 y, sv;
            const float ta = (float)s.t_acc;
#pragma unroll
            for (int k = 0; k < 3; ++k) { y.v[k] = g.v[k] - g0v.v[k]; sv.v[k] = vd.v[k] * ta; }
            const double ys = lb_dot(y, sv);
            if (ys > 1e-10) {
                int slot;
                if (s.hist_count == P.history) {            // drop the oldest pair
                    slot = s.hist_head;
                    s.hist_head = (s.hist_head + 1) % P.history;
                } else {
                    slot = (s.hist_head + s.hist_count) % P.history;
                    s.hist_count += 1;
                }
                lb_store(Y + (size_t)slot * kLbStride, lane, y);
                lb_store(S + (size_t)slot * kLbStride, lane, sv);
                if (lane == 0) ro[slot] = 1.0 / ys;
                __syncwarp();
                s.H = ys / lb_dot(y, y);
            }
            Vec3 q;
#pragma unroll
            for (int k = 0; k < 3; ++k) q.v[k] = -g.v[k];
            // al[i] of the reference lives in ro's second half: one double per pair
            double *al = ro + P.history;
            for (int i = s.hist_count - 1; i >= 0; --i) {
                const int slot = (s.hist_head + i) % P.history;
                const double a = lb_dot(lb_load(S + (size_t)slot * kLbStride, lane), q) * ro[slot];
                if (lane == 0) al[slot] = a;
                q = lb_axpy(q, -a, lb_load(Y + (size_t)slot * kLbStride, lane));
            }
            __syncwarp();
            const float Hf = (float)s.H;
#pragma unroll
            for (int k = 0; k < 3; ++k) q.v[k] *= Hf;
            for (int i = 0; i < s.hist_count; ++i) {
                const int slot = (s.hist_head + i) % P.history;
                const double be = lb_dot(lb_load(Y + (size_t)slot * kLbStride, lane), q) * ro[slot];
                q = lb_axpy(q, al[slot] - be, lb_load(S + (size_t)slot * kLbStride, lane));
            }
            dn = q;
        }
        vd = dn;
        lb_store(d, lane, dn);
        lb_store(g0, lane, g);                                  // prev_flat_grad
        s.f0 = f;                                               // prev_loss
        const double t = (s.n_iter == 1 && P.reset_lr) ? fmin(1.0, 1.0 / lb_abs_sum(g)) * P.lr : P.lr;
        s.gtd0 = lb_dot(g, dn);
        if (s.gtd0 > -P.tol_change) { s.phase = LB_DONE; return vx_init; }
        s.d_norm = lb_abs_max(dn);
        s.t = t; s.t_prev = 0.0; s.f_prev = f; s.gtd_prev = s.gtd0;
        lb_store(g_prev, lane, g);
        s.ls_iter = 0; s.phase = LB_BRACKET;
        s.t_best = 0.0; s.f_best = f;
        return trial(t);
    }

    // lbfgs_ls.py:400-455: accept the step, test the stopping rules, start the next outer iteration
    __device__ Vec3 finish(double t, double f, const Vec3 &g) {
        s.t_acc = t;
        vx_init = trial(t);
        lb_store(x_init, lane, vx_init);
        Vec3 dt;
        const float tf = (float)t;
#pragma unroll
        for (int k = 0; k < 3; ++k) dt.v[k] = vd.v[k] * tf;
        const bool stop = s.n_iter == P.max_iter || lb_abs_sum(g) <= P.tol_grad || s.gtd0 > -P.tol_change ||
                          lb_abs_sum(dt) <= P.tol_change || fabs(f - s.f0) < P.tol_change;
        s.t_best = 0.0; s.f_best = f;
        if (stop) { s.phase = LB_DONE; return vx_init; }
        return new_direction(f, g, false);
    }

    __device__ Vec3 zoom_setup(double t0, double f0b, const Vec3 &ga, double gtd_a, double t1, double f1b, const Vec3 &gb,
                               double gtd_b) {
        s.bt[0] = t0; s.bf[0] = f0b; s.bgtd[0] = gtd_a;
        s.bt[1] = t1; s.bf[1] = f1b; s.bgtd[1] = gtd_b;
        lb_store(bg, lane, ga);
        lb_store(bg + kLbStride, lane, gb);
        s.insuf = 0;
        s.low = s.bf[0] <= s.bf[1] ? 0 : 1;
        s.high = 1 - s.low;
        s.phase = LB_ZOOM;
        return zoom_trial();
    }

    // top half of the zoom loop (lbfgs_ls.py:128-148): the next trial step, or the end of the search
    __device__ Vec3 zoom_trial() {
        if (s.ls_iter >= P.zoom_max_iter) return finish(s.bt[s.low], s.bf[s.low], lb_load(bg + s.low * kLbStride, lane));
        double t = lb_cubic(s.bt[0], s.bf[0], s.bgtd[0], s.bt[1], s.bf[1], s.bgtd[1], false, 0.0, 0.0);
        const double hi = fmax(s.bt[0], s.bt[1]), lo = fmin(s.bt[0], s.bt[1]);
        const double eps = 0.1 * (hi - lo);
        if (fmin(hi - t, t - lo) < eps) {
            if (s.insuf || t >= hi || t <= lo) {
                t = fabs(t - hi) < fabs(t - lo) ? hi - eps : lo + eps;
                s.insuf = 0;
            } else {
                s.insuf = 1;
            }
        } else {
            s.insuf = 0;
        }
        s.t = t;
        return trial(t);
    }

    __device__ Vec3 feed(double f, const Vec3 &g) {
        s.evals += 1;
        if (s.phase == LB_DONE) return vx_init;
        if (s.phase == LB_START) {
            if (lb_abs_sum(g) <= P.tol_grad) { s.phase = LB_DONE; return vx_init; }
            return new_direction(f, g, true);
        }
        const double gtd_new = lb_dot(g, vd);
        const double t = s.t;
        if (f <= s.f0 + P.c1 * t * s.gtd0 && f < s.f_best) { s.f_best = f; s.t_best = t; }
        if (s.phase == LB_BRACKET) {                            // one pass of lbfgs_ls.py:70-113
            const Vec3 gp = lb_load(g_prev, lane);
            if (s.ls_iter == P.max_ls)                          // :115-120
                return zoom_setup(0.0, s.f0, lb_load(g0, lane), s.gtd0, t, f, g, gtd_new);
            if (f > s.f0 + P.c1 * t * s.gtd0 || (s.ls_iter > 1 && f >= s.f_prev))
                return zoom_setup(s.t_prev, s.f_prev, gp, s.gtd_prev, t, f, g, gtd_new);
            if (fabs(gtd_new) <= -P.c2 * s.gtd0) return finish(t, f, g);
            if (gtd_new >= 0.0) return zoom_setup(s.t_prev, s.f_prev, gp, s.gtd_prev, t, f, g, gtd_new);
            const double tn = lb_cubic(s.t_prev, s.f_prev, s.gtd_prev, t, f, gtd_new, true, t + 0.01 * (t - s.t_prev), t * 10.0);
            s.t_prev = t; s.f_prev = f; s.gtd_prev = gtd_new;
            lb_store(g_prev, lane, g);
            s.t = tn;
            s.ls_iter += 1;
            return trial(tn);
        }
        // LB_ZOOM: bottom half of the zoom loop (lbfgs_ls.py:150-177)
        s.ls_iter += 1;
        bool done = false;
        if (f > s.f0 + P.c1 * t * s.gtd0 || f >= s.bf[s.low]) {
            const int h = s.high;
            s.bt[h] = t; s.bf[h] = f; s.bgtd[h] = gtd_new;
            lb_store(bg + h * kLbStride, lane, g);
            s.low = s.bf[0] <= s.bf[1] ? 0 : 1;
            s.high = 1 - s.low;
        } else {
            if (fabs(gtd_new) <= -P.c2 * s.gtd0) {
                done = true;
            } else if (gtd_new * (s.bt[s.high] - s.bt[s.low]) >= 0.0) {
                const int h = s.high, l = s.low;
                s.bt[h] = s.bt[l]; s.bf[h] = s.bf[l]; s.bgtd[h] = s.bgtd[l];
                lb_store(bg + h * kLbStride, lane, lb_load(bg + l * kLbStride, lane));
            }
            const int l = s.low;
            s.bt[l] = t; s.bf[l] = f; s.bgtd[l] = gtd_new;
            lb_store(bg + l * kLbStride, lane, g);
        }
        __syncwarp();
        if (done || fabs(s.bt[1] - s.bt[0]) * s.d_norm < P.tol_change)
            return finish(s.bt[s.low], s.bf[s.low], lb_load(bg + s.low * kLbStride, lane));
        return zoom_trial();
    }
};

// One evaluation of body b: (f, gradient in g_smem[0..xdim)) -> next trial point in x_smem[0..xdim) and global x / xbest.
// Whole warp; components >= xdim are zero.
__device__ inline void lbfgs_feed_body(const LbfgsParams &P, const LbfgsState &st, int b, int xdim, double f,
                                       const float *g_smem, float *x_smem, float *x_glob, int lane) {
    LbfgsScalars *sp = st.sc + b;
    LbfgsMachine m{P, *sp,
                   st.x_init + (size_t)b * kLbStride, st.d + (size_t)b * kLbStride, st.g0 + (size_t)b * kLbStride,
                   st.g_prev + (size_t)b * kLbStride, st.bg + (size_t)b * 2 * kLbStride,
                   st.Y + (size_t)b * P.history * kLbStride, st.S + (size_t)b * P.history * kLbStride,
                   st.ro + (size_t)b * 2 * P.history, lane, Vec3(), Vec3()};
    Vec3 g;
#pragma unroll
    for (int k = 0; k < 3; ++k) g.v[k] = (lane + 32 * k < xdim) ? g_smem[lane + 32 * k] : 0.f;
    if (m.s.phase == LB_START) {           // x_init = the starting point (x_smem holds it)
#pragma unroll
        for (int k = 0; k < 3; ++k) m.vx_init.v[k] = (lane + 32 * k < xdim) ? x_smem[lane + 32 * k] : 0.f;
        lb_store(m.x_init, lane, m.vx_init);
#pragma unroll
        for (int k = 0; k < 3; ++k) m.vd.v[k] = 0.f;
    } else {
        m.vx_init = lb_load(m.x_init, lane);
        m.vd = lb_load(m.d, lane);
    }
    const Vec3 xn = m.feed(f, g);
    // the best point so far: the last accepted one plus the best sufficient-decrease step of the running search
    const Vec3 xb = lb_axpy(m.vx_init, m.s.phase == LB_DONE ? 0.0 : m.s.t_best, m.vd);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = lane + 32 * k;
        if (e < xdim) {
            x_smem[e] = xn.v[k];
            x_glob[(size_t)b * xdim + e] = xn.v[k];
            st.xbest[(size_t)b * xdim + e] = xb.v[k];
        }
    }
    if (lane == 0) *sp = m.s;
}

}  // namespace psi
