"""CPU ORACLE of the L-BFGS / strong-Wolfe optimiser -- TEST INFRASTRUCTURE, not a product path.

Restates human_body_prior/optimizers/lbfgs_ls.py (a copy of pytorch PR #8824, optimizers/readme:1):
`_cubic_interpolate` :24-51, `_strong_Wolfe` :54-183, `LBFGS.step` :275-463 -- as an explicit STATE MACHINE that
consumes ONE closure evaluation (loss, gradient) per call and answers with the next point to evaluate.  That is
the form the fused fitting loop needs (one evaluation of all bodies per graph iteration, every body at its own
stage of its own line search); `psi-release_b200/csrc/fit_lbfgs.cuh` is the device version of the same machine
and is tested against this file.  Scalars (step length, losses, directional derivatives) are Python floats
(double), vectors keep the dtype they arrive in.

Pin: with `reset_lr=True` (first step min(1, 1/|g|_1) * lr, lbfgs_ls.py:374-375) and max_iter <= 25 this file
must reproduce `torch.optim.LBFGS(line_search_fn='strong_wolfe')` -- the merged form of the same PR -- step
for step (tests/test_oracle.py, on the Rosenbrock function the reference file defines at :15-18).

Differences from the reference file, all deliberate: `abs_grad_sum` is recomputed after every line search
(the file tests a value from before the loop, :440); the evaluation budget is counted in closure calls.
"""
from __future__ import annotations

import numpy as np

START, BRACKET, ZOOM, DONE = 0, 1, 2, 3


def cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """lbfgs_ls.py:24-51."""
    if bounds is not None:
        xmin_bound, xmax_bound = bounds
    else:
        xmin_bound, xmax_bound = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    d2_square = d1 ** 2 - g1 * g2
    if d2_square >= 0:
        d2 = np.sqrt(d2_square)
        if x1 <= x2:
            min_pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            min_pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(min_pos, xmin_bound), xmax_bound)
    return (xmin_bound + xmax_bound) / 2.0


class LBFGSMachine:
    """One optimisation problem.  `x = m.start(x0)`; then repeatedly `x = m.feed(loss(x), grad(x))`.
    `m.best()` is the last ACCEPTED point (plus the best Armijo point of a line search in progress)."""

    def __init__(self, lr=1.0, history_size=100, tolerance_grad=1e-5, tolerance_change=1e-9, max_iter=300,
                 max_ls=25, c1=1e-4, c2=0.9, reset_lr=False, zoom_max_iter=None):
        self.lr, self.m, self.tg, self.tc = float(lr), int(history_size), float(tolerance_grad), float(tolerance_change)
        self.max_iter, self.max_ls, self.c1, self.c2, self.reset_lr = int(max_iter), int(max_ls), c1, c2, reset_lr
        self.zoom_max_iter = int(max_iter if zoom_max_iter is None else zoom_max_iter)      # lbfgs_ls.py:400: max_iter=max_iter
        self.phase = START

    # ------------------------------------------------------------------------------------------ API
    def start(self, x0):
        self.x_init = np.array(x0, copy=True)
        self.x = self.x_init.copy()
        self.n_iter, self.evals, self.phase = 0, 0, START
        self.t_best, self.f_best = 0.0, float("inf")
        self.Y, self.S, self.ro = [], [], []
        self.H = 1.0
        return self.x

    def best(self):
        """The last accepted point, or -- inside a line search -- its best evaluated sufficient-decrease step."""
        if self.phase in (BRACKET, ZOOM) and self.t_best != 0.0:
            return self.x_init + self.x_init.dtype.type(self.t_best) * self.d
        return self.x_init

    def feed(self, f, g):
        """(loss, gradient) at the point the previous call returned -> the next point to evaluate."""
        f = float(f)
        g = np.array(g, copy=True)
        self.evals += 1
        if self.phase == DONE:
            return self.x
        if self.phase == START:
            if float(np.abs(g).sum()) <= self.tg:                    # lbfgs_ls.py:318-319
                self.phase = DONE
                return self.x
            return self._new_direction(f, g, first=True)
        gtd_new = float(np.dot(g, self.d))
        if f <= self.f0 + self.c1 * self.t * self.gtd0 and f < self.f_best:
            self.f_best, self.t_best = f, self.t
        if self.phase == BRACKET:
            return self._bracket(f, g, gtd_new)
        return self._zoom(f, g, gtd_new)

    # ------------------------------------------------------------------------------- outer iteration
    def _new_direction(self, f, g, first):
        """lbfgs_ls.py:332-389: curvature pair, two-loop recursion, first trial step of the line search."""
        self.n_iter += 1
        if first:
            d = -g
            self.Y, self.S, self.ro, self.H = [], [], [], 1.0
        else:
            y = g - self.g0
            s = self.d * self.x_init.dtype.type(self.t_acc)
            ys = float(np.dot(y, s))
            if ys > 1e-10:
                if len(self.Y) == self.m:
                    self.Y.pop(0); self.S.pop(0); self.ro.pop(0)
                self.Y.append(y); self.S.append(s); self.ro.append(1.0 / ys)
                self.H = ys / float(np.dot(y, y))
            q = -g
            al = [0.0] * len(self.Y)
            for i in range(len(self.Y) - 1, -1, -1):
                al[i] = float(np.dot(self.S[i], q)) * self.ro[i]
                q = q - q.dtype.type(al[i]) * self.Y[i]
            r = q * q.dtype.type(self.H)
            for i in range(len(self.Y)):
                be = float(np.dot(self.Y[i], r)) * self.ro[i]
                r = r + r.dtype.type(al[i] - be) * self.S[i]
            d = r
        self.d = d
        self.g0, self.f0 = g, f                                      # prev_flat_grad / prev_loss (:366-370)
        t = min(1.0, 1.0 / float(np.abs(g).sum())) * self.lr if (self.n_iter == 1 and self.reset_lr) else self.lr
        self.gtd0 = float(np.dot(g, d))
        if self.gtd0 > -self.tc:                                     # :381-382
            self.phase = DONE
            return self.x
        # _strong_Wolfe prologue (:58-68)
        self.d_norm = float(np.abs(d).max())
        self.t, self.t_prev, self.f_prev, self.g_prev, self.gtd_prev = t, 0.0, f, g, self.gtd0
        self.ls_iter, self.phase = 0, BRACKET
        self.t_best, self.f_best = 0.0, f
        self.x = self.x_init + self.x_init.dtype.type(t) * d
        return self.x

    def _finish_line_search(self, t, f, g):
        """lbfgs_ls.py:400-455: accept the step, test the stopping rules, start the next outer iteration."""
        self.t_acc = t
        self.x_init = self.x_init + self.x_init.dtype.type(t) * self.d
        self.x = self.x_init.copy()
        stop = (self.n_iter == self.max_iter or float(np.abs(g).sum()) <= self.tg or self.gtd0 > -self.tc
                or float(np.abs(self.d * self.d.dtype.type(t)).sum()) <= self.tc or abs(f - self.f0) < self.tc)
        if stop:
            self.phase = DONE
            return self.x
        return self._new_direction(f, g, first=False)

    # ----------------------------------------------------------------------------------- line search
    def _bracket(self, f_new, g_new, gtd_new):
        """One pass of the bracketing loop (lbfgs_ls.py:70-113) after an evaluation at self.t."""
        t = self.t
        if self.ls_iter == self.max_ls:                              # :115-120
            return self._zoom_setup([0.0, t], [self.f0, f_new], [self.g0, g_new], [self.gtd0, gtd_new])
        if f_new > (self.f0 + self.c1 * t * self.gtd0) or (self.ls_iter > 1 and f_new >= self.f_prev):
            return self._zoom_setup([self.t_prev, t], [self.f_prev, f_new], [self.g_prev, g_new], [self.gtd_prev, gtd_new])
        if abs(gtd_new) <= -self.c2 * self.gtd0:
            return self._finish_line_search(t, f_new, g_new)
        if gtd_new >= 0:
            return self._zoom_setup([self.t_prev, t], [self.f_prev, f_new], [self.g_prev, g_new], [self.gtd_prev, gtd_new])
        min_step, max_step = t + 0.01 * (t - self.t_prev), t * 10
        t_new = cubic_interpolate(self.t_prev, self.f_prev, self.gtd_prev, t, f_new, gtd_new, bounds=(min_step, max_step))
        self.t_prev, self.f_prev, self.g_prev, self.gtd_prev = t, f_new, g_new, gtd_new
        self.t = float(t_new)
        self.ls_iter += 1
        self.x = self.x_init + self.x_init.dtype.type(self.t) * self.d
        return self.x

    def _zoom_setup(self, bt, bf, bg, bgtd):
        self.bt, self.bf, self.bg, self.bgtd = bt, bf, bg, bgtd
        self.insuf = False
        self.low, self.high = (0, 1) if bf[0] <= bf[1] else (1, 0)   # :127
        self.phase = ZOOM
        return self._zoom_trial()

    def _zoom_trial(self):
        """Top half of the zoom loop (lbfgs_ls.py:128-148): the next trial step, or the end of the search."""
        if self.ls_iter >= self.zoom_max_iter:
            return self._finish_line_search(self.bt[self.low], self.bf[self.low], self.bg[self.low])
        bt = self.bt
        t = float(cubic_interpolate(bt[0], self.bf[0], self.bgtd[0], bt[1], self.bf[1], self.bgtd[1]))
        hi, lo = max(bt), min(bt)
        eps = 0.1 * (hi - lo)
        if min(hi - t, t - lo) < eps:
            if self.insuf or t >= hi or t <= lo:
                t = hi - eps if abs(t - hi) < abs(t - lo) else lo + eps
                self.insuf = False
            else:
                self.insuf = True
        else:
            self.insuf = False
        self.t = t
        self.x = self.x_init + self.x_init.dtype.type(t) * self.d
        return self.x

    def _zoom(self, f_new, g_new, gtd_new):
        """Bottom half of the zoom loop (lbfgs_ls.py:150-177) after an evaluation at self.t."""
        t = self.t
        self.ls_iter += 1
        done = False
        if f_new > (self.f0 + self.c1 * t * self.gtd0) or f_new >= self.bf[self.low]:
            h = self.high
            self.bt[h], self.bf[h], self.bg[h], self.bgtd[h] = t, f_new, g_new, gtd_new
            self.low, self.high = (0, 1) if self.bf[0] <= self.bf[1] else (1, 0)
        else:
            if abs(gtd_new) <= -self.c2 * self.gtd0:
                done = True
            elif gtd_new * (self.bt[self.high] - self.bt[self.low]) >= 0:
                h, l = self.high, self.low
                self.bt[h], self.bf[h], self.bg[h], self.bgtd[h] = self.bt[l], self.bf[l], self.bg[l], self.bgtd[l]
            l = self.low
            self.bt[l], self.bf[l], self.bg[l], self.bgtd[l] = t, f_new, g_new, gtd_new
        if done or abs(self.bt[1] - self.bt[0]) * self.d_norm < self.tc:
            return self._finish_line_search(self.bt[self.low], self.bf[self.low], self.bg[self.low])
        return self._zoom_trial()


def minimize(fun, x0, num_evals, **kw):
    """fun(x) -> (loss, grad).  Runs `num_evals` closure evaluations; returns (best point, machine)."""
    m = LBFGSMachine(**kw)
    x = m.start(x0)
    for _ in range(num_evals):
        f, g = fun(x)
        x = m.feed(f, g)
    return m.best(), m
