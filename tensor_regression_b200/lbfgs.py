"""L-BFGS with strong-Wolfe line search on the flat device parameter vector — the optimizer behind
``fit`` (std:366,392 / mn:355,381).  Same algorithm, defaults, state carried across ``step`` calls and
stopping rules as ``torch.optim.LBFGS`` (torch/optim/lbfgs.py:333-536, ``_strong_wolfe`` 40-209,
``_cubic_interpolate`` 10-37), but the update history, the two-loop recursion and every vector
operation run in the library's kernels (``tr_lbfgs_direction`` / ``tr_lbfgs_point`` / ``tr_lbfgs_gtd``);
the host keeps only the scalar control flow and reads back four doubles per closure evaluation.
"""
import math

import torch


def _cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """lbfgs.py:10-37 on host floats."""
    if bounds is not None:
        xmin_bound, xmax_bound = bounds
    else:
        xmin_bound, xmax_bound = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    d2_square = d1 ** 2 - g1 * g2
    if d2_square >= 0:
        d2 = math.sqrt(d2_square)
        if x1 <= x2:
            min_pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            min_pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(min_pos, xmin_bound), xmax_bound)
    return (xmin_bound + xmax_bound) / 2.0


class LBFGS:
    """``closure(grad_out, loss_out)`` must evaluate the objective at the CURRENT ``theta``, write the flat
    gradient into ``grad_out`` (P, dtype) and ``[data loss, data loss + penalty]`` into ``loss_out`` (2
    device doubles) without synchronising; the optimizer reads ``loss_out[1]``."""

    def __init__(self, engine, theta, lr=1, max_iter=20, max_eval=None, tolerance_grad=1e-7,
                 tolerance_change=1e-9, history_size=100, line_search_fn=None):
        if not 0.0 <= lr:
            raise ValueError(f'Invalid learning rate: {lr}')
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        if line_search_fn is not None and line_search_fn != 'strong_wolfe':
            raise RuntimeError("only 'strong_wolfe' is supported")
        self.eng, self.theta = engine, theta
        self.lr, self.max_iter, self.max_eval = float(lr), int(max_iter), int(max_eval)
        self.tolerance_grad, self.tolerance_change = float(tolerance_grad), float(tolerance_change)
        self.history_size, self.line_search_fn = int(history_size), line_search_fn
        P, dev, dt = theta.numel(), theta.device, theta.dtype
        self.g = torch.empty(P, dtype=dt, device=dev)
        self.prev_g = torch.zeros(P, dtype=dt, device=dev)
        self.d = torch.zeros(P, dtype=dt, device=dev)
        self.S = torch.zeros((self.history_size, P), dtype=dt, device=dev)
        self.Y = torch.zeros((self.history_size, P), dtype=dt, device=dev)
        self.lstate = torch.zeros(4 + self.history_size, dtype=torch.float64, device=dev)
        self.scal = torch.zeros(8, dtype=torch.float64, device=dev)    # [0:2] losses, [2:4] gtd/gmax, [4:8] direction
        self.state = {'func_evals': 0, 'n_iter': 0, 't': None, 'prev_loss': None}

    # one closure evaluation at the current theta -> (loss, g.d, max|g|); one device->host read
    def _evaluate(self, closure, with_d):
        closure(self.g, self.scal[0:2])
        self.eng.lbfgs_gtd(self.g, self.d if with_d else None, self.scal[2:4])
        v = self.scal[0:4].tolist()
        return v[1], v[2], v[3]

    def _strong_wolfe(self, closure, x, t, f, gtd, d_norm, c1=1e-4, c2=0.9, max_ls=25):
        """lbfgs.py:40-209.  self.g holds g(x) on entry and the accepted point's gradient on return."""
        # torch.optim.LBFGS.step does not forward its tolerance_change to _strong_wolfe: the bracket-width exit
        # always uses the function's default of 1e-9 (lbfgs.py:40, call at 455-457)
        tol = 1e-9
        eng, theta, d = self.eng, self.theta, self.d

        def obj(tt):
            eng.lbfgs_point(theta, x, tt, d)
            fn, gtdn, _ = self._evaluate(closure, True)
            return fn, gtdn

        g0 = self.g.clone()
        f_new, gtd_new = obj(t)
        ls_func_evals = 1
        t_prev, f_prev, g_prev, gtd_prev = 0.0, f, g0, gtd
        done = False
        ls_iter = 0
        bracket = bracket_f = bracket_g = bracket_gtd = None
        while ls_iter < max_ls:
            if f_new > (f + c1 * t * gtd) or (ls_iter > 1 and f_new >= f_prev):
                bracket, bracket_f = [t_prev, t], [f_prev, f_new]
                bracket_g, bracket_gtd = [g_prev, self.g.clone()], [gtd_prev, gtd_new]
                break
            if abs(gtd_new) <= -c2 * gtd:
                bracket, bracket_f, bracket_g = [t], [f_new], [self.g.clone()]
                done = True
                break
            if gtd_new >= 0:
                bracket, bracket_f = [t_prev, t], [f_prev, f_new]
                bracket_g, bracket_gtd = [g_prev, self.g.clone()], [gtd_prev, gtd_new]
                break
            min_step = t + 0.01 * (t - t_prev)
            max_step = t * 10
            tmp = t
            t = _cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(min_step, max_step))
            t_prev, f_prev, g_prev, gtd_prev = tmp, f_new, self.g.clone(), gtd_new
            f_new, gtd_new = obj(t)
            ls_func_evals += 1
            ls_iter += 1
        if ls_iter == max_ls:
            bracket, bracket_f, bracket_g = [0.0, t], [f, f_new], [g0, self.g.clone()]
            bracket_gtd = [gtd, gtd_new]

        insuf_progress = False
        low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[-1] else (1, 0)
        while not done and ls_iter < max_ls:
            if abs(bracket[1] - bracket[0]) * d_norm < tol:
                break
            t = _cubic_interpolate(bracket[0], bracket_f[0], bracket_gtd[0], bracket[1], bracket_f[1], bracket_gtd[1])
            eps = 0.1 * (max(bracket) - min(bracket))
            if min(max(bracket) - t, t - min(bracket)) < eps:
                if insuf_progress or t >= max(bracket) or t <= min(bracket):
                    if abs(t - max(bracket)) < abs(t - min(bracket)):
                        t = max(bracket) - eps
                    else:
                        t = min(bracket) + eps
                    insuf_progress = False
                else:
                    insuf_progress = True
            else:
                insuf_progress = False
            f_new, gtd_new = obj(t)
            ls_func_evals += 1
            ls_iter += 1
            if f_new > (f + c1 * t * gtd) or f_new >= bracket_f[low_pos]:
                bracket[high_pos], bracket_f[high_pos] = t, f_new
                bracket_g[high_pos], bracket_gtd[high_pos] = self.g.clone(), gtd_new
                low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[1] else (1, 0)
            else:
                if abs(gtd_new) <= -c2 * gtd:
                    done = True
                elif gtd_new * (bracket[high_pos] - bracket[low_pos]) >= 0:
                    bracket[high_pos], bracket_f[high_pos] = bracket[low_pos], bracket_f[low_pos]
                    bracket_g[high_pos], bracket_gtd[high_pos] = bracket_g[low_pos], bracket_gtd[low_pos]
                bracket[low_pos], bracket_f[low_pos] = t, f_new
                bracket_g[low_pos], bracket_gtd[low_pos] = self.g.clone(), gtd_new
        t = bracket[low_pos]
        f_new = bracket_f[low_pos]
        self.g.copy_(bracket_g[low_pos])
        return f_new, t, ls_func_evals

    def step(self, closure):
        """lbfgs.py:333-536.  Returns the loss (float) at the entry point, like torch's ``orig_loss``."""
        st = self.state
        eng = self.eng
        orig_loss, _, gmax = self._evaluate(closure, False)
        loss = orig_loss
        current_evals = 1
        st['func_evals'] += 1
        if gmax <= self.tolerance_grad:
            return orig_loss
        t = st['t']
        n_iter = 0
        while n_iter < self.max_iter:
            n_iter += 1
            st['n_iter'] += 1
            first = st['n_iter'] == 1
            eng.lbfgs_direction(self.g, self.prev_g, self.d, 0.0 if t is None else t, first, self.S, self.Y,
                                self.lstate, self.history_size, self.scal[4:8])
            gtd, g1, gmax, dmax = self.scal[4:8].tolist()
            prev_loss = loss
            if first:
                t = min(1.0, 1.0 / g1) * self.lr
            else:
                t = self.lr
            if gtd > -self.tolerance_change:
                break
            ls_func_evals = 0
            if self.line_search_fn is not None:
                x_init = self.theta.clone()
                loss, t, ls_func_evals = self._strong_wolfe(closure, x_init, t, loss, gtd, dmax,
                                                            max_ls=self.max_eval - current_evals)
                eng.lbfgs_point(self.theta, x_init, t, self.d)
                eng.lbfgs_gtd(self.g, None, self.scal[2:4])
                gmax = self.scal[3].item()
                opt_cond = gmax <= self.tolerance_grad
            else:
                eng.lbfgs_point(self.theta, self.theta, t, self.d)
                opt_cond = False
                if n_iter != self.max_iter:
                    loss, _, gmax = self._evaluate(closure, False)
                    opt_cond = gmax <= self.tolerance_grad
                    ls_func_evals = 1
            current_evals += ls_func_evals
            st['func_evals'] += ls_func_evals
            if n_iter == self.max_iter:
                break
            if current_evals >= self.max_eval:
                break
            if opt_cond:
                break
            if dmax * abs(t) <= self.tolerance_change:
                break
            if abs(loss - prev_loss) < self.tolerance_change:
                break
        st['t'] = t
        st['prev_loss'] = prev_loss if n_iter else st['prev_loss']
        return orig_loss
