"""Host control flow of tensor_regression_b200.lbfgs.LBFGS (strong-Wolfe line search, stopping rules,
state carried across step() calls) against torch.optim.LBFGS on CPU.  The three vector kernels are
replaced by a torch stand-in with the semantics documented in include/tr_b200.h, so this checks the
port of torch/optim/lbfgs.py, not the CUDA code (tests/test_gpu_parity.py does that)."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_lbfgs():
    # the package __init__ loads the CUDA library; lbfgs.py itself is pure host logic
    spec = importlib.util.spec_from_file_location('_trb200_lbfgs', os.path.join(ROOT, 'tensor_regression_b200', 'lbfgs.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeVectorEngine:
    """torch restatement of tr_lbfgs_direction / tr_lbfgs_point / tr_lbfgs_gtd."""

    def lbfgs_direction(self, g, prev_g, d, t, first, S, Y, ls, hist, scal):
        if first:
            d.copy_(-g)
            ls[0], ls[1], ls[2] = 1.0, 0.0, 0.0
        else:
            num_old, head = int(ls[1]), int(ls[2])
            y = g - prev_g
            s = d * t
            ys = float(y.dot(s))
            if ys > 1e-10:
                if num_old == hist:
                    slot = head
                    head = (head + 1) % hist
                else:
                    slot = (head + num_old) % hist
                    num_old += 1
                Y[slot].copy_(y)
                S[slot].copy_(s)
                ls[4 + slot] = 1.0 / ys
                ls[0] = ys / float(y.dot(y))
            q = -g.clone()
            al = [0.0] * num_old
            for i in range(num_old - 1, -1, -1):
                j = (head + i) % hist
                al[i] = float(S[j].dot(q)) * float(ls[4 + j])
                q.add_(Y[j], alpha=-al[i])
            r = q * float(ls[0])
            for i in range(num_old):
                j = (head + i) % hist
                be = float(Y[j].dot(r)) * float(ls[4 + j])
                r.add_(S[j], alpha=al[i] - be)
            d.copy_(r)
            ls[1], ls[2] = float(num_old), float(head)
        prev_g.copy_(g)
        scal[0], scal[1], scal[2], scal[3] = float(g.dot(d)), float(g.abs().sum()), float(g.abs().max()), float(d.abs().max())

    def lbfgs_point(self, out, x, t, d):
        out.copy_(x + t * d)

    def lbfgs_gtd(self, g, d, scal):
        scal[0] = float(g.dot(d)) if d is not None else 0.0
        scal[1] = float(g.abs().max())


def rosenbrock_like(x):
    return ((1 - x[:-1]) ** 2).sum() + 100 * ((x[1:] - x[:-1] ** 2) ** 2).sum() + 0.1 * torch.sqrt((x ** 2).sum())


CASES = [
    {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
     'history_size': 100, 'line_search_fn': 'strong_wolfe'},
    {'lr': 1, 'max_iter': 20, 'max_eval': 20, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
     'history_size': 3, 'line_search_fn': 'strong_wolfe'},
    {'lr': 0.05, 'max_iter': 7, 'history_size': 5, 'line_search_fn': None},
    # tolerance_change != 1e-9: torch does NOT forward it to _strong_wolfe (the bracket-width exit keeps its
    # default 1e-9), only the outer stopping rules use it
    {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-06,
     'history_size': 100, 'line_search_fn': 'strong_wolfe'},
    {'lr': 1, 'max_iter': 30, 'max_eval': 60, 'tolerance_grad': 1e-09, 'tolerance_change': 1e-03,
     'history_size': 10, 'line_search_fn': 'strong_wolfe'},
]


@pytest.mark.parametrize('kw', CASES, ids=['reference_kwargs', 'short_history', 'fixed_step', 'tolchange_1e-6', 'tolchange_1e-3'])
def test_native_lbfgs_matches_torch_lbfgs(kw):
    L = _load_lbfgs()
    torch.manual_seed(3)
    x0 = torch.randn(12, dtype=torch.float64) * 0.5
    # torch
    xt = x0.clone().requires_grad_(True)
    opt = torch.optim.LBFGS([xt], **kw)
    ref_losses = []

    def tclosure():
        opt.zero_grad()
        loss = rosenbrock_like(xt)
        loss.backward()
        return loss

    for _ in range(6):
        ref_losses.append(float(opt.step(tclosure)))
    # ours
    theta = x0.clone()
    mine = L.LBFGS(FakeVectorEngine(), theta, **kw)

    def closure(grad_out, loss_out):
        xv = theta.clone().requires_grad_(True)
        loss = rosenbrock_like(xv)
        loss.backward()
        grad_out.copy_(xv.grad)
        loss_out[0] = loss.item()
        loss_out[1] = loss.item()

    got = [mine.step(closure) for _ in range(6)]
    assert max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(got, ref_losses)) < 1e-9, (got, ref_losses)
    assert torch.allclose(theta, xt.detach(), rtol=1e-8, atol=1e-10)
    assert mine.state['func_evals'] == opt.state[xt]['func_evals']
    assert mine.state['n_iter'] == opt.state[xt]['n_iter']
