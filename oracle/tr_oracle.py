"""CPU restatement of the reference's CP tensor-regression fit iteration.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Two independent layers:

* Part 1 — ``port``: the reference's algorithm op for op in torch on the CPU
  (materialise the dense coefficient tensor with a Khatri-Rao chain, one
  ``X_flat @ B_flat`` matmul, torch autograd for the gradients, ``torch.optim``
  for the update).  This is what ``bench.py`` times as the CPU baseline when
  /root/reference is not on the box, because it does the same CPU work as the
  reference.  Follows
    standard_tensor_regression.py:53-85 (non_neg_fn), 87-130 (lin_model),
      180-196 (L2_penalty), 368-375 / 459-463 (loss, backward, step)
    multinomial_tensor_regression.py:116-146, 148-187 (model), 189-204,
      357-366 / 454-458 (class-weighted CE on the already-softmaxed output).
  tensorly (not vendored, version unpinned) is restated in
  oracle/tensorly_standin.py from its published semantics.

* Part 2 — ``closed form``: SURVEY.md Appendix A written out with explicit
  contractions and no autograd, producing the *unnormalised local sums* in the
  packed layout the C-ABI kernels emit (``gradsum``), plus the finishing step
  (normalisation, softplus chain rule, penalty gradient) and the Adam update
  (torch/optim/adam.py single-tensor path).  Part 2 is validated against Part 1's
  autograd in tests/test_oracle_pinned.py; the CUDA kernels are compared with both.

Parity pinning: Part 1 == unmodified reference run through the stand-in
(tests/golden/*.npz made by oracle/make_golden.py) and reproduces the notebook
known-answer loss 0.041904340578888165 (demo_TensorRegression.ipynb cell 8).
"""
import math

import numpy as np
import torch

from .tensorly_standin import cp_to_tensor, inner

DEFAULT_SOFTPLUS = {'beta': 50, 'threshold': 1}


# ----------------------------------------------------------------------------
# Part 1: reference port (torch ops + autograd, CPU)
# ----------------------------------------------------------------------------

def non_neg(Bcp, non_negative, softplus_kwargs=None):
    """std:53-85 / mn:116-146 — softplus on flagged list positions."""
    if softplus_kwargs is None:
        softplus_kwargs = DEFAULT_SOFTPLUS
    out = []
    for ii in range(len(Bcp)):
        if non_negative[ii]:
            out.append(torch.nn.functional.softplus(Bcp[ii], **softplus_kwargs))
        else:
            out.append(Bcp[ii])
    return out


def lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs=None):
    """std:123-130."""
    B = cp_to_tensor((weights, non_neg(Bcp, non_negative, softplus_kwargs)))[..., None]
    return inner(X, B, n_modes=len(Bcp)).squeeze() + bias


def mn_model(X, Bcp, weights, non_negative, softplus_kwargs=None):
    """mn:180-187 — the last factor is the class factor; returns softmax probabilities."""
    B = cp_to_tensor((weights, non_neg(Bcp, non_negative, softplus_kwargs)))
    return torch.nn.functional.softmax(inner(X, B, n_modes=len(Bcp) - 1), dim=1)


def L2_penalty(Bcp):
    """std:193-196 / mn:201-204 — sum of UN-squared Frobenius norms of the raw factors."""
    tot = 0
    for comp in Bcp:
        tot = tot + torch.sqrt(torch.sum(comp ** 2))
    return tot


def _leaf(ts):
    return [t.detach().clone().requires_grad_(True) for t in ts]


def std_loss_grad(X, y, Bcp, bias, weights, non_negative, lambda_L2, softplus_kwargs=None):
    """One closure evaluation of std:368-373 / 459-462.  Returns python dict."""
    Bcp = _leaf(Bcp)
    bias = bias.detach().clone().requires_grad_(True)
    y_hat = lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs)
    mse = torch.nn.MSELoss()(y_hat, y)
    loss = mse + lambda_L2 * L2_penalty(Bcp)
    loss.backward()
    return {'y_hat': y_hat.detach(), 'loss_data': mse.detach(), 'loss': loss.detach(),
            'grads': [b.grad.detach() for b in Bcp], 'dbias': bias.grad.detach()}


def mn_loss_grad(X, y, Bcp, weights, non_negative, class_weights, lambda_L2, softplus_kwargs=None):
    """One closure evaluation of mn:357-362 / 454-457 (CrossEntropyLoss on probabilities)."""
    Bcp = _leaf(Bcp)
    P = mn_model(X, Bcp, weights, non_negative, softplus_kwargs)
    loss_fn = torch.nn.CrossEntropyLoss(weight=torch.as_tensor(class_weights, dtype=P.dtype))
    ce = loss_fn(P, y)
    loss = ce + lambda_L2 * L2_penalty(Bcp)
    loss.backward()
    return {'P': P.detach(), 'loss_data': ce.detach(), 'loss': loss.detach(),
            'grads': [b.grad.detach() for b in Bcp]}


def fit_adam_std(X, y, Bcp, bias, weights, non_negative, lambda_L2, n_iter, adam_kwargs, softplus_kwargs=None):
    """std:453-464 for exactly n_iter iterations (no convergence test)."""
    Bcp = _leaf(Bcp)
    bias = bias.detach().clone().requires_grad_(True)
    opt = torch.optim.Adam(Bcp + [bias], **adam_kwargs)
    loss_fn = torch.nn.MSELoss()
    losses = []
    for _ in range(n_iter):
        opt.zero_grad()
        y_hat = lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs)
        loss = loss_fn(y_hat, y) + lambda_L2 * L2_penalty(Bcp)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return {'Bcp': [b.detach() for b in Bcp], 'bias': bias.detach(), 'loss_running': losses}


def fit_adam_mn(X, y, Bcp, weights, non_negative, class_weights, lambda_L2, n_iter, adam_kwargs,
                softplus_kwargs=None):
    """mn:447-459 for exactly n_iter iterations."""
    Bcp = _leaf(Bcp)
    opt = torch.optim.Adam(Bcp, **adam_kwargs)
    loss_fn = torch.nn.CrossEntropyLoss(weight=torch.as_tensor(class_weights, dtype=torch.float32))
    losses = []
    for _ in range(n_iter):
        opt.zero_grad()
        P = mn_model(X, Bcp, weights, non_negative, softplus_kwargs)
        loss = loss_fn(P, y) + lambda_L2 * L2_penalty(Bcp)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return {'Bcp': [b.detach() for b in Bcp], 'loss_running': losses}


def fit_lbfgs_std(X, y, Bcp, bias, weights, non_negative, lambda_L2, max_iter, tol, patience,
                  lbfgs_kwargs, softplus_kwargs=None, running_loss_logging_interval=1):
    """std:366-392 (L-BFGS outer loop with the reference's logging + convergence rule)."""
    Bcp = _leaf(Bcp)
    bias = bias.detach().clone().requires_grad_(True)
    opt = torch.optim.LBFGS(Bcp + [bias], **lbfgs_kwargs)
    loss_fn = torch.nn.MSELoss()

    def closure():
        opt.zero_grad()
        y_hat = lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs)
        loss = loss_fn(y_hat, y) + lambda_L2 * L2_penalty(Bcp)
        loss.backward()
        return loss

    losses, converged = [], False
    for ii in range(max_iter):
        if ii % running_loss_logging_interval == 0:
            with torch.no_grad():
                y_hat = lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs)
                losses.append(loss_fn(y_hat, y).item())
        if ii > patience:
            if np.sum(np.abs(np.diff(losses[ii - patience:]))) < tol:
                converged = True
                break
        opt.step(closure)
    return {'Bcp': [b.detach() for b in Bcp], 'bias': bias.detach(), 'loss_running': losses,
            'converged': converged}


def fit_lbfgs_mn(X, y, Bcp, weights, non_negative, class_weights, lambda_L2, max_iter, tol, patience,
                 lbfgs_kwargs, softplus_kwargs=None, running_loss_logging_interval=1):
    """mn:355-381 (L-BFGS outer loop, CE on probabilities, logging without the penalty)."""
    Bcp = _leaf(Bcp)
    opt = torch.optim.LBFGS(Bcp, **lbfgs_kwargs)
    loss_fn = torch.nn.CrossEntropyLoss(weight=torch.as_tensor(class_weights, dtype=X.dtype))

    def closure():
        opt.zero_grad()
        P = mn_model(X, Bcp, weights, non_negative, softplus_kwargs)
        loss = loss_fn(P, y) + lambda_L2 * L2_penalty(Bcp)
        loss.backward()
        return loss

    losses, converged = [], False
    for ii in range(max_iter):
        if ii % running_loss_logging_interval == 0:
            with torch.no_grad():
                losses.append(loss_fn(mn_model(X, Bcp, weights, non_negative, softplus_kwargs), y).item())
        if ii > patience:
            if np.sum(np.abs(np.diff(losses[ii - patience:]))) < tol:
                converged = True
                break
        opt.step(closure)
    return {'Bcp': [b.detach() for b in Bcp], 'loss_running': losses, 'converged': converged}


# ----------------------------------------------------------------------------
# Part 2: closed form, no autograd, packed layout
# ----------------------------------------------------------------------------

def softplus_fwd(x, beta, thr):
    """torch.nn.functional.softplus: x if x*beta > thr else log1p(exp(x*beta))/beta."""
    bx = x * beta
    return torch.where(bx > thr, x, torch.log1p(torch.exp(torch.clamp(bx, max=thr))) / beta)


def softplus_grad(x, beta, thr):
    bx = x * beta
    return torch.where(bx > thr, torch.ones_like(x), torch.sigmoid(bx))


def tilde(Bcp, non_negative, softplus_kwargs=None):
    sp = softplus_kwargs or DEFAULT_SOFTPLUS
    return [softplus_fwd(F, sp['beta'], sp['threshold']) if non_negative[m] else F for m, F in enumerate(Bcp)]


def inner_u(X, Ft):
    """u[n,r] = sum_i X[n,i_1..i_k] prod_m Ft[m][i_m,r]; innermost mode contracted first."""
    k = len(Ft)
    R = Ft[0].shape[1]
    N = X.shape[0]
    T = torch.einsum('...j,jr->...r', X, Ft[k - 1])  # (N, I_1..I_{k-1}, R)
    for m in range(k - 2, -1, -1):
        T = torch.einsum('...jr,jr->...r', T, Ft[m])
    return T.reshape(N, R)


def mttkrp_all(G, Ft):
    """For a (I_1..I_k) tensor G (optionally with a trailing rank axis) return, for every
    mode m, M_m[i,r] = sum_{i_j, j != m} G[..i..(,r)] prod_{j != m} Ft[j][i_j, r]."""
    k = len(Ft)
    has_r = (G.ndim == k + 1)
    out = []
    for m in range(k):
        if k == 1:      # single feature mode: nothing to contract
            out.append(G if has_r else G.unsqueeze(-1).expand(G.shape[0], Ft[0].shape[1]))
            continue
        letters = 'abcdefgh'[:k]
        expr_in = letters + ('r' if has_r else '')
        ops = [G]
        terms = [expr_in]
        for j in range(k):
            if j != m:
                ops.append(Ft[j])
                terms.append(letters[j] + 'r')
        out.append(torch.einsum(','.join(terms) + '->' + letters[m] + 'r', *ops))
    return out


def pack_sizes(dims, R, C):
    """Offsets of each factor inside the flat parameter vector theta =
    [F_0 | ... | F_{k-1} | (F_C) | bias(std only)] (row-major (I_m, R) blocks)."""
    sizes = [int(d) * R for d in dims] + ([C * R] if C > 0 else [])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return sizes, offs


def pack(Bcp, bias=None):
    parts = [F.reshape(-1) for F in Bcp]
    if bias is not None:
        parts.append(bias.reshape(-1))
    return torch.cat(parts)


def closed_form_std(X, y, Bcp, bias, weights, non_negative, softplus_kwargs=None):
    """Unnormalised local sums for the standard model.

    gradsum = [ dFt_0 | ... | dFt_{k-1} | sum_n res_n | sum_n res_n^2 ] with
    dFt_m[i,r] = w_r * sum_n res_n * sum_{others} X[n,..] prod_{j!=m} Ft_j, res_n = yhat_n - y_n.
    (Multiply the factor blocks and dbias by 2/N and the last entry by 1/N to get the
    reference's dL/dFt, dL/dbias and MSE — done in ``finish``.)
    """
    Ft = tilde(Bcp, non_negative, softplus_kwargs)
    u = inner_u(X, Ft)
    y_hat = u @ weights.to(u.dtype) + bias
    res = y_hat - y
    G = torch.einsum('n,n...->...', res, X)
    M = mttkrp_all(G, Ft)
    dFt = [Mm * weights.to(u.dtype)[None, :] for Mm in M]
    gradsum = torch.cat([d.reshape(-1) for d in dFt] + [res.sum().reshape(1), (res * res).sum().reshape(1)])
    return {'u': u, 'y_hat': y_hat, 'res': res, 'gradsum': gradsum}


def closed_form_mn(X, y, Bcp, weights, non_negative, class_weights, softplus_kwargs=None):
    """Unnormalised local sums for the multinomial model (double softmax, mn:180,187 + 364-366).

    gradsum = [ dFt_0 | ... | dFt_{k-1} | dFt_C | sum_n -omega[y_n] log Q[n,y_n] ], where
    dP = omega[y_n](Q - onehot) (NOT yet divided by W = sum_n omega[y_n]).
    """
    Ft_all = tilde(Bcp, non_negative, softplus_kwargs)
    Ft, FC = Ft_all[:-1], Ft_all[-1]
    w = weights.to(X.dtype)
    u = inner_u(X, Ft)                                   # (N,R)
    Z = (u * w[None, :]) @ FC.T                          # (N,C)
    P = torch.softmax(Z, dim=1)
    Q = torch.softmax(P, dim=1)
    om = torch.as_tensor(class_weights, dtype=X.dtype)[y]  # (N,)
    N = X.shape[0]
    logQy = torch.log(Q[torch.arange(N), y])
    loss_sum = -(om * logQy).sum()
    onehot = torch.zeros_like(Q)
    onehot[torch.arange(N), y] = 1
    dP = om[:, None] * (Q - onehot)
    dZ = P * (dP - (dP * P).sum(dim=1, keepdim=True))
    v = (dZ @ FC) * w[None, :]                           # (N,R)
    dFC = (dZ.T @ u) * w[None, :]                        # (C,R)
    G = torch.einsum('nr,n...->...r', v, X)              # (I_1..I_k, R)
    M = mttkrp_all(G, Ft)
    gradsum = torch.cat([d.reshape(-1) for d in M] + [dFC.reshape(-1), loss_sum.reshape(1)])
    return {'u': u, 'Z': Z, 'P': P, 'Q': Q, 'dZ': dZ, 'v': v, 'gradsum': gradsum, 'W': om.sum()}


def finish(gradsum, Bcp, non_negative, lambda_L2, grad_scale, loss_scale, has_bias, softplus_kwargs=None):
    """gradsum -> (flat gradient wrt the RAW parameters, data loss, data loss + penalty).

    grad[F_m] = grad_scale * dFt_m * softplus'(F_m) + lambda * F_m / ||F_m||_F   (Appendix A)
    grad[bias] = grad_scale * gradsum[Pf];  loss_data = loss_scale * gradsum[-1].
    """
    sp = softplus_kwargs or DEFAULT_SOFTPLUS
    outs, off, pen = [], 0, 0.0
    for m, F in enumerate(Bcp):
        n = F.numel()
        d = gradsum[off:off + n].reshape(F.shape).to(F.dtype) * grad_scale
        if non_negative[m]:
            d = d * softplus_grad(F, sp['beta'], sp['threshold'])
        nrm = torch.sqrt(torch.sum(F * F))
        pen = pen + nrm
        outs.append((d + lambda_L2 * F / nrm).reshape(-1))
        off += n
    if has_bias:
        outs.append((gradsum[off] * grad_scale).reshape(1).to(Bcp[0].dtype))
        off += 1
    loss_data = gradsum[-1] * loss_scale
    return torch.cat(outs), loss_data, loss_data + lambda_L2 * pen


def adam_step(theta, grad, m, v, vmax, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              amsgrad=False):
    """torch/optim/adam.py _single_tensor_adam (non-capturable, no maximize), in place; step is 1-based."""
    b1, b2 = betas
    if weight_decay != 0:
        grad = grad + weight_decay * theta
    m.lerp_(grad, 1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    if amsgrad:
        torch.maximum(vmax, v, out=vmax)
        denom = (vmax.sqrt() / bc2_sqrt).add_(eps)
    else:
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
    theta.addcdiv_(m, denom, value=-step_size)
    return theta


# ----------------------------------------------------------------------------
# Synthetic inputs — shared by tests and bench
# ----------------------------------------------------------------------------

def synth_std(N, dims, R, seed, dtype=torch.float32, noise=0.01, bias=0.1):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn((N, *dims), generator=g, dtype=dtype)
    Fstar = [0.3 * torch.randn((d, R), generator=g, dtype=dtype) for d in dims]
    w = torch.ones(R, dtype=dtype)
    y = lin_model(X, Fstar, w, [False] * len(dims), torch.tensor([bias], dtype=dtype))
    y = y + noise * torch.randn(y.shape, generator=g, dtype=dtype)
    return X, y, Fstar


def synth_mn(N, dims, R, C, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn((N, *dims), generator=g, dtype=torch.float32)
    Fstar = [0.3 * torch.randn((d, R), generator=g) for d in list(dims) + [C]]
    w = torch.ones(R)
    P = mn_model(X, Fstar, w, [False] * (len(dims) + 1))
    y = torch.argmax(P, dim=1)
    for c in range(C):          # guarantee every class appears (mn:279 uses len(unique(y)))
        if not (y == c).any():
            y[c % N] = c
    return X, y, Fstar


def init_std(dims, R, non_negative, scale=1.0, dtype=torch.float32, seed=321):
    """std:42-43 orthogonal init drawn on the CPU under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    B = [torch.nn.init.orthogonal_(torch.empty(d, R, dtype=dtype), gain=scale) for d in dims]
    B = [(B[ii] + torch.std(B[ii]) * 2 * non_negative[ii]) / (non_negative[ii] + 1) if B[0].shape[0] > 1 else B[ii]
         for ii in range(len(B))]
    return B


def init_mn(dims_with_C, R, non_negative, scale=1.0, seed=321):
    """mn:111 uniform init."""
    torch.manual_seed(seed)
    return [torch.rand((d, R)) * scale - (1 - non_negative[ii]) * (scale / 2) for ii, d in enumerate(dims_with_C)]
