"""CPU restatement of the fit path of the reference's spectral_tensor_regression.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Op-for-op port in torch on the CPU (einsum chain, torch autograd for
the gradients, ``torch.optim`` for the update), following
  spectral_tensor_regression.py:62-94 (non_neg_fn), 118-165 (lin_model), 168-221 (spectral_model),
    284-337 (stepwise_latents_model), 339-390 (stepwise_spectral_model), 393-416 (L2_penalty),
    573-586 / 727-733 (closure: MSE over (T, n_out) + lambda (L2(Bcp_n) + L2(Bcp_c)), backward, step),
    601-632 / 726-752 (fit loops).
Parity pinning: this port == the unmodified reference run through the tensorly stand-in
(tests/golden/spec_*.npz made by oracle/make_golden.py; tests/test_oracle_pinned.py).
"""
import numpy as np
import torch

from .tensorly_standin import cp_to_tensor, inner
from .tr_oracle import DEFAULT_SOFTPLUS, non_neg


def lin_model(X, Bcp_n, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:118-165."""
    if Bcp_n[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    fac = non_neg([b[:, :, 0] for b in Bcp_n], non_negative, softplus_kwargs)
    return inner(X, cp_to_tensor((weights, fac)), n_modes=X.ndim - 1).squeeze() + bias


def spectral_model(X, Bcp_c, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:168-221 (what predict adds to lin_model): the norm over the complex axis of the complete CP
    contraction of every complex slice, plus bias."""
    if Bcp_c[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    y_hat_all = []
    for ii in range(Bcp_c[0].shape[2]):
        fac = non_neg([Bcp_c[0][:, :, ii]] + [Bcp_c[jj][:, :, 0] for jj in range(1, len(Bcp_c))], non_negative,
                      softplus_kwargs)
        y_hat_all += [inner(X, cp_to_tensor((weights, fac)), n_modes=X.ndim - 1).squeeze()[None, ...]]
    return torch.norm(torch.vstack(y_hat_all), dim=0) + bias


def stepwise_spectral_model(X, Bcp_c, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:339-390."""
    if Bcp_c[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    B = non_neg(Bcp_c, non_negative, softplus_kwargs)
    X_1a = torch.norm(torch.einsum('twd,wrc -> tdrc', X, B[0]), dim=3)
    X_1b = torch.einsum('tdr,drs -> tr', X_1a, B[1])
    return torch.einsum('tr,nrs -> tn', X_1b, B[2]) + bias


def L2_penalty(Bcp):
    """spectral:393-416."""
    ii = 0
    for comp in Bcp:
        ii = ii + torch.sqrt(torch.sum(comp ** 2))
    return ii


def model(X, Bcp_n, Bcp_c, weights, non_negative, bias, softplus_kwargs=None):
    """The expression of the fit closures (spectral:577-578 / 727-728)."""
    rn = Bcp_n[0].shape[1]
    return lin_model(X, Bcp_n, weights[:rn], non_negative, bias, softplus_kwargs) + \
        stepwise_spectral_model(X, Bcp_c, weights[rn:], non_negative, bias, softplus_kwargs)


def _leaf(ts):
    return [t.detach().clone().requires_grad_(True) for t in ts]


def loss_grad(X, y, Bcp_n, Bcp_c, bias, weights, non_negative, lambda_L2, softplus_kwargs=None):
    """One closure evaluation: prediction, data loss, total loss, gradients wrt the raw parameters."""
    Bn, Bc = _leaf(Bcp_n), _leaf(Bcp_c)
    b = bias.detach().clone().requires_grad_(True)
    y_hat = model(X, Bn, Bc, weights, non_negative, b, softplus_kwargs)
    mse = torch.nn.MSELoss()(y_hat, y)
    loss = mse + lambda_L2 * (L2_penalty(Bn) + L2_penalty(Bc))
    loss.backward()
    return {'y_hat': y_hat.detach(), 'loss_data': mse.item(), 'loss': loss.item(),
            'grad_n': [t.grad if t.grad is not None else torch.zeros_like(t) for t in Bn],
            'grad_c': [t.grad if t.grad is not None else torch.zeros_like(t) for t in Bc],
            'dbias': b.grad}


def fit_adam(X, y, Bcp_n, Bcp_c, bias, weights, non_negative, lambda_L2, n_iter, adam_kwargs, softplus_kwargs=None):
    """spectral:722-733 for a fixed number of iterations (no convergence test)."""
    Bn, Bc = _leaf(Bcp_n), _leaf(Bcp_c)
    b = bias.detach().clone().requires_grad_(True)
    opt = torch.optim.Adam(Bn + Bc + [b], **adam_kwargs)
    losses = []
    for _ in range(n_iter):
        opt.zero_grad()
        y_hat = model(X, Bn, Bc, weights, non_negative, b, softplus_kwargs)
        loss = torch.nn.MSELoss()(y_hat, y) + lambda_L2 * (L2_penalty(Bn) + L2_penalty(Bc))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return {'loss_running': losses, 'Bcp_n': [t.detach() for t in Bn], 'Bcp_c': [t.detach() for t in Bc],
            'bias': b.detach()}


def fit_lbfgs(X, y, Bcp_n, Bcp_c, bias, weights, non_negative, lambda_L2, max_iter, tol, patience, log_interval,
              lbfgs_kwargs, softplus_kwargs=None):
    """spectral:601-640 (loss logged WITH the penalty, convergence on the last patience-1 logged losses)."""
    Bn, Bc = _leaf(Bcp_n), _leaf(Bcp_c)
    b = bias.detach().clone().requires_grad_(True)
    opt = torch.optim.LBFGS(Bn + Bc + [b], **lbfgs_kwargs)

    def closure():
        opt.zero_grad()
        y_hat = model(X, Bn, Bc, weights, non_negative, b, softplus_kwargs)
        loss = torch.nn.MSELoss()(y_hat, y) + lambda_L2 * (L2_penalty(Bn) + L2_penalty(Bc))
        loss.backward()
        return loss

    losses, converged = [], False
    for ii in range(max_iter):
        if ii % log_interval == 0:
            y_hat = model(X, Bn, Bc, weights, non_negative, b, softplus_kwargs).detach()
            losses.append((torch.nn.MSELoss()(y_hat, y) + lambda_L2 * (L2_penalty(Bn) + L2_penalty(Bc))).item())
        if len(losses) > patience:
            if np.sum(np.abs(np.diff(losses[-patience + 1:]))) < tol:
                converged = True
                break
        elif np.isnan(losses[-1]):
            break
        opt.step(closure)
    return {'loss_running': losses, 'converged': converged, 'Bcp_n': [t.detach() for t in Bn],
            'Bcp_c': [t.detach() for t in Bc], 'bias': b.detach()}


def synth(T, W, D, n_out, rn, rs, cc, seed, dtype=torch.float32, noise=0.01):
    """Seeded X (T, W, D), y (T, n_out) from a planted model of the same family."""
    g = torch.Generator().manual_seed(seed)
    X = torch.randn((T, W, D), generator=g, dtype=torch.float64)
    Bn = [0.4 * torch.randn((d, rn, 1), generator=g, dtype=torch.float64) for d in (W, D, n_out)]
    Bc = [0.4 * torch.randn((d, rs, c), generator=g, dtype=torch.float64) for d, c in ((W, cc), (D, 1), (n_out, 1))]
    nn = [False, False, False]
    y = model(X, Bn, Bc, torch.ones(rn + rs, dtype=torch.float64), nn, torch.zeros(n_out, dtype=torch.float64))
    if y.ndim == 1:
        y = y.reshape(T, -1)
    y = y + noise * torch.randn(y.shape, generator=g, dtype=torch.float64)
    return X.to(dtype), y.to(dtype)


def init(W, D, n_out, rn, rs, cc, scale=0.5, dtype=torch.float32, seed=321):
    """Seeded dense initial factors in the reference's list layout ([Bcp_n, Bcp_c])."""
    g = torch.Generator().manual_seed(seed)
    Bn = [(scale * (torch.rand((d, rn, 1), generator=g, dtype=torch.float64) - 0.3)).to(dtype) for d in (W, D, n_out)]
    Bc = [(scale * (torch.rand((d, rs, c), generator=g, dtype=torch.float64) - 0.3)).to(dtype)
          for d, c in ((W, cc), (D, 1), (n_out, 1))]
    return Bn, Bc


def pack(Bcp_n, Bcp_c, bias):
    """theta layout of the C ABI: [Bcp_n[0..2] | Bcp_c[0..2] | bias], row-major."""
    return torch.cat([t.reshape(-1) for t in list(Bcp_n) + list(Bcp_c)] + [bias.reshape(-1)])
