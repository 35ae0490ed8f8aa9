"""Pin the spectral oracle port (oracle/tr_oracle_spectral.py): it reproduces what the UNMODIFIED
/root/reference/spectral_tensor_regression.py returned for the seeded inputs stored in tests/golden/spec_*.npz
(made by `python -m oracle.make_golden spec`), and, where /root/reference is present, agrees with it live."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import tr_oracle_spectral as OS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SPEC = sorted(glob.glob(os.path.join(GOLDEN, 'spec_*.npz')))
ADAM = {'lr': 0.01, 'amsgrad': True}
LBFGS = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
         'history_size': 100, 'line_search_fn': 'strong_wolfe'}


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300) if b.size else 0.0


def case(path):
    z = np.load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    Bn = [torch.from_numpy(z[f'Bn_init_{i}']) for i in range(3)]
    Bc = [torch.from_numpy(z[f'Bc_init_{i}']) for i in range(3)]
    w = torch.from_numpy(z['weights'])
    nn = [bool(v) for v in z['non_negative']]
    return z, X, y, Bn, Bc, w, nn, float(z['lambda_L2'])


def test_spectral_fixture_inventory():
    assert len(SPEC) >= 4


@pytest.mark.parametrize('path', SPEC, ids=[os.path.basename(p)[:-4] for p in SPEC])
def test_spectral_port_matches_reference_golden(path):
    torch.set_num_threads(1)
    z, X, y, Bn, Bc, w, nn, lam = case(path)
    tol = 1e-12 if X.dtype == torch.float64 else 2e-6
    bias = torch.zeros(y.shape[1], dtype=X.dtype)
    r = OS.loss_grad(X, y, Bn, Bc, bias, w, nn, lam)
    assert rel(r['y_hat'], z['y_hat']) < tol
    assert abs(r['loss'] - float(z['loss'])) < tol * abs(float(z['loss']))
    for i in range(3):
        assert rel(r['grad_n'][i], z[f'grad_n_{i}']) < tol
        assert rel(r['grad_c'][i], z[f'grad_c_{i}']) < tol
    assert rel(r['dbias'], z['dbias']) < tol
    f = OS.fit_adam(X, y, Bn, Bc, bias, w, nn, lam, 20, ADAM)
    assert rel(f['loss_running'], z['adam_loss_running']) < 10 * tol
    for i in range(3):
        assert rel(f['Bcp_n'][i], z[f'adam_Bn_{i}']) < 10 * tol
        assert rel(f['Bcp_c'][i], z[f'adam_Bc_{i}']) < 10 * tol
    assert rel(f['bias'], z['adam_bias']) < 10 * tol
    if 'lbfgs_loss_running' in z.files:
        lb = OS.fit_lbfgs(X, y, Bn, Bc, bias, w, nn, lam, 6, 1e-50, 10, 1, LBFGS)
        assert rel(lb['loss_running'], z['lbfgs_loss_running']) < 1e-9
        assert rel(lb['bias'], z['lbfgs_bias']) < 1e-7


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
def test_spectral_port_vs_live_reference():
    torch.set_num_threads(1)
    SPR = ref_loader.spectral()
    X, y = OS.synth(30, 6, 7, 2, 2, 2, 2, 77, dtype=torch.float64)
    Bn, Bc = OS.init(6, 7, 2, 2, 2, 2, dtype=torch.float64, seed=5)
    nn = [False, True, False]
    w = torch.tensor([0.5, 2.0, 1.0, 1.0], dtype=torch.float64)
    bias = 0.1 * torch.ones(2, dtype=torch.float64)
    want = SPR.lin_model(X, Bn, w[:2], nn, bias) + SPR.stepwise_spectral_model(X, Bc, w[2:], nn, bias)
    got = OS.model(X, Bn, Bc, w, nn, bias)
    assert rel(got, want) < 1e-13
    assert rel(OS.spectral_model(X, Bc, w[2:], nn, bias), SPR.spectral_model(X, Bc, w[2:], nn, bias)) < 1e-13
    assert rel(OS.lin_model(X, Bn, w[:2], nn, bias), SPR.lin_model(X, Bn, w[:2], nn, bias)) < 1e-13
    assert rel(OS.L2_penalty(Bn + Bc), SPR.L2_penalty(Bn) + SPR.L2_penalty(Bc)) < 1e-13
