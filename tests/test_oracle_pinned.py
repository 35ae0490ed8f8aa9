"""Pin the oracle: (1) the port reproduces the unmodified reference's outputs stored in
tests/golden (made by oracle/make_golden.py), (2) the closed form agrees
with the port's autograd in fp64, (3) where /root/reference is present, the port is compared
with the reference modules live, (4) the notebook known-answer loss."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import tr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STD = sorted(glob.glob(os.path.join(GOLDEN, 'std_*.npz')))
MN = sorted(glob.glob(os.path.join(GOLDEN, 'mn_*.npz')))
ADAM = {'lr': 0.01, 'amsgrad': True}


def load(path):
    z = np.load(path)
    k = len([n for n in z.files if n.startswith('Bcp_init_')])
    return z, k


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_fixture_inventory():
    assert len(STD) >= 6 and len(MN) >= 4


@pytest.mark.parametrize('path', STD, ids=[os.path.basename(p)[:-4] for p in STD])
def test_port_matches_reference_std(path):
    torch.set_num_threads(1)
    z, k = load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    tol = 1e-12 if X.dtype == torch.float64 else 2e-6
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    w = torch.from_numpy(z['weights'])
    nn = [bool(v) for v in z['non_negative']]
    bias = torch.tensor([float(z['bias_init'])], dtype=X.dtype)
    lam = float(z['lambda_L2'])
    r = O.std_loss_grad(X, y, B0, bias, w, nn, lam)
    assert rel(r['y_hat'], z['y_hat']) < tol
    assert abs(r['loss'].item() - float(z['loss'])) < tol * abs(float(z['loss']))
    for i in range(k):
        assert rel(r['grads'][i], z[f'grad_{i}']) < tol
    assert rel(r['dbias'], z['dbias']) < tol
    f = O.fit_adam_std(X, y, B0, bias, w, nn, lam, 20, ADAM)
    assert rel(f['loss_running'], z['adam_loss_running']) < 10 * tol
    for i in range(k):
        assert rel(f['Bcp'][i], z[f'adam_Bcp_{i}']) < 10 * tol
    assert rel(f['bias'], z['adam_bias']) < 10 * tol
    if 'lbfgs_loss_running' in z.files:
        lb = O.fit_lbfgs_std(X, y, B0, bias, w, nn, lam, 6, 1e-50, 10,
                             {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07,
                              'tolerance_change': 1e-09, 'history_size': 100, 'line_search_fn': 'strong_wolfe'})
        assert rel(lb['loss_running'], z['lbfgs_loss_running']) < 1e-9
        for i in range(k):
            assert rel(lb['Bcp'][i], z[f'lbfgs_Bcp_{i}']) < 1e-8


@pytest.mark.parametrize('path', MN, ids=[os.path.basename(p)[:-4] for p in MN])
def test_port_matches_reference_mn(path):
    torch.set_num_threads(1)
    z, k = load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    w = torch.from_numpy(z['weights'])
    nn = [bool(v) for v in z['non_negative']]
    lam = float(z['lambda_L2'])
    r = O.mn_loss_grad(X, y, B0, w, nn, z['class_weights'], lam)
    assert rel(r['P'], z['P']) < 2e-6
    assert abs(r['loss'].item() - float(z['loss'])) < 2e-6 * abs(float(z['loss']))
    for i in range(k):
        assert rel(r['grads'][i], z[f'grad_{i}']) < 2e-6
    f = O.fit_adam_mn(X, y, B0, w, nn, z['class_weights'], lam, 20, ADAM)
    assert rel(f['loss_running'], z['adam_loss_running']) < 2e-5
    for i in range(k):
        assert rel(f['Bcp'][i], z[f'adam_Bcp_{i}']) < 2e-5


def test_port_matches_reference_hierarchical():
    """multinomial_tensor_regression_hierarchical.py (unweighted CE, three Adam parameter groups with one
    learning rate) == the multinomial port with class weights of one."""
    z, k = load(os.path.join(GOLDEN, 'hier_2mode.npz'))
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    nn = [bool(v) for v in z['non_negative']]
    R, C = int(z['R']), int(z['C'])
    f = O.fit_adam_mn(X, y, B0, torch.ones(R), nn, np.ones(C, dtype=np.float32), float(z['lambda_L2']), 20, ADAM)
    assert rel(f['loss_running'], z['adam_loss_running']) < 2e-5
    for i in range(k):
        assert rel(f['Bcp'][i], z[f'adam_Bcp_{i}']) < 2e-5


@pytest.mark.parametrize('path', STD, ids=[os.path.basename(p)[:-4] for p in STD])
def test_closed_form_vs_autograd_std(path):
    z, k = load(path)
    d = torch.float64
    X, y = torch.from_numpy(z['X']).to(d), torch.from_numpy(z['y']).to(d)
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']).to(d) for i in range(k)]
    w = torch.from_numpy(z['weights']).to(d)
    nn = [bool(v) for v in z['non_negative']]
    bias = torch.tensor([float(z['bias_init'])], dtype=d)
    lam = float(z['lambda_L2'])
    N = X.shape[0]
    ref = O.std_loss_grad(X, y, B0, bias, w, nn, lam)
    cf = O.closed_form_std(X, y, B0, bias, w, nn)
    grad, loss_data, loss = O.finish(cf['gradsum'], B0, nn, lam, 2.0 / N, 1.0 / N, True)
    want = torch.cat([g.reshape(-1) for g in ref['grads']] + [ref['dbias'].reshape(-1)])
    assert rel(cf['y_hat'], ref['y_hat']) < 1e-13
    assert rel(grad, want) < 1e-12
    assert abs(loss_data.item() - ref['loss_data'].item()) < 1e-12 * abs(ref['loss_data'].item())
    assert abs(loss.item() - ref['loss'].item()) < 1e-12 * abs(ref['loss'].item())


@pytest.mark.parametrize('path', MN, ids=[os.path.basename(p)[:-4] for p in MN])
def test_closed_form_vs_autograd_mn(path):
    z, k = load(path)
    d = torch.float64
    X, y = torch.from_numpy(z['X']).to(d), torch.from_numpy(z['y'])
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']).to(d) for i in range(k)]
    w = torch.from_numpy(z['weights']).to(d)
    nn = [bool(v) for v in z['non_negative']]
    lam = float(z['lambda_L2'])
    cw = z['class_weights'].astype(np.float64)
    ref = O.mn_loss_grad(X, y, B0, w, nn, cw, lam)
    cf = O.closed_form_mn(X, y, B0, w, nn, cw)
    W = cf['W'].item()
    grad, loss_data, loss = O.finish(cf['gradsum'], B0, nn, lam, 1.0 / W, 1.0 / W, False)
    want = torch.cat([g.reshape(-1) for g in ref['grads']])
    assert rel(cf['P'], ref['P']) < 1e-13
    assert rel(grad, want) < 1e-12
    assert abs(loss.item() - ref['loss'].item()) < 1e-12 * abs(ref['loss'].item())


def test_adam_step_restatement_matches_torch():
    torch.manual_seed(0)
    for amsgrad in (False, True):
        for wd in (0.0, 0.1):
            p = torch.randn(37, dtype=torch.float64)
            q = p.clone().requires_grad_(True)
            opt = torch.optim.Adam([q], lr=0.01, betas=(0.8, 0.99), eps=1e-7, weight_decay=wd, amsgrad=amsgrad)
            m, v, vm = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
            for step in range(1, 6):
                g = torch.randn(37, dtype=torch.float64)
                q.grad = g.clone()
                opt.step()
                O.adam_step(p, g, m, v, vm, step, lr=0.01, betas=(0.8, 0.99), eps=1e-7, weight_decay=wd,
                            amsgrad=amsgrad)
                assert torch.allclose(p, q.detach(), rtol=0, atol=1e-15)


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
def test_port_vs_live_reference():
    """Fresh shapes (not in the fixtures) through the unmodified reference modules."""
    torch.set_num_threads(1)
    STR, MTR = ref_loader.standard(), ref_loader.multinomial()
    X, y, _ = O.synth_std(20, (3, 4, 5, 2), 3, 99, dtype=torch.float64)
    nn = [True, False, True, False, False]
    B0 = O.init_std((3, 4, 5, 2), 3, nn, dtype=torch.float64)
    w = torch.tensor([0.5, 1.0, 2.0], dtype=torch.float64)
    bias = torch.tensor([0.3], dtype=torch.float64)
    want = STR.lin_model(X, B0, w, nn, bias)
    got = O.lin_model(X, B0, w, nn, bias)
    assert rel(got, want) < 1e-14
    cf = O.closed_form_std(X, y, B0, bias, w, nn)
    assert rel(cf['y_hat'], want) < 1e-13
    Xm, ym, _ = O.synth_mn(30, (4, 3), 2, 3, 98)
    nnm = [False, True, False]
    Bm = O.init_mn([4, 3, 3], 2, nnm)
    wm = torch.ones(2)
    assert rel(O.mn_model(Xm, Bm, wm, nnm), MTR.model(Xm, Bm, wm, nnm)) < 1e-6
    assert abs(O.L2_penalty(Bm).item() - MTR.L2_penalty(Bm).item()) < 1e-6


@pytest.mark.slow
def test_notebook_known_answer():
    """demo_TensorRegression.ipynb cells 5+8: fp64, seed 321, rank 10, L-BFGS strong-Wolfe.
    Saved log: 560125.5196947237, 1699.8925874402807, 0.041904340578888165 (x11), 'Convergence reached'."""
    import scipy.signal
    torch.manual_seed(321)
    np.random.seed(321)
    dims = [2000, 500, 500]
    Xcp = [torch.rand(dims[0], 4) - 0.5,
           torch.vstack([torch.sin(torch.linspace(0, 140, dims[1])),
                         torch.cos(torch.linspace(2, 19, dims[1])),
                         torch.linspace(0, 1, dims[1]),
                         torch.cos(torch.linspace(0, 17, dims[1])) > 0]).T,
           torch.tensor(scipy.signal.savgol_filter(np.random.rand(dims[2], 4), 15, 3, axis=0)) - 0.5]
    Bcp_true = Xcp[1:]
    from oracle.tensorly_standin import cp_to_tensor, inner
    X_fake = cp_to_tensor((np.ones(4), Xcp))
    y = inner(X_fake + torch.rand(dims) / 100, cp_to_tensor((np.ones(4), Bcp_true)), n_modes=2)
    X = X_fake - X_fake.mean(0)
    del X_fake
    assert X.dtype == torch.float64
    nn = [False, False]
    B0 = [torch.nn.init.orthogonal_(torch.empty(d, 10, dtype=torch.float64), gain=0.005) for d in dims[1:]]
    B0 = [(B0[i] + torch.std(B0[i]) * 2 * nn[i]) / (nn[i] + 1) for i in range(2)]
    out = O.fit_lbfgs_std(X, y, B0, torch.tensor([0.0], dtype=torch.float64), torch.ones(10, dtype=torch.float64),
                          nn, 1e-5, 200, 1e-50, 10,
                          {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07,
                           'tolerance_change': 1e-09, 'history_size': 100, 'line_search_fn': 'strong_wolfe'})
    L = out['loss_running']
    assert out['converged'] and len(L) == 13
    assert abs(L[0] - 560125.5196947237) / 560125.5196947237 < 1e-3
    assert abs(L[-1] - 0.041904340578888165) / 0.041904340578888165 < 1e-5
    assert all(v == L[2] for v in L[2:])
