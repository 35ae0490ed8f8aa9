"""GPU parity of the spectral variant (tensor_regression_b200/spectral_tensor_regression.py, tr_spec_* in the C ABI)
against the outputs of the UNMODIFIED reference stored in tests/golden/spec_*.npz and against the oracle port
(oracle/tr_oracle_spectral.py) on seeded inputs.  Tolerances: the north star's 1e-5 (fp32) / 1e-10 (fp64), norm-relative;
fitted factors after 20 Adam iterations 1e-4 / 1e-9; L-BFGS (fp64) logged losses 1e-7."""
import ctypes
import glob
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import tr_oracle_spectral as OS

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SPEC = sorted(glob.glob(os.path.join(GOLDEN, 'spec_*.npz')))
ADAM = {'lr': 0.01, 'amsgrad': True}
LBFGS = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
         'history_size': 100, 'line_search_fn': 'strong_wolfe'}
DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300) if b.size else 0.0


def case(path):
    z = np.load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    Bn = [torch.from_numpy(z[f'Bn_init_{i}']) for i in range(3)]
    Bc = [torch.from_numpy(z[f'Bc_init_{i}']) for i in range(3)]
    w = z['weights']
    nn = [bool(v) for v in z['non_negative']]
    return z, X, y, Bn, Bc, w, nn, float(z['lambda_L2'])


def model_of(z, X, y, Bn, Bc, w, nn):
    from tensor_regression_b200 import spectral_tensor_regression as SPR
    return SPR.CP_linear_regression(X.shape, y.shape, dtype=X.dtype, rank_normal=int(z['rank_normal']),
                                    rank_spectral=int(z['rank_spectral']), non_negative=nn, weights=w,
                                    Bcp_init=[[b.clone() for b in Bn], [b.clone() for b in Bc]],
                                    n_complex_dim=int(z['n_complex_dim']), device=DEV)


@pytest.mark.parametrize('path', SPEC, ids=[os.path.basename(p)[:-4] for p in SPEC])
def test_spectral_closure_vs_reference_golden(path):
    """One closure evaluation through the C ABI: prediction, losses and the gradient wrt every raw parameter."""
    z, X, y, Bn, Bc, w, nn, lam = case(path)
    tol = 1e-10 if X.dtype == torch.float64 else 1e-5
    m = model_of(z, X, y, Bn, Bc, w, nn)
    eng = m._engine()
    Xd, yd = X.to(DEV), y.to(DEV)
    beta, thr = m._sp()
    yhat = torch.empty_like(yd)
    gs = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr, yhat=yhat)
    launches_fused = eng.launch_info()['launches']
    q = int(z['rank_normal']) + int(z['rank_spectral']) * (int(z['n_complex_dim']) + 1)
    assert (eng.launch_info()['df1_slabs'] == 0) == (q <= 8)           # fused first pass whenever the channels fit
    n_total = y.numel()
    grad, loss = eng.finish(gs, 2.0 / n_total, 1.0 / n_total, m.theta, lam, m._mask(), beta, thr)
    assert rel(yhat, z['y_hat']) < tol
    assert abs(loss[0].item() - float(z['loss_data'])) < tol * abs(float(z['loss_data']))
    assert abs(loss[1].item() - float(z['loss'])) < tol * abs(float(z['loss']))
    want = np.concatenate([z[f'grad_n_{i}'].reshape(-1) for i in range(3)] + [z[f'grad_c_{i}'].reshape(-1) for i in range(3)]
                          + [z['dbias'].reshape(-1)])
    assert rel(grad, want) < tol
    # block by block as well (a small block must not hide behind a large one)
    off = 0
    for name in [f'grad_n_{i}' for i in range(3)] + [f'grad_c_{i}' for i in range(3)] + ['dbias']:
        n = z[name].size
        if n:
            assert rel(grad[off:off + n], z[name].reshape(-1)) < 10 * tol, name
        off += n
    # forward-only entry point (separate epilogue kernel, sums in double) gives the same prediction
    out = eng.forward(Xd, m.theta, m.weights, m._mask(), beta, thr, want=('yhat',))
    assert rel(out['yhat'], yhat) < tol
    # ... and so does the unfused fit path (window contraction, epilogue and second-mode gradient as separate kernels)
    eng.set_option('fused', 0)
    yhat0 = torch.empty_like(yd)
    gs0 = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr, yhat=yhat0)
    grad0, loss0 = eng.finish(gs0, 2.0 / n_total, 1.0 / n_total, m.theta, lam, m._mask(), beta, thr)
    assert eng.launch_info()['df1_slabs'] > 0 and eng.launch_info()['launches'] >= launches_fused
    assert rel(yhat0, z['y_hat']) < tol and rel(grad0, want) < tol
    assert abs(loss0[1].item() - float(z['loss'])) < tol * abs(float(z['loss']))


@pytest.mark.parametrize('path', SPEC, ids=[os.path.basename(p)[:-4] for p in SPEC])
def test_spectral_fit_adam_vs_reference_golden(path):
    z, X, y, Bn, Bc, w, nn, lam = case(path)
    f64 = X.dtype == torch.float64
    m = model_of(z, X, y, Bn, Bc, w, nn)
    conv = m.fit_Adam(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, verbose=False,
                      Adam_kwargs=ADAM)
    assert conv is False
    assert rel(m.loss_running, z['adam_loss_running']) < (1e-9 if f64 else 1e-5)
    ftol = 1e-9 if f64 else 1e-4
    for i in range(3):
        assert rel(m.Bcp_n[i], z[f'adam_Bn_{i}']) < ftol
        assert rel(m.Bcp_c[i], z[f'adam_Bc_{i}']) < ftol
    assert rel(m.bias, z['adam_bias']) < ftol
    # the reference's predict (lin_model + spectral_model) and predict_latents on the fitted model
    if 'adam_predict' in z.files:
        p = m.predict(X.to(DEV))
        assert isinstance(p, torch.Tensor) and p.device.type == 'cpu'
        assert rel(p, z['adam_predict']) < (1e-8 if f64 else 2e-4)
        assert rel(m.predict(X.numpy()), z['adam_predict']) < (1e-8 if f64 else 2e-4)       # host array, streamed
    if 'adam_latents' in z.files:
        assert rel(m.predict_latents(X.to(DEV)), z['adam_latents']) < (1e-8 if f64 else 2e-4)


@pytest.mark.parametrize('path', [p for p in SPEC if 'f64' in p], ids=lambda p: os.path.basename(p)[:-4])
def test_spectral_fit_lbfgs_vs_reference_golden(path):
    z, X, y, Bn, Bc, w, nn, lam = case(path)
    m = model_of(z, X, y, Bn, Bc, w, nn)
    conv = m.fit(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=6, tol=1e-50, patience=10, verbose=False,
                 running_loss_logging_interval=1, LBFGS_kwargs=LBFGS)
    assert bool(conv) == bool(z['lbfgs_converged'])
    # The norm makes this objective much less forgiving than the other models': perturbing X by 1e-15 (relative) moves
    # the REFERENCE's own logged losses by 2e-14, 3e-12, 2e-8 / 5e-7, 3e-6 / 5e-4, 1e-4 at logged iterations 1..5
    # (tools/README: measured with the oracle port, the same torch.optim.LBFGS).  The first four logged losses pin
    # the path; the last two only have to stay on the same descent.
    want = z['lbfgs_loss_running']
    assert len(m.loss_running) == len(want)
    assert rel(m.loss_running[:3], want[:3]) < 1e-9
    assert abs(m.loss_running[3] - want[3]) < 2e-5 * abs(want[3])
    assert rel(m.loss_running[4:], want[4:]) < 2e-2


@pytest.mark.parametrize('shape', [
    # T, W, D, n_out, rank_normal, rank_spectral, complex, dtype
    (37, 13, 21, 3, 2, 2, 2, torch.float32),       # D = 21: element loads, ragged tiles
    (64, 50, 128, 4, 2, 2, 2, torch.float32),      # 16-byte rows, one full warp tile
    (50, 9, 260, 2, 3, 3, 3, torch.float32),       # 12 channels: two passes over X per direction; D spans three tiles
    (33, 17, 36, 5, 4, 1, 1, torch.float64),
    (20, 5, 8, 2, 1, 1, 4, torch.float64),         # four complex columns
], ids=lambda s: 'x'.join(str(v) for v in s[:7]) + ('_f64' if s[7] == torch.float64 else '_f32'))
def test_spectral_kernels_vs_oracle(shape):
    """Seeded geometry sweep against the oracle port's autograd (fp64 truth for the fp32 cases as well)."""
    from tensor_regression_b200 import spectral_tensor_regression as SPR
    T, W, D, NO, rn, rs, cc, dtype = shape
    X, y = OS.synth(T, W, D, NO, rn, rs, cc, 4242, dtype=dtype)
    Bn, Bc = OS.init(W, D, NO, rn, rs, cc, dtype=dtype, seed=99)
    nn = [True, False, False]
    w = np.linspace(0.5, 1.5, rn + rs)
    lam = 0.01
    bias = 0.1 * torch.arange(1, NO + 1, dtype=dtype)
    r = OS.loss_grad(X.double(), y.double(), [b.double() for b in Bn], [b.double() for b in Bc], bias.double(),
                     torch.tensor(w, dtype=torch.float64), nn, lam)
    m = SPR.CP_linear_regression(X.shape, y.shape, dtype=dtype, rank_normal=rn, rank_spectral=rs, non_negative=nn,
                                 weights=w, Bcp_init=[Bn, Bc], n_complex_dim=cc - 1, device=DEV)
    m.bias.copy_(bias.to(DEV))
    eng = m._engine()
    beta, thr = m._sp()
    Xd, yd = X.to(DEV), y.to(DEV)
    yhat = torch.empty_like(yd)
    gs = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr, yhat=yhat)
    grad, loss = eng.finish(gs, 2.0 / y.numel(), 1.0 / y.numel(), m.theta, lam, m._mask(), beta, thr)
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert rel(yhat.reshape(-1), r['y_hat'].reshape(-1)) < tol
    assert abs(loss[1].item() - r['loss']) < tol * abs(r['loss'])
    want = torch.cat([g.reshape(-1) for g in r['grad_n'] + r['grad_c']] + [r['dbias'].reshape(-1)])
    assert rel(grad, want) < tol
    # relaunch: bit-identical (fixed summation orders, no atomics)
    gs2 = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr)
    assert torch.equal(gs, gs2)
    # shard sums add up: the gradsum of two halves of the sample axis sums to the whole (what the all-reduce relies on)
    h = T // 2
    ga = eng.fwd_grad(Xd[:h].contiguous(), yd[:h].contiguous(), m.theta, m.weights, m._mask(), beta, thr).clone()
    gb = eng.fwd_grad(Xd[h:].contiguous(), yd[h:].contiguous(), m.theta, m.weights, m._mask(), beta, thr)
    assert rel(ga + gb, gs) < 1e-12 if dtype == torch.float64 else rel(ga + gb, gs) < 1e-6
    # module-level functions
    got = SPR.lin_model(Xd, [b.to(DEV) for b in Bn], m.weights[:rn], nn, m.bias) + \
        SPR.stepwise_spectral_model(Xd, [b.to(DEV) for b in Bc], m.weights[rn:], nn, m.bias)
    assert rel(got.reshape(-1), r['y_hat'].reshape(-1)) < tol
    sp = OS.spectral_model(X.double(), [b.double() for b in Bc], torch.tensor(w[rn:], dtype=torch.float64), nn, bias.double())
    assert rel(SPR.spectral_model(Xd, [b.to(DEV) for b in Bc], m.weights[rn:], nn, m.bias).reshape(-1), sp.reshape(-1)) < tol


def test_spectral_api_surface_and_errors():
    from tensor_regression_b200 import spectral_tensor_regression as SPR
    from tensor_regression_b200._lib import TRError
    m = SPR.CP_linear_regression((10, 6, 8), (10, 2), rank_normal=2, rank_spectral=1, n_complex_dim=1, device=DEV)
    assert [tuple(b.shape) for b in m.Bcp_n] == [(6, 2, 1), (8, 2, 1), (2, 2, 1)]
    assert [tuple(b.shape) for b in m.Bcp_c] == [(6, 1, 2), (8, 1, 1), (2, 1, 1)]
    assert tuple(m.bias.shape) == (2,) and float(m.bias.abs().sum()) == 0.0
    assert m.theta.numel() == 2 * (6 + 8 + 2) + (12 + 8 + 2) + 2
    X = torch.randn(10, 6, 8, device=DEV)
    y = torch.randn(10, 2, device=DEV)
    with pytest.raises(TypeError):
        m.fit_Adam(X, y)                      # Adam_kwargs=None raises, as in the reference (spectral:705-712)
    with pytest.raises(TypeError):
        m.fit(X, y)
    with pytest.raises(ValueError):
        m.fit_Adam(torch.randn(10, 6, 9, device=DEV), y, Adam_kwargs=ADAM)
    with pytest.raises(ValueError):
        SPR.CP_linear_regression((10, 6, 8, 3), (10, 2), device=DEV)
    with pytest.raises(TRError):
        SPR.CP_linear_regression((10, 6, 8), (10, 2), device='cpu')
    m.fit_Adam(X, y, max_iter=3, Adam_kwargs=ADAM)
    assert len(m.loss_running) == 3
    p = m.get_params()
    assert set(p) == {'weights', 'Bcp_n', 'Bcp_c', 'non_negative', 'softplus_kwargs', 'rank', 'device', 'loss_running'}
    m2 = pickle.loads(pickle.dumps(m))
    assert torch.equal(m2.theta, m.theta)
    assert rel(m2.predict(X), m.predict(X)) == 0.0
    m3 = SPR.CP_linear_regression((10, 6, 8), (10, 2), rank_normal=2, rank_spectral=1, n_complex_dim=1, device=DEV)
    m3.set_params(p)
    assert all(torch.equal(a, b) for a, b in zip(m3.Bcp_n + m3.Bcp_c, m.Bcp_n + m.Bcp_c))
    fin_n, fin_c = m.return_Bcp_final()
    assert len(fin_n) == 3 and fin_c[0].shape == (6, 1, 2)
    # a standard-model entry point refuses a spectral handle
    from tensor_regression_b200 import _lib
    rc = _lib.lib.tr_forward_std(m._engine()._h, X.data_ptr(), 10, m.theta.data_ptr(), m.weights.data_ptr(), 0, 50.0, 1.0,
                                 y.data_ptr(), None)
    assert rc != 0 and b'spectral' in _lib.lib.tr_last_error(m._engine()._h)


def test_spectral_convergence_rule_and_nan_stop(capsys):
    """fit_Adam stops on the reference's rule (spectral:743-745) and both fits report a NaN loss like the reference."""
    from tensor_regression_b200 import spectral_tensor_regression as SPR
    X, y = OS.synth(40, 6, 8, 2, 1, 1, 2, 7)
    m = SPR.CP_linear_regression(X.shape, y.shape, rank_normal=1, rank_spectral=1, n_complex_dim=1, device=DEV)
    conv = m.fit_Adam(X.to(DEV), y.to(DEV), lambda_L2=0.0, max_iter=400, tol=1e3, patience=5, Adam_kwargs={'lr': 1e-3})
    assert conv is True and len(m.loss_running) == 7          # first test at ii = patience + 1
    m2 = SPR.CP_linear_regression(X.shape, y.shape, rank_normal=1, rank_spectral=1, n_complex_dim=1, device=DEV)
    Xn = X.clone()
    Xn[0, 0, 0] = float('nan')
    conv = m2.fit_Adam(Xn.to(DEV), y.to(DEV), max_iter=50, Adam_kwargs={'lr': 1e-3})
    assert conv is False and len(m2.loss_running) == 1
    assert 'Loss is NaN. Stopping.' in capsys.readouterr().out


@pytest.mark.parametrize('shape', [
    (1, 1, 1, 1, 1, 0, 1, torch.float64),          # one sample, one element, normal part only
    (3, 5, 7, 40, 2, 1, 2, torch.float32),         # more outputs than lanes in a warp; fewer samples than warps
    (9, 3, 300, 2, 0, 2, 3, torch.float32),        # spectral part only, D spans three warp tiles (separate-kernel path)
    (17, 65, 4, 3, 2, 2, 2, torch.float64),        # window longer than the batch of rows, tiny D
], ids=lambda s: 'x'.join(str(v) for v in s[:7]))
def test_spectral_edge_geometries(shape):
    from tensor_regression_b200 import spectral_tensor_regression as SPR
    T, W, D, NO, rn, rs, cc, dtype = shape
    g = torch.Generator().manual_seed(3)
    X = torch.randn((T, W, D), generator=g, dtype=dtype)
    y = torch.randn((T, NO), generator=g, dtype=dtype)
    Bn, Bc = OS.init(W, D, NO, rn, rs, cc, dtype=dtype, seed=17)
    nn = [False, True, False]
    w = np.linspace(0.8, 1.2, rn + rs)
    bias = torch.linspace(-0.2, 0.2, NO, dtype=dtype)
    both = rn > 0 and rs > 0
    if NO == 1 and both:
        pytest.skip('the reference expression broadcasts (T,) + (T,1) to (T,T) for one output')
    r = OS.loss_grad(X.double(), y.double(), [b.double() for b in Bn], [b.double() for b in Bc], bias.double(),
                     torch.tensor(w, dtype=torch.float64), nn, 0.01)
    m = SPR.CP_linear_regression(X.shape, y.shape, dtype=dtype, rank_normal=rn, rank_spectral=rs, non_negative=nn,
                                 weights=w, Bcp_init=[Bn, Bc], n_complex_dim=cc - 1, device=DEV)
    m.bias.copy_(bias.to(DEV))
    eng = m._engine()
    beta, thr = m._sp()
    yhat = torch.empty((T, NO), dtype=dtype, device=DEV)
    gs = eng.fwd_grad(X.to(DEV), y.to(DEV), m.theta, m.weights, m._mask(), beta, thr, yhat=yhat)
    grad, loss = eng.finish(gs, 2.0 / y.numel(), 1.0 / y.numel(), m.theta, 0.01, m._mask(), beta, thr)
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert rel(yhat.reshape(-1), r['y_hat'].reshape(-1)) < tol
    assert abs(loss[1].item() - r['loss']) < tol * abs(r['loss'])
    want = torch.cat([t.reshape(-1) for t in r['grad_n'] + r['grad_c']] + [r['dbias'].reshape(-1)])
    assert rel(grad, want) < tol


def test_spectral_empty_batch_and_limits():
    import ctypes
    from tensor_regression_b200 import _lib, engine
    eng = engine.SpectralEngine(6, 8, 2, 1, 1, 2, torch.float32, DEV)
    th = torch.zeros(eng.P, device=DEV)
    w = torch.ones(2, device=DEV)
    gs = torch.full((eng.n_gradsum,), 7.0, dtype=torch.float64, device=DEV)
    X = torch.empty((0, 6, 8), device=DEV)
    y = torch.empty((0, 2), device=DEV)
    rc = _lib.lib.tr_spec_fwd_grad(eng._h, None, None, 0, th.data_ptr(), w.data_ptr(), 0, 50.0, 1.0, gs.data_ptr(), None, None)
    assert rc == 0 and float(gs.abs().sum()) == 0.0          # N = 0: the sums are zero
    h = ctypes.c_void_p()
    assert _lib.lib.tr_spec_create(ctypes.byref(h), 0, 6, 8, 2, 9, 8, 1, 0) != 0        # 17 components
    assert b'rank' in _lib.lib.tr_last_error(None)
    assert _lib.lib.tr_spec_create(ctypes.byref(h), 0, 6, 8, 2, 0, 0, 1, 0) != 0        # no component at all
    assert _lib.lib.tr_spec_create(ctypes.byref(h), 0, 6, 8, 200, 1, 1, 1, 0) != 0      # too many outputs
    # tr_spec_* refuses a standard / multinomial handle
    e2 = engine.Engine((6, 8), 2, 0, torch.float32, DEV)
    rc = _lib.lib.tr_spec_fwd_grad(e2._h, X.data_ptr(), y.data_ptr(), 0, th.data_ptr(), w.data_ptr(), 0, 50.0, 1.0,
                                   gs.data_ptr(), None, None)
    assert rc != 0 and b'tr_spec_create' in _lib.lib.tr_last_error(e2._h)


# ---------------------------------------------------------------------------------------------
# single-pass kernel (tr_spectral_single.cuh): ring of whole samples in shared memory, X read once
# ---------------------------------------------------------------------------------------------
def _spectral_model(shape, seed=4242):
    from tensor_regression_b200 import spectral_tensor_regression as SPR
    T, W, D, NO, rn, rs, cc, dtype = shape
    X, y = OS.synth(T, W, D, NO, rn, rs, cc, seed, dtype=dtype)
    Bn, Bc = OS.init(W, D, NO, rn, rs, cc, dtype=dtype, seed=99)
    nn = [True, False, False]
    w = np.linspace(0.5, 1.5, rn + rs)
    bias = 0.1 * torch.arange(1, NO + 1, dtype=dtype)
    m = SPR.CP_linear_regression(X.shape, y.shape, dtype=dtype, rank_normal=rn, rank_spectral=rs, non_negative=nn,
                                 weights=w, Bcp_init=[Bn, Bc], n_complex_dim=cc - 1, device=DEV)
    m.bias.copy_(bias.to(DEV))
    return m, X, y, Bn, Bc, bias, w, nn


@pytest.mark.parametrize('shape', [
    # T, W, D, n_out, rank_normal, rank_spectral, complex, dtype
    (700, 64, 128, 4, 2, 2, 2, torch.float32),     # the bench workload's sample shape: 8 full row tiles, full warp tile
    (301, 50, 100, 3, 2, 2, 2, torch.float32),     # 25 active lanes, last row tile holds 2 rows, batch remainder of 2 rows
    (257, 13, 32, 5, 1, 1, 3, torch.float32),      # short window (one batch + remainder), 8 active lanes
    (3, 64, 128, 2, 2, 2, 2, torch.float32),       # fewer samples than forward warps
    (1, 9, 16, 2, 1, 1, 2, torch.float32),         # one sample
    (200, 33, 36, 2, 3, 1, 2, torch.float64),      # fp64: 2 elements per lane, 18 active lanes
    (150, 24, 64, 100, 2, 1, 2, torch.float64),    # more outputs than lanes
    (90, 40, 60, 3, 0, 2, 4, torch.float64),       # spectral part only, four complex columns (8 channels)
    (120, 16, 128, 3, 8, 0, 1, torch.float32),     # normal part only, 8 channels
], ids=lambda s: 'x'.join(str(v) for v in s[:7]) + ('_f64' if s[7] == torch.float64 else '_f32'))
@pytest.mark.parametrize('stages', [0, 2, 3])
def test_spectral_single_pass_vs_oracle_and_two_pass(shape, stages):
    T, W, D, NO, rn, rs, cc, dtype = shape
    m, X, y, Bn, Bc, bias, w, nn = _spectral_model(shape)
    lam = 0.01
    r = OS.loss_grad(X.double(), y.double(), [b.double() for b in Bn], [b.double() for b in Bc], bias.double(),
                     torch.tensor(w, dtype=torch.float64), nn, lam)
    eng = m._engine()
    beta, thr = m._sp()
    Xd, yd = X.to(DEV), y.to(DEV)
    eng.set_option('spec_single', 1)
    eng.set_option('spec_single_ns', stages)
    yhat = torch.full_like(yd, float('nan'))
    gs = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr, yhat=yhat).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-pass') and info['stages'] >= 2
    if stages:
        assert info['stages'] <= stages
    grad, loss = eng.finish(gs, 2.0 / y.numel(), 1.0 / y.numel(), m.theta, lam, m._mask(), beta, thr)
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert rel(yhat.reshape(-1), r['y_hat'].reshape(-1)) < tol
    assert abs(loss[1].item() - r['loss']) < tol * abs(r['loss'])
    want = torch.cat([g.reshape(-1) for g in r['grad_n'] + r['grad_c']] + [r['dbias'].reshape(-1)])
    assert rel(grad, want) < tol
    off = 0
    for blk in r['grad_n'] + r['grad_c'] + [r['dbias']]:
        if blk.numel():
            assert rel(grad[off:off + blk.numel()], blk.reshape(-1)) < 10 * tol
        off += blk.numel()
    # relaunch: bit-identical
    gs2 = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr)
    assert torch.equal(gs, gs2)
    # the two-pass path on the same inputs
    eng.set_option('spec_single', 0)
    gs0 = eng.fwd_grad(Xd, yd, m.theta, m.weights, m._mask(), beta, thr)
    assert not eng.launch_info()['path'].startswith('single-pass')
    assert rel(gs, gs0) < (1e-12 if dtype == torch.float64 else 2e-6)


@pytest.mark.parametrize('path', SPEC, ids=[os.path.basename(p)[:-4] for p in SPEC])
def test_spectral_single_pass_fit_adam_vs_reference_golden(path):
    """20 Adam iterations of the reference's fit_Adam on the single-pass kernel against the unmodified reference's outputs."""
    z, X, y, Bn, Bc, w, nn, lam = case(path)
    f64 = X.dtype == torch.float64
    q = int(z['rank_normal']) + int(z['rank_spectral']) * (int(z['n_complex_dim']) + 1)
    vec = 2 if f64 else 4
    if q > 8 or X.shape[2] > 32 * vec or X.shape[2] % vec or X.shape[1] > 64:
        pytest.skip('geometry outside the single-pass kernel (runs on the two-pass path)')
    m = model_of(z, X, y, Bn, Bc, w, nn)
    m._engine().set_option('spec_single', 1)
    m.fit_Adam(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, verbose=False, Adam_kwargs=ADAM)
    assert m._engine().launch_info()['path'].startswith('single-pass')
    assert rel(m.loss_running, z['adam_loss_running']) < (1e-9 if f64 else 1e-5)
    ftol = 1e-9 if f64 else 1e-4
    for i in range(3):
        assert rel(m.Bcp_n[i], z[f'adam_Bn_{i}']) < ftol
        assert rel(m.Bcp_c[i], z[f'adam_Bc_{i}']) < ftol
    assert rel(m.bias, z['adam_bias']) < ftol


def test_spectral_single_pass_refuses_geometry_loudly():
    from tensor_regression_b200._lib import TRError
    m, X, y, *_ = _spectral_model((20, 70, 64, 2, 1, 1, 2, torch.float32))        # 70 window rows > 64
    eng = m._engine()
    eng.set_option('spec_single', 1)
    beta, thr = m._sp()
    with pytest.raises(TRError, match='spec_single'):
        eng.fwd_grad(X.to(DEV), y.to(DEV), m.theta, m.weights, m._mask(), beta, thr)


def test_spectral_full_size_properties_single_pass():
    """The bench workload spec1 at full size (200 000 samples of 64 x 128 fp32, 6.55 GB): the single-pass kernel is the
    automatic choice, agrees with the two-pass path, relaunches bit-identically, its sums over two halves of the sample
    axis add up, and a slice of it agrees with the fp64 oracle."""
    from tensor_regression_b200 import engine
    T, W, D, NO, rn, rs, cc = 200000, 64, 128, 4, 2, 2, 2
    g = torch.Generator(device=DEV).manual_seed(5)
    X = torch.randn((T, W, D), generator=g, device=DEV)
    y = torch.randn((T, NO), generator=g, device=DEV)
    eng = engine.SpectralEngine(W, D, NO, rn, rs, cc, torch.float32, DEV)
    Bn, Bc = OS.init(W, D, NO, rn, rs, cc, seed=21)
    th = torch.cat([b.reshape(-1) for b in Bn + Bc] + [0.05 * torch.arange(1, NO + 1, dtype=torch.float32)]).to(DEV)
    assert th.numel() == eng.P
    w = torch.linspace(0.7, 1.3, rn + rs, device=DEV)
    gs = eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0).clone()
    assert eng.launch_info()['path'].startswith('single-pass')
    assert torch.equal(gs, eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0))
    eng.set_option('spec_single', 0)
    gs2 = eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0).clone()
    assert not eng.launch_info()['path'].startswith('single-pass')
    assert rel(gs, gs2) < 2e-6
    eng.set_option('spec_single', -1)
    h = 123457                                       # uneven halves
    ga = eng.fwd_grad(X[:h], y[:h], th, w, 0, 50.0, 1.0).clone()
    gb = eng.fwd_grad(X[h:], y[h:], th, w, 0, 50.0, 1.0)
    assert rel(ga + gb, gs) < 1e-6
    # a slice against the fp64 oracle (closure value and every gradient block through finish)
    n = 4000
    Xs, ys = X[70000:70000 + n].contiguous(), y[70000:70000 + n].contiguous()
    gsl = eng.fwd_grad(Xs, ys, th, w, 0, 50.0, 1.0)
    assert eng.launch_info()['path'].startswith('single-pass')
    grad, loss = eng.finish(gsl, 2.0 / ys.numel(), 1.0 / ys.numel(), th, 0.01, 0, 50.0, 1.0)
    bias = (0.05 * torch.arange(1, NO + 1, dtype=torch.float64))
    r = OS.loss_grad(Xs.double().cpu(), ys.double().cpu(), [b.double() for b in Bn], [b.double() for b in Bc], bias,
                     w.double().cpu(), [False, False, False], 0.01)
    want = torch.cat([t.reshape(-1) for t in r['grad_n'] + r['grad_c']] + [r['dbias'].reshape(-1)])
    assert abs(loss[1].item() - r['loss']) < 1e-5 * abs(r['loss'])
    assert rel(grad, want) < 1e-5
