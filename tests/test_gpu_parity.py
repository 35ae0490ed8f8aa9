"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and through the
reference-shaped Python API, against (a) the golden outputs of the unmodified reference
(tests/golden), (b) the CPU oracle on freshly seeded inputs, (c) size-independent properties.

Tolerances (north star): norm-relative 1e-5 for fp32, 1e-10 for fp64, i.e.
max|a-b| / max|b| (per-element relative error is meaningless where y_hat ~ 0)."""
import glob
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import tr_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STD = sorted(glob.glob(os.path.join(GOLDEN, 'std_*.npz')))
MN = sorted(glob.glob(os.path.join(GOLDEN, 'mn_*.npz')))
ADAM = {'lr': 0.01, 'amsgrad': True}
LBFGS = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
         'history_size': 100, 'line_search_fn': 'strong_wolfe'}
DEV = 'cuda:0'
TOL = {torch.float32: 1e-5, torch.float64: 1e-10}


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def load(path):
    z = np.load(path)
    k = len([n for n in z.files if n.startswith('Bcp_init_')])
    return z, k


def engine_for(dims, R, C, dtype):
    from tensor_regression_b200 import engine
    return engine.Engine(dims, R, C, dtype, DEV)


def dev(t, dtype=None):
    t = torch.as_tensor(t)
    return t.to(device=DEV, dtype=dtype or t.dtype).contiguous()


_FLOW = []


def flow_built():
    """The experimental dataflow kernel (tr_flow.cuh) is only in builds made with `make FLOW=1`."""
    if not _FLOW:
        from tensor_regression_b200 import engine
        eng = engine.Engine((4, 4), 2, 0, torch.float32, DEV)
        try:
            eng.set_option('flow', 1)
            _FLOW.append(True)
        except engine.TRError:
            _FLOW.append(False)
        eng.close()
    return _FLOW[0]


def need_flow():
    if not flow_built():
        pytest.skip('library built without the experimental dataflow kernel (make FLOW=1)')


# ------------------------------------------------------------------------------------------
# (a) golden fixtures: kernels through the C ABI
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', STD, ids=[os.path.basename(p)[:-4] for p in STD])
def test_std_kernels_vs_reference_golden(path):
    from tensor_regression_b200 import engine
    z, k = load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    dt = X.dtype
    tol = TOL[dt]
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    w = torch.from_numpy(z['weights'])
    nn = [bool(v) for v in z['non_negative']]
    bias = torch.tensor([float(z['bias_init'])], dtype=dt)
    lam, N = float(z['lambda_L2']), X.shape[0]
    eng = engine_for(X.shape[1:], int(z['R']), 0, dt)
    theta = dev(O.pack(B0, bias))
    mask = engine.nn_mask_of(nn, k)
    Xd, yd, wd = dev(X), dev(y), dev(w)
    yhat = eng.forward_std(Xd, theta, wd, mask, 50.0, 1.0)
    assert rel(yhat, z['y_hat']) < tol
    yh2 = torch.empty_like(yd)
    gs = eng.fwd_grad_std(Xd, yd, theta, wd, mask, 50.0, 1.0, yhat=yh2)
    assert rel(yh2, z['y_hat']) < tol
    # unnormalised sums vs the fp64 closed form
    cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn)
    assert rel(gs, cf['gradsum']) < tol
    grad, loss = eng.finish(gs, 2.0 / N, 1.0 / N, theta, lam, mask, 50.0, 1.0)
    want = np.concatenate([z[f'grad_{i}'].reshape(-1) for i in range(k)] + [z['dbias'].reshape(-1)])
    assert rel(grad, want) < tol
    assert abs(loss[0].item() - float(z['loss_data'])) < tol * abs(float(z['loss_data']))
    assert abs(loss[1].item() - float(z['loss'])) < tol * abs(float(z['loss']))


@pytest.mark.parametrize('path', MN, ids=[os.path.basename(p)[:-4] for p in MN])
@pytest.mark.parametrize('dt', [torch.float32, torch.float64], ids=['f32', 'f64'])
def test_mn_kernels_vs_reference_golden(path, dt):
    """fp32 against the reference's own fp32 outputs; the fp64 instantiation of the same kernels
    against the fp64 oracle at 1e-10 (pins the double-softmax / dZ / MTTKRP algebra exactly)."""
    from tensor_regression_b200 import engine
    z, k = load(path)
    X, y = torch.from_numpy(z['X']).to(dt), torch.from_numpy(z['y'])
    tol = TOL[dt]
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']).to(dt) for i in range(k)]
    w = torch.from_numpy(z['weights']).to(dt)
    nn = [bool(v) for v in z['non_negative']]
    cw = torch.from_numpy(z['class_weights']).to(dt)
    lam, C, R = float(z['lambda_L2']), int(z['C']), int(z['R'])
    eng = engine_for(X.shape[1:], R, C, dt)
    theta = dev(O.pack(B0))
    mask = engine.nn_mask_of(nn, k)
    Xd, yd, wd, cwd = dev(X), dev(y), dev(w), dev(cw)
    ref = O.mn_loss_grad(X.double(), y, [b.double() for b in B0], w.double(), nn, cw.double().numpy(), lam)
    cf = O.closed_form_mn(X.double(), y, [b.double() for b in B0], w.double(), nn, cw.double())
    P, pred = eng.forward_mn(Xd, theta, wd, mask, 50.0, 1.0)
    assert rel(P, ref['P']) < tol
    if dt == torch.float32:
        assert rel(P, z['P']) < tol
    assert np.array_equal(pred.cpu().numpy(), np.argmax(P.cpu().numpy(), axis=1))
    P2 = torch.empty_like(P)
    gs = eng.fwd_grad_mn(Xd, yd, cwd, theta, wd, mask, 50.0, 1.0, P=P2)
    assert rel(P2, ref['P']) < tol
    assert rel(gs, cf['gradsum']) < tol
    W = cf['W'].item()
    grad, loss = eng.finish(gs, 1.0 / W, 1.0 / W, theta, lam, mask, 50.0, 1.0)
    want64 = torch.cat([g.reshape(-1) for g in ref['grads']])
    assert rel(grad, want64) < tol
    assert abs(loss[1].item() - ref['loss'].item()) < tol * abs(ref['loss'].item())
    if dt == torch.float32:
        want = np.concatenate([z[f'grad_{i}'].reshape(-1) for i in range(k)])
        assert rel(grad, want) < tol
        assert abs(loss[1].item() - float(z['loss'])) < tol * abs(float(z['loss']))


# ------------------------------------------------------------------------------------------
# (a') golden fixtures: the reference-shaped API (fit_Adam / fit / predict)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', STD, ids=[os.path.basename(p)[:-4] for p in STD])
def test_std_api_fit_vs_reference_golden(path):
    from tensor_regression_b200 import standard_tensor_regression as STR
    z, k = load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    dt = X.dtype
    tol = 10 * TOL[dt]
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    nn = [bool(v) for v in z['non_negative']]
    wts = None if np.all(z['weights'] == 1) else z['weights']
    lam = float(z['lambda_L2'])
    m = STR.CP_linear_regression(X.shape, dtype=dt, rank=int(z['R']), non_negative=nn, weights=wts,
                                 Bcp_init=[b.clone() for b in B0], bias_init=float(z['bias_init']), device=DEV)
    conv = m.fit_Adam(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, verbose=False,
                      Adam_kwargs=ADAM)
    assert conv is False and len(m.loss_running) == 20
    assert rel(m.loss_running, z['adam_loss_running']) < tol
    for i in range(k):
        assert rel(m.Bcp[i], z[f'adam_Bcp_{i}']) < tol
    assert rel(m.bias, z['adam_bias']) < tol
    assert rel(m.predict(z['X']), z['adam_predict']) < tol                  # numpy X: streamed path
    assert rel(m.predict(X.to(DEV)), z['adam_predict']) < tol
    if 'lbfgs_loss_running' in z.files:
        m2 = STR.CP_linear_regression(X.shape, dtype=dt, rank=int(z['R']), non_negative=nn, weights=wts,
                                      Bcp_init=[b.clone() for b in B0], bias_init=float(z['bias_init']), device=DEV)
        m2.fit(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=6, tol=1e-50, patience=10, verbose=False,
               running_loss_logging_interval=1, LBFGS_kwargs=LBFGS)
        assert rel(m2.loss_running, z['lbfgs_loss_running']) < 1e-7
        for i in range(k):
            assert rel(m2.Bcp[i], z[f'lbfgs_Bcp_{i}']) < 1e-6


@pytest.mark.parametrize('path', MN, ids=[os.path.basename(p)[:-4] for p in MN])
def test_mn_api_fit_vs_reference_golden(path):
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    z, k = load(path)
    nn = [bool(v) for v in z['non_negative']]
    wts = None if np.all(z['weights'] == 1) else z['weights']
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    m = MTR.CP_logistic_regression(z['X'], z['y'], rank=int(z['R']), non_negative=nn, weights=wts,
                                   Bcp_init=[b.clone() for b in B0], device=DEV)
    assert m.n_classes == int(z['C'])
    m.fit_Adam(lambda_L2=float(z['lambda_L2']), max_iter=20, tol=1e-50, patience=100, weights=z['class_weights'],
               verbose=False, Adam_kwargs=ADAM)
    assert rel(m.loss_running, z['adam_loss_running']) < 1e-4
    for i in range(k):
        assert rel(m.Bcp[i], z[f'adam_Bcp_{i}']) < 1e-4
    prob, pred = m.predict()
    assert rel(prob, z['adam_prob']) < 1e-4
    assert np.mean(pred == z['adam_pred']) > 0.98
    cm, acc = m.make_confusion_matrix()
    assert cm.shape == (m.n_classes, m.n_classes) and 0 <= acc <= 1
    with pytest.raises((TypeError, RuntimeError)):
        m.fit_Adam(Adam_kwargs=ADAM)                    # class weights None raises like mn:449


def test_hierarchical_api_fit_vs_reference_golden():
    """Drop-in for multinomial_tensor_regression_hierarchical.py: no class weights in the
    signatures, unweighted CE, fitted factors / probabilities equal to the reference's."""
    from tensor_regression_b200 import multinomial_tensor_regression_hierarchical as HTR
    z, k = load(os.path.join(GOLDEN, 'hier_2mode.npz'))
    nn = [bool(v) for v in z['non_negative']]
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    m = HTR.CP_logistic_regression(z['X'], z['y'], rank=int(z['R']), non_negative=nn,
                                   Bcp_init=[b.clone() for b in B0], device=DEV)
    m.fit_Adam(lambda_L2=float(z['lambda_L2']), max_iter=20, tol=1e-50, patience=100, verbose=False, Adam_kwargs=ADAM)
    assert rel(m.loss_running, z['adam_loss_running']) < 1e-4
    for i in range(k):
        assert rel(m.Bcp[i], z[f'adam_Bcp_{i}']) < 1e-4
    prob, pred = m.predict(plot_pref=False)
    assert rel(prob, z['adam_prob']) < 1e-4 and np.mean(pred == z['adam_pred']) > 0.98
    with pytest.raises(TypeError):
        m.fit_Adam(weights=np.ones(3), Adam_kwargs=ADAM)      # the hierarchical signatures have no `weights`
    m4 = HTR.CP_logistic_regression(np.zeros((8, 4, 3, 2), dtype=np.float32), np.arange(8) % 2, rank=2, device=DEV)
    with pytest.raises(IndexError):
        m4.fit_Adam(Adam_kwargs=ADAM)                           # three parameter groups only (hier:436-440)


# ------------------------------------------------------------------------------------------
# (b) oracle on BASELINE.json shapes at sizes the CPU finishes in seconds
# ------------------------------------------------------------------------------------------
SHAPES_STD = [
    ('cfg1_full', 2000, (20, 30, 40), 5, torch.float32),
    ('cfg2_shape', 300, (64, 64, 32), 8, torch.float32),
    ('cfg4_shape_f64', 48, (16, 16, 16, 32), 12, torch.float64),
    ('demo_shape_f64', 16, (500, 500), 10, torch.float64),
    ('ragged_tail', 131, (9, 7, 11), 3, torch.float32),
]


@pytest.mark.parametrize('name,N,dims,R,dt', SHAPES_STD, ids=[s[0] for s in SHAPES_STD])
def test_std_vs_oracle_on_baseline_shapes(name, N, dims, R, dt):
    from tensor_regression_b200 import engine
    X, y, _ = O.synth_std(N, dims, R, 1234 + 2, dtype=dt)
    nn = [False] * (len(dims) + 1)
    B0 = O.init_std(dims, R, nn, dtype=dt)
    bias = torch.tensor([0.0], dtype=dt)
    w = torch.ones(R, dtype=dt)
    eng = engine_for(dims, R, 0, dt)
    theta = dev(O.pack(B0, bias))
    gs = eng.fwd_grad_std(dev(X), dev(y), theta, dev(w), 0, 50.0, 1.0)
    grad, loss = eng.finish(gs, 2.0 / N, 1.0 / N, theta, 0.01, 0, 50.0, 1.0)
    truth = O.std_loss_grad(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn, 0.01)
    want = torch.cat([g.reshape(-1) for g in truth['grads']] + [truth['dbias'].reshape(-1)])
    e_truth = rel(grad, want)
    assert e_truth < TOL[dt], e_truth
    assert abs(loss[1].item() - truth['loss'].item()) < TOL[dt] * abs(truth['loss'].item())
    if dt == torch.float32:
        # error of the reference's own fp32 evaluation against the same fp64 truth, for context
        ref32 = O.std_loss_grad(X, y, B0, bias, w, nn, 0.01)
        got32 = torch.cat([g.reshape(-1) for g in ref32['grads']] + [ref32['dbias'].reshape(-1)])
        print(f'{name}: kernel vs fp64 {e_truth:.2e}; reference fp32 vs fp64 {rel(got32, want):.2e}; '
              f'kernel vs reference fp32 {rel(grad, got32):.2e}')
        assert rel(grad, got32) < TOL[dt]
    info = eng.launch_info()
    assert info['launches'] >= 5 and info['path'] in ('two-pass', 'single-pass (cluster-resident sample)')


SHAPES_MN = [
    ('cfg3_shape', 256, (100, 50, 20), 10, 6),
    ('cfg5_shape', 200, (100, 50, 20), 4, 4),
    ('rank16', 64, (12, 10), 3, 16),
    ('many_classes', 150, (8, 6, 4), 40, 5),
]


@pytest.mark.parametrize('name,N,dims,C,R', SHAPES_MN, ids=[s[0] for s in SHAPES_MN])
def test_mn_vs_oracle_on_baseline_shapes(name, N, dims, C, R):
    from tensor_regression_b200 import engine
    X, y, _ = O.synth_mn(N, dims, R, C, 1234 + 3)
    nn = [False] * (len(dims) + 1)
    B0 = O.init_mn(list(dims) + [C], R, nn, scale=0.2)
    w = torch.ones(R)
    counts = np.bincount(y.numpy(), minlength=C).astype(np.float64)
    cw = torch.tensor(N / (C * np.maximum(counts, 1)), dtype=torch.float32)
    eng = engine_for(dims, R, C, torch.float32)
    theta = dev(O.pack(B0))
    gs = eng.fwd_grad_mn(dev(X), dev(y), dev(cw), theta, dev(w), 0, 50.0, 1.0)
    truth = O.mn_loss_grad(X.double(), y, [b.double() for b in B0], w.double(), nn, cw.double().numpy(), 0.01)
    W = cw[y].double().sum().item()
    grad, loss = eng.finish(gs, 1.0 / W, 1.0 / W, theta, 0.01, 0, 50.0, 1.0)
    want = torch.cat([g.reshape(-1) for g in truth['grads']])
    assert rel(grad, want) < 1e-5, rel(grad, want)
    assert abs(loss[1].item() - truth['loss'].item()) < 1e-5 * abs(truth['loss'].item())


# ------------------------------------------------------------------------------------------
# (c) properties and edge cases
# ------------------------------------------------------------------------------------------
def test_shard_sum_equals_whole():
    """Virtual ranks on one GPU: the packed sums of disjoint sample slices add up to the whole
    (what the NCCL all-reduce relies on), incl. an unaligned slice start (scalar-load path)."""
    from tensor_regression_b200 import engine
    N, dims, R = 1003, (7, 9, 5), 4            # D = 315: rows are not 16-byte multiples
    X, y, _ = O.synth_std(N, dims, R, 77, dtype=torch.float64)
    B0 = O.init_std(dims, R, [False] * 4, dtype=torch.float64)
    eng = engine_for(dims, R, 0, torch.float64)
    theta = dev(O.pack(B0, torch.tensor([0.1], dtype=torch.float64)))
    Xd, yd, wd = dev(X), dev(y), dev(torch.ones(R, dtype=torch.float64))
    whole = eng.fwd_grad_std(Xd, yd, theta, wd, 0, 50.0, 1.0).clone()
    parts = torch.zeros_like(whole)
    for r in range(3):
        lo, hi = engine.shard_bounds(N, r, 3)
        parts += eng.fwd_grad_std(Xd[lo:hi], yd[lo:hi].contiguous(), theta, wd, 0, 50.0, 1.0)
    assert rel(parts, whole) < 1e-12
    empty = eng.fwd_grad_std(Xd[:0], yd[:0].contiguous(), theta, wd, 0, 50.0, 1.0)
    assert float(empty.abs().max()) == 0.0


def test_vector_and_scalar_load_paths_agree():
    from tensor_regression_b200 import engine
    N, dims, R = 257, (8, 8, 8), 3
    X, y, _ = O.synth_std(N + 1, dims, R, 5)
    eng = engine_for(dims, R, 0, torch.float32)
    B0 = O.init_std(dims, R, [False] * 4)
    theta = dev(O.pack(B0, torch.tensor([0.0])))
    wd = dev(torch.ones(R))
    flat = dev(torch.cat([torch.zeros(1), X.reshape(-1)]))
    X_al = flat[1 + 512:].view(N, *dims)                 # 4-byte offset: not 16-byte aligned
    assert X_al.data_ptr() % 16 != 0
    a = eng.forward_std(X_al, theta, wd, 0, 50.0, 1.0)
    assert eng.launch_info()['vector_width'] == 1
    b = eng.forward_std(X_al.clone(), theta, wd, 0, 50.0, 1.0)
    assert eng.launch_info()['vector_width'] == 4
    assert rel(a, b) < 1e-6


def test_backward_is_linear_and_matches_autograd():
    from tensor_regression_b200 import standard_tensor_regression as STR
    N, dims, R = 64, (6, 5, 8), 3
    X, y, _ = O.synth_std(N, dims, R, 9, dtype=torch.float64)
    nn = [True, False, False, False]
    B0 = O.init_std(dims, R, nn, dtype=torch.float64)
    bias = torch.tensor([0.3], dtype=torch.float64)
    w = torch.tensor([0.5, 1.0, 2.0], dtype=torch.float64)
    Bd = [dev(b).requires_grad_(True) for b in B0]
    bd = dev(bias).requires_grad_(True)
    yh = STR.lin_model(dev(X), Bd, dev(w), nn, bd)
    loss = torch.nn.MSELoss()(yh, dev(y)) + 0.01 * STR.L2_penalty(Bd)
    loss.backward()
    ref = O.std_loss_grad(X, y, B0, bias, w, nn, 0.01)
    assert rel(yh, ref['y_hat']) < 1e-12
    for i in range(3):
        assert rel(Bd[i].grad, ref['grads'][i]) < 1e-10
    assert rel(bd.grad, ref['dbias']) < 1e-10


def test_n_smaller_than_unroll_and_single_sample():
    from tensor_regression_b200 import engine
    dims, R = (5, 4, 8), 2
    eng = engine_for(dims, R, 0, torch.float32)
    for N in (1, 2, 3, 5):
        X, y, _ = O.synth_std(N, dims, R, 40 + N)
        B0 = O.init_std(dims, R, [False] * 4)
        theta = dev(O.pack(B0, torch.tensor([0.2])))
        gs = eng.fwd_grad_std(dev(X), dev(y).reshape(-1), theta, dev(torch.ones(R)), 0, 50.0, 1.0)
        cf = O.closed_form_std(X.double(), y.double().reshape(-1), [b.double() for b in B0],
                               torch.tensor([0.2], dtype=torch.float64), torch.ones(R, dtype=torch.float64), [False] * 4)
        assert rel(gs, cf['gradsum']) < 1e-5


def test_large_sample_count_long_sums():
    """Many samples, small D: exercises the chunked fp32 gradient sums (slots) and several groups."""
    from tensor_regression_b200 import engine
    N, dims, R = 60000, (16, 8), 3
    X, y, _ = O.synth_std(N, dims, R, 3)
    B0 = O.init_std(dims, R, [False] * 3)
    eng = engine_for(dims, R, 0, torch.float32)
    theta = dev(O.pack(B0, torch.tensor([0.0])))
    gs = eng.fwd_grad_std(dev(X), dev(y), theta, dev(torch.ones(R)), 0, 50.0, 1.0)
    cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], torch.tensor([0.0], dtype=torch.float64),
                           torch.ones(R, dtype=torch.float64), [False] * 3)
    assert rel(gs, cf['gradsum']) < 1e-5, rel(gs, cf['gradsum'])


def test_errors_and_pickle():
    from tensor_regression_b200 import engine
    from tensor_regression_b200 import standard_tensor_regression as STR
    X, y, _ = O.synth_std(32, (4, 5), 2, 1)
    m = STR.CP_linear_regression(X.shape, rank=2, device=DEV)
    with pytest.raises(TypeError):
        m.fit_Adam(X.to(DEV), y.to(DEV))                       # Adam_kwargs=None raises like std:453
    with pytest.raises(TypeError):
        m.fit(X.to(DEV), y.to(DEV))                            # LBFGS_kwargs=None raises like std:366
    with pytest.raises(engine.TRError):
        m.predict(torch.zeros(3, 4, 6, device=DEV))            # wrong feature dims
    m.fit_Adam(X.to(DEV), y.to(DEV), max_iter=3, Adam_kwargs=ADAM)
    m2 = pickle.loads(pickle.dumps(m))
    assert rel(m2.predict(X.numpy()), m.predict(X.numpy())) == 0.0
    p = m.get_params()
    m3 = STR.CP_linear_regression(X.shape, rank=2, device=DEV)
    m3.set_params(p)
    m3.bias.copy_(m.bias)
    assert rel(m3.predict(X.numpy()), m.predict(X.numpy())) == 0.0
    assert [b.shape for b in m.return_Bcp_final()] == [(4, 2), (5, 2)]


def test_convergence_rule_matches_reference():
    """std:467-470: stop when sum |diff| of the last patience+1 losses < tol."""
    from tensor_regression_b200 import standard_tensor_regression as STR
    X, y, _ = O.synth_std(64, (4, 5), 2, 1)
    nn = [False] * 3
    B0 = O.init_std((4, 5), 2, nn)
    m = STR.CP_linear_regression(X.shape, rank=2, Bcp_init=[b.clone() for b in B0], device=DEV)
    conv = m.fit_Adam(X.to(DEV), y.to(DEV), lambda_L2=0.01, max_iter=400, tol=1e-2, patience=5,
                      Adam_kwargs={'lr': 0.05})
    ref = O.fit_adam_std(X, y, B0, torch.tensor([0.0]), torch.ones(2), nn, 0.01, len(m.loss_running), {'lr': 0.05})
    L = ref['loss_running']
    stop = next((ii for ii in range(len(L)) if ii > 5 and np.sum(np.abs(np.diff(L[ii - 5:ii + 1]))) < 1e-2), None)
    assert conv is True and stop is not None and abs(len(m.loss_running) - (stop + 1)) <= 1


# ------------------------------------------------------------------------------------------
# single-pass cluster kernel (tr_fwd_grad_std, option fused=1) vs two-pass vs oracle
# ------------------------------------------------------------------------------------------
FUSED_CASES = [
    ('cfg2_shape_cl8', 300, (64, 64, 32), 8, torch.float32),
    ('cfg1_cl2', 2000, (20, 30, 40), 5, torch.float32),
    ('cfg4_shape_f64', 40, (16, 16, 16, 32), 12, torch.float64),
    ('tiny_cl1', 257, (8, 6, 16), 3, torch.float32),
    ('single_sample', 1, (64, 64, 32), 2, torch.float32),
    ('fewer_samples_than_clusters', 5, (32, 16, 8), 2, torch.float64),
    ('long_sums_chunked', 40000, (16, 16), 3, torch.float32),
    ('ragged_slices_cl16_demo_shape_f32', 150, (500, 500), 3, torch.float32),     # 62 500 chunks over 16 CTAs
    ('ragged_slices_f64', 70, (300, 333), 2, torch.float64),                      # 49 950 chunks over 16 CTAs
]


@pytest.mark.parametrize('name,N,dims,R,dt', FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_single_pass_matches_two_pass_and_oracle(name, N, dims, R, dt):
    X, y, _ = O.synth_std(N, dims, R, 1234 + 7, dtype=dt)
    y = y.reshape(-1)
    nn = [True] + [False] * len(dims)
    B0 = O.init_std(dims, R, nn, dtype=dt)
    bias = torch.tensor([0.07], dtype=dt)
    w = torch.linspace(0.5, 1.5, R, dtype=dt)
    eng = engine_for(dims, R, 0, dt)
    theta = dev(O.pack(B0, bias))
    Xd, yd, wd = dev(X), dev(y), dev(w)
    eng.set_option('fused', 0)
    yh2 = torch.empty_like(yd)
    two = eng.fwd_grad_std(Xd, yd, theta, wd, 1, 50.0, 1.0, yhat=yh2).clone()
    assert eng.launch_info()['path'] == 'two-pass'
    eng.set_option('fused', 1)
    yh1 = torch.empty_like(yd)
    one = eng.fwd_grad_std(Xd, yd, theta, wd, 1, 50.0, 1.0, yhat=yh1).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-pass'), info
    cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn)
    tol = TOL[dt]
    assert rel(yh1, cf['y_hat']) < tol and rel(yh2, cf['y_hat']) < tol
    assert rel(one, cf['gradsum']) < tol, (rel(one, cf['gradsum']), info)
    assert rel(two, cf['gradsum']) < tol
    assert rel(one, two) < tol
    # deterministic: same bits on a second launch
    again = eng.fwd_grad_std(Xd, yd, theta, wd, 1, 50.0, 1.0)
    assert torch.equal(again, one)
    eng.set_option('fused', -1)


@pytest.mark.parametrize('cl', [2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 15, 16])
@pytest.mark.parametrize('dims,R,dt', [((64, 64, 32), 4, torch.float32), ((20, 30, 41 * 4), 3, torch.float32),
                                       ((16, 16, 16, 32), 5, torch.float64)],
                         ids=['cfg2_shape', 'ragged_24600_chunks', 'cfg4_shape_f64'])
def test_fused_single_pass_any_cluster_size(cl, dims, R, dt):
    """The cluster size is not restricted to powers of two (the automatic choice takes the size that covers the most SMs,
    e.g. 15 clusters of 9 instead of 15 clusters of 8 on cfg 2): every size, forced, against the two-pass kernels."""
    from tensor_regression_b200 import engine
    N = 97
    X, y, _ = O.synth_std(N, dims, R, 4321, dtype=dt)
    y = y.reshape(-1)
    nn = [False] * (len(dims) + 1)
    B0 = O.init_std(dims, R, nn, dtype=dt)
    bias = torch.tensor([-0.03], dtype=dt)
    w = torch.linspace(0.8, 1.2, R, dtype=dt)
    eng = engine_for(dims, R, 0, dt)
    theta = dev(O.pack(B0, bias))
    Xd, yd, wd = dev(X), dev(y), dev(w)
    eng.set_option('fused', 0)
    yh2 = torch.empty_like(yd)
    two = eng.fwd_grad_std(Xd, yd, theta, wd, 0, 50.0, 1.0, yhat=yh2).clone()
    eng.set_option('fused_cl', cl)
    eng.set_option('fused', -1)
    try:
        # does this size fit (three stages, resident clusters)?  fused=1 on a size that does not fails loudly
        eng.set_option('fused', 1)
        yh1 = torch.full_like(yd, float('nan'))
        try:
            one = eng.fwd_grad_std(Xd, yd, theta, wd, 0, 50.0, 1.0, yhat=yh1).clone()
        except engine.TRError:
            pytest.skip(f'cluster size {cl} does not fit this geometry')
        info = eng.launch_info()
        assert info['path'].startswith('single-pass') and info['cluster_size'] == cl, info
        tol = TOL[dt]
        assert rel(yh1, yh2) < tol and rel(one, two) < tol
        again = eng.fwd_grad_std(Xd, yd, theta, wd, 0, 50.0, 1.0)
        assert torch.equal(again, one)
    finally:
        eng.set_option('fused_cl', 0)
        eng.set_option('fused', -1)


# ------------------------------------------------------------------------------------------
# single-launch dataflow kernel (option flow=1: forward, epilogue and gradient in one persistent
# kernel, second read of X from L2) vs two-pass vs oracle
# ------------------------------------------------------------------------------------------
FLOW_STD = [
    ('cfg2_shape', 300, (64, 64, 32), 8, torch.float32),
    ('cfg1', 2000, (20, 30, 40), 5, torch.float32),
    ('cfg4_shape_f64', 40, (16, 16, 16, 32), 12, torch.float64),
    ('tiny', 257, (8, 6, 16), 3, torch.float32),
    ('single_sample', 1, (64, 64, 32), 2, torch.float32),
    ('long_sums_chunked', 40000, (16, 16), 3, torch.float32),
    ('ragged_tiles', 333, (10, 10, 12), 4, torch.float32),
]


@pytest.mark.parametrize('name,N,dims,R,dt', FLOW_STD, ids=[c[0] for c in FLOW_STD])
@pytest.mark.parametrize('window', [32, 1], ids=['win32', 'win1'])
def test_flow_std_matches_two_pass_and_oracle(name, N, dims, R, dt, window):
    need_flow()
    X, y, _ = O.synth_std(N, dims, R, 1234 + 7, dtype=dt)
    y = y.reshape(-1)
    nn = [True] + [False] * len(dims)
    B0 = O.init_std(dims, R, nn, dtype=dt)
    bias = torch.tensor([0.07], dtype=dt)
    w = torch.linspace(0.5, 1.5, R, dtype=dt)
    eng = engine_for(dims, R, 0, dt)
    theta = dev(O.pack(B0, bias))
    Xd, yd, wd = dev(X), dev(y), dev(w)
    eng.set_option('fused', 0)
    eng.set_option('flow', 0)
    yh2 = torch.empty_like(yd)
    two = eng.fwd_grad_std(Xd, yd, theta, wd, 1, 50.0, 1.0, yhat=yh2).clone()
    assert eng.launch_info()['path'] == 'two-pass'
    eng.set_option('flow', 1)
    eng.set_option('flow_window_mb', window)
    yh1 = torch.empty_like(yd)
    one = eng.fwd_grad_std(Xd, yd, theta, wd, 1, 50.0, 1.0, yhat=yh1).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-launch dataflow'), info
    cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn)
    tol = TOL[dt]
    assert rel(yh1, cf['y_hat']) < tol
    assert rel(one, cf['gradsum']) < tol, (rel(one, cf['gradsum']), info)
    assert rel(one, two) < tol
    again = eng.fwd_grad_std(Xd, yd, theta, wd, 1, 50.0, 1.0)          # deterministic: same bits
    assert torch.equal(again, one)


FLOW_MN = [
    ('cfg3_shape', 256, (100, 50, 20), 10, 6, torch.float32),
    ('cfg5_shape', 200, (100, 50, 20), 4, 4, torch.float32),
    ('rank16', 64, (12, 10), 3, 16, torch.float32),
    ('many_classes', 150, (8, 6, 4), 40, 5, torch.float32),
    ('padded_channels_f64', 90, (12, 8, 4), 7, 5, torch.float64),
    ('long', 30000, (10, 8), 3, 2, torch.float32),
]


@pytest.mark.parametrize('name,N,dims,C,R,dt', FLOW_MN, ids=[c[0] for c in FLOW_MN])
def test_flow_mn_matches_two_pass_and_oracle(name, N, dims, C, R, dt):
    need_flow()
    X, y, _ = O.synth_mn(N, dims, R, C, 1234 + 3)
    X = X.to(dt)
    nn = [False, True] + [False] * (len(dims) - 1)
    B0 = [b.to(dt) for b in O.init_mn(list(dims) + [C], R, nn, scale=0.2)]
    w = torch.linspace(0.7, 1.2, R, dtype=dt)
    counts = np.bincount(y.numpy(), minlength=C).astype(np.float64)
    cw = torch.tensor(N / (C * np.maximum(counts, 1)), dtype=dt)
    eng = engine_for(dims, R, C, dt)
    theta = dev(O.pack(B0))
    Xd, yd, cwd, wd = dev(X), dev(y), dev(cw), dev(w)
    eng.set_option('flow', 0)
    P2 = torch.empty((N, C), dtype=dt, device=DEV)
    two = eng.fwd_grad_mn(Xd, yd, cwd, theta, wd, 2, 50.0, 1.0, P=P2).clone()
    assert eng.launch_info()['path'] == 'two-pass'
    eng.set_option('flow', 1)
    P1 = torch.empty((N, C), dtype=dt, device=DEV)
    one = eng.fwd_grad_mn(Xd, yd, cwd, theta, wd, 2, 50.0, 1.0, P=P1).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-launch dataflow'), info
    cf = O.closed_form_mn(X.double(), y, [b.double() for b in B0], w.double(), nn, cw.double().numpy())
    tol = TOL[dt]
    assert rel(P1, cf['P']) < tol
    assert rel(one, cf['gradsum']) < tol, (rel(one, cf['gradsum']), info)
    assert rel(one, two) < tol
    assert torch.equal(P1, P2)                                            # same epilogue arithmetic
    again = eng.fwd_grad_mn(Xd, yd, cwd, theta, wd, 2, 50.0, 1.0)
    assert torch.equal(again, one)


def test_flow_not_eligible_fails_loudly():
    need_flow()
    from tensor_regression_b200 import engine
    dims, R = (5, 7, 3), 2                       # D = 105: rows are not 16-byte multiples
    X, y, _ = O.synth_std(64, dims, R, 3)
    eng = engine_for(dims, R, 0, torch.float32)
    theta = dev(O.pack(O.init_std(dims, R, [False] * 4), torch.tensor([0.0])))
    eng.set_option('fused', 0)
    eng.set_option('flow', 1)
    with pytest.raises(engine.TRError):
        eng.fwd_grad_std(dev(X), dev(y).reshape(-1), theta, dev(torch.ones(R)), 0, 50.0, 1.0)


def test_fused_not_eligible_falls_back_loudly():
    from tensor_regression_b200 import engine
    dims, R = (5, 7, 3), 2                       # D = 105: rows are not 16-byte multiples
    X, y, _ = O.synth_std(64, dims, R, 3)
    eng = engine_for(dims, R, 0, torch.float32)
    theta = dev(O.pack(O.init_std(dims, R, [False] * 4), torch.tensor([0.0])))
    eng.set_option('fused', 1)
    with pytest.raises(engine.TRError):
        eng.fwd_grad_std(dev(X), dev(y).reshape(-1), theta, dev(torch.ones(R)), 0, 50.0, 1.0)
    eng.set_option('fused', -1)
    eng.fwd_grad_std(dev(X), dev(y).reshape(-1), theta, dev(torch.ones(R)), 0, 50.0, 1.0)
    assert eng.launch_info()['path'] == 'two-pass'


def test_mn_lbfgs_fit_tracks_reference_algorithm():
    """fit (L-BFGS) of the multinomial class in fp32: strong-Wolfe branches may flip on fp32 noise
    (SURVEY H3), so the check is on the logged losses, not the factors."""
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    X, y, _ = O.synth_mn(160, (6, 5, 4), 3, 4, 77)
    nn = [False] * 4
    B0 = O.init_mn([6, 5, 4, 4], 3, nn, scale=0.5)
    cw = np.ones(4, dtype=np.float32)
    ref = O.fit_lbfgs_mn(X, y, B0, torch.ones(3), nn, cw, 0.01, 4, 1e-50, 10, LBFGS)
    m = MTR.CP_logistic_regression(X, y, rank=3, Bcp_init=[b.clone() for b in B0], device=DEV)
    m.fit(lambda_L2=0.01, max_iter=4, tol=1e-50, patience=10, weights=cw, running_loss_logging_interval=1,
          LBFGS_kwargs=LBFGS)
    assert len(m.loss_running) == len(ref['loss_running']) == 4
    assert rel(m.loss_running[:2], ref['loss_running'][:2]) < 1e-4
    assert m.loss_running[-1] < m.loss_running[0]
    assert abs(m.loss_running[-1] - ref['loss_running'][-1]) < 2e-2 * abs(ref['loss_running'][-1])


# ------------------------------------------------------------------------------------------
# out-of-bounds detection without compute-sanitizer (closed on this pool): NaN poison around every
# input the kernels read, sentinels around every output they write
# ------------------------------------------------------------------------------------------
def _guarded(t, pad=4096, poison=float('nan')):
    """Return (view, whole): `view` has t's contents and sits in the middle of a poisoned buffer."""
    flat = torch.full((t.numel() + 2 * pad,), poison, dtype=t.dtype, device=DEV) if t.is_floating_point() else \
        torch.full((t.numel() + 2 * pad,), -1, dtype=t.dtype, device=DEV)
    flat[pad:pad + t.numel()] = t.reshape(-1).to(DEV)
    return flat[pad:pad + t.numel()].view(t.shape), flat


@pytest.mark.parametrize('fused', [0, 1], ids=['two_pass', 'single_pass'])
@pytest.mark.parametrize('N,dims', [(37, (8, 4, 8)), (19, (5, 7, 3)), (3, (64, 64, 32))], ids=['small', 'odd_D', 'cfg2_shape'])
def test_std_kernels_do_not_touch_memory_outside_their_buffers(fused, N, dims):
    from tensor_regression_b200 import engine
    R = 3
    D = int(np.prod(dims))
    if fused and D % 4:
        pytest.skip('single-pass kernel needs 16-byte rows')
    X, y, _ = O.synth_std(N, dims, R, 5)
    y = y.reshape(-1)
    B0 = O.init_std(dims, R, [False] * (len(dims) + 1))
    eng = engine_for(dims, R, 0, torch.float32)
    eng.set_option('fused', fused)
    pad = 4096
    Xg, Xw = _guarded(X, pad)
    yg, _ = _guarded(y, pad)
    thg, _ = _guarded(O.pack(B0, torch.tensor([0.1])), pad)
    wg, _ = _guarded(torch.ones(R), pad)
    SENT = -777.0
    gs_w = torch.full((eng.n_gradsum + 2 * pad,), SENT, dtype=torch.float64, device=DEV)
    yh_w = torch.full((N + 2 * pad,), SENT, dtype=torch.float32, device=DEV)
    gs, yh = gs_w[pad:pad + eng.n_gradsum], yh_w[pad:pad + N]
    eng.fwd_grad_std(Xg, yg, thg, wg, 0, 50.0, 1.0, gradsum=gs, yhat=yh)
    torch.cuda.synchronize()
    cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], torch.tensor([0.1], dtype=torch.float64),
                           torch.ones(R, dtype=torch.float64), [False] * (len(dims) + 1))
    assert torch.isfinite(gs).all() and torch.isfinite(yh).all()          # a poisoned read would show up as NaN
    assert rel(gs, cf['gradsum']) < 1e-5 and rel(yh, cf['y_hat']) < 1e-5
    for w in (gs_w, yh_w):                                                # sentinels around the outputs intact
        assert bool((w[:pad] == SENT).all()) and bool((w[-pad:] == SENT).all())
    assert bool(torch.isnan(Xw[:pad]).all()) and bool(torch.isnan(Xw[-pad:]).all())
    eng.set_option('fused', -1)


def test_mn_kernels_do_not_touch_memory_outside_their_buffers():
    from tensor_regression_b200 import engine
    N, dims, C, R = 29, (6, 5, 4), 3, 5
    X, y, _ = O.synth_mn(N, dims, R, C, 6)
    B0 = O.init_mn(list(dims) + [C], R, [False] * 4, scale=0.5)
    eng = engine_for(dims, R, C, torch.float32)
    pad = 4096
    Xg, _ = _guarded(X, pad)
    yg, _ = _guarded(y, pad)
    thg, _ = _guarded(O.pack(B0), pad)
    wg, _ = _guarded(torch.ones(R), pad)
    cwg, _ = _guarded(torch.ones(C), pad)
    SENT = -777.0
    gs_w = torch.full((eng.n_gradsum + 2 * pad,), SENT, dtype=torch.float64, device=DEV)
    P_w = torch.full((N * C + 2 * pad,), SENT, dtype=torch.float32, device=DEV)
    gs, P = gs_w[pad:pad + eng.n_gradsum], P_w[pad:pad + N * C].view(N, C)
    eng.fwd_grad_mn(Xg, yg, cwg, thg, wg, 0, 50.0, 1.0, gradsum=gs, P=P)
    torch.cuda.synchronize()
    cf = O.closed_form_mn(X.double(), y, [b.double() for b in B0], torch.ones(R, dtype=torch.float64), [False] * 4,
                          torch.ones(C, dtype=torch.float64))
    assert torch.isfinite(gs).all() and torch.isfinite(P).all()
    assert rel(gs, cf['gradsum']) < 1e-5 and rel(P, cf['P']) < 1e-5
    for w in (gs_w, P_w):
        assert bool((w[:pad] == SENT).all()) and bool((w[-pad:] == SENT).all())


# ------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at FULL size (N=200000, 64x64x32, rank 8, fp32: 104.9 GB resident in HBM),
# through size-independent properties (the CPU oracle cannot finish this size in seconds)
# ------------------------------------------------------------------------------------------
def test_cfg2_full_size_properties():
    from tensor_regression_b200 import engine
    free, _ = torch.cuda.mem_get_info()
    N, dims, R = 200000, (64, 64, 32), 8
    D = int(np.prod(dims))
    if free < N * D * 4 + (8 << 30):
        pytest.skip('needs ~113 GB of free HBM')
    g = torch.Generator(device=DEV).manual_seed(2024)
    X = torch.empty((N, *dims), dtype=torch.float32, device=DEV)
    for lo in range(0, N, 2048):
        X[lo:lo + 2048].normal_(generator=g)
    nn = [False] * 4
    B0 = O.init_std(dims, R, nn)
    Fs = [0.3 * torch.randn(d, R, generator=torch.Generator().manual_seed(7)) for d in dims]
    eng = engine_for(dims, R, 0, torch.float32)
    w = dev(torch.ones(R))
    theta_star = dev(O.pack(Fs, torch.tensor([0.1])))
    theta = dev(O.pack(B0, torch.tensor([0.0])))
    y = eng.forward_std(X, theta_star, w, 0, 50.0, 1.0)
    # (1) forward at full size == oracle on a regenerated-by-copy sub-range (first / last 300 samples)
    for sl in (slice(0, 300), slice(N - 300, N)):
        Xs = X[sl].cpu()
        want = O.lin_model(Xs.double(), [f.double() for f in Fs], torch.ones(R, dtype=torch.float64), nn,
                           torch.tensor([0.1], dtype=torch.float64))
        assert rel(y[sl], want) < 1e-5
    # (2) single-pass == two-pass at full size, and both deterministic
    eng.set_option('fused', 1)
    yh1 = torch.empty_like(y)
    one = eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0, yhat=yh1).clone()
    assert eng.launch_info()['path'].startswith('single-pass')
    eng.set_option('fused', 0)
    yh2 = torch.empty_like(y)
    two = eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0, yhat=yh2).clone()
    assert rel(one, two) < 1e-5 and rel(yh1, yh2) < 1e-5
    # (3) the packed sums of disjoint shards add up to the whole (what the all-reduce relies on)
    parts = torch.zeros_like(two)
    for r in range(3):
        lo, hi = engine.shard_bounds(N, r, 3)
        eng.set_option('fused', r % 2)
        parts += eng.fwd_grad_std(X[lo:hi], y[lo:hi], theta, w, 0, 50.0, 1.0)
    assert rel(parts, two) < 1e-5
    # (4) linearity of the gradient pass in the residual: backward(a*r1 + r2) == a*backward(r1) + backward(r2)
    r1 = torch.randn(N, device=DEV, generator=g)
    r2 = torch.randn(N, device=DEV, generator=g)
    b1 = eng.backward_std(X, r1, theta, w, 0, 50.0, 1.0).clone()
    b2 = eng.backward_std(X, r2, theta, w, 0, 50.0, 1.0).clone()
    b12 = eng.backward_std(X, (0.5 * r1 + r2).contiguous(), theta, w, 0, 50.0, 1.0)
    assert rel(b12, 0.5 * b1 + b2) < 1e-5
    # (5) the gradient at the generating factors is (numerically) zero relative to the gradient at the init
    g_star = eng.fwd_grad_std(X, y, theta_star, w, 0, 50.0, 1.0)
    assert float(g_star[:-1].abs().max()) < 1e-4 * float(two[:-1].abs().max())
    eng.set_option('fused', -1)
    del X
    torch.cuda.empty_cache()


def test_cfg3_full_size_properties():
    """One GPU's shard of configs[2] (62 500 x (100, 50, 20) fp32 = 25 GB, 10 classes, rank 6): what can be
    checked at a size the oracle cannot run."""
    from tensor_regression_b200 import engine
    free, _ = torch.cuda.mem_get_info()
    N, dims, C, R = 62500, (100, 50, 20), 10, 6
    D = int(np.prod(dims))
    if free < N * D * 4 + (8 << 30):
        pytest.skip('needs ~33 GB of free HBM')
    g = torch.Generator(device=DEV).manual_seed(2025)
    X = torch.empty((N, *dims), dtype=torch.float32, device=DEV)
    for lo in range(0, N, 2048):
        X[lo:lo + 2048].normal_(generator=g)
    nn = [False] * 4
    gc = torch.Generator().manual_seed(11)
    Fs = [0.3 * torch.randn(d, R, generator=gc) for d in list(dims) + [C]]
    B0 = O.init_mn(list(dims) + [C], R, nn, scale=0.2)
    eng = engine_for(dims, R, C, torch.float32)
    eng.set_option('fused', 0)      # the two-pass kernels (fp64 epilogue); the single-pass kernel has its own full-size test
    w = dev(torch.ones(R))
    theta_star, theta = dev(O.pack(Fs)), dev(O.pack(B0))
    P, y = eng.forward_mn(X, theta_star, w, 0, 50.0, 1.0)
    # (1) probabilities at full size == oracle on the first / last 200 samples; rows sum to one; argmax == labels
    for sl in (slice(0, 200), slice(N - 200, N)):
        want = O.mn_model(X[sl].cpu().double(), [f.double() for f in Fs], torch.ones(R, dtype=torch.float64), nn)
        assert rel(P[sl], want) < 1e-5
    assert float((P.sum(1) - 1).abs().max()) < 1e-5 and torch.equal(P.argmax(1), y)
    counts = torch.bincount(y, minlength=C).double()
    cw = dev((N / (C * counts.clamp(min=1))).float())
    # (2) deterministic: same bits on a second launch
    full = eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0).clone()
    assert torch.equal(eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0), full)
    # (3) the packed sums of disjoint shards add up to the whole (what the all-reduce relies on)
    parts = torch.zeros_like(full)
    for r in range(3):
        lo, hi = engine.shard_bounds(N, r, 3)
        parts += eng.fwd_grad_mn(X[lo:hi], y[lo:hi], cw, theta, w, 0, 50.0, 1.0)
    assert rel(parts, full) < 1e-5
    # (4) the vector-Jacobian product is linear in the upstream gradient
    d1 = torch.randn((N, C), device=DEV, generator=g)
    d2 = torch.randn((N, C), device=DEV, generator=g)
    b1 = eng.backward_mn(X, d1, theta, w, 0, 50.0, 1.0).clone()
    b2 = eng.backward_mn(X, d2, theta, w, 0, 50.0, 1.0).clone()
    b12 = eng.backward_mn(X, (0.5 * d1 + d2).contiguous(), theta, w, 0, 50.0, 1.0)
    assert rel(b12[:-1], 0.5 * b1[:-1] + b2[:-1]) < 1e-5
    # (5) loss bounds of the double softmax: log(1 + (C-1)/e) <= CE <= log C at unit class weights
    ones = dev(torch.ones(C))
    for th in (theta, theta_star):
        gs = eng.fwd_grad_mn(X, y, ones, th, w, 0, 50.0, 1.0)
        ce = float(gs[-1]) / N
        assert np.log(1 + (C - 1) / np.e) - 1e-6 <= ce <= np.log(C) + 1e-6, ce
    # (6) with a zero class factor every logit is zero: P uniform, CE == log C, feature-factor gradients vanish
    B_zero = [b.clone() for b in B0]
    B_zero[-1].zero_()
    gz = eng.fwd_grad_mn(X, y, ones, dev(O.pack(B_zero)), w, 0, 50.0, 1.0)
    assert abs(float(gz[-1]) / N - np.log(C)) < 1e-9 and float(gz[:eng.P - C * R].abs().max()) == 0.0
    del X
    torch.cuda.empty_cache()


def test_mn_model_autograd_matches_reference_autograd():
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    N, dims, C, R = 48, (5, 4, 6), 4, 3
    X, y, _ = O.synth_mn(N, dims, R, C, 31)
    X = X.double()
    nn = [False, True, False, False]
    B0 = [b.double() for b in O.init_mn(list(dims) + [C], R, nn, scale=0.6)]
    w = torch.tensor([0.5, 1.0, 2.0], dtype=torch.float64)
    cw = torch.tensor([1.0, 0.5, 2.0, 1.5], dtype=torch.float64)
    Bd = [dev(b).requires_grad_(True) for b in B0]
    P = MTR.model(dev(X), Bd, dev(w), nn)
    loss = torch.nn.CrossEntropyLoss(weight=dev(cw))(P, dev(y)) + 0.01 * MTR.L2_penalty(Bd)
    loss.backward()
    ref = O.mn_loss_grad(X, y, B0, w, nn, cw.numpy(), 0.01)
    assert rel(P, ref['P']) < 1e-12
    assert abs(loss.item() - ref['loss'].item()) < 1e-12
    for i in range(4):
        assert rel(Bd[i].grad, ref['grads'][i]) < 1e-10


def test_upload_resident_places_every_chunk():
    """Host data of a resident fit goes up through pinned staging in chunks: float64 numpy (converted), a pinned
    fp32 tensor (copied in place), a chunk size that does not divide N, an array-like that is neither."""
    from tensor_regression_b200 import engine
    g = np.random.default_rng(5)
    A = g.standard_normal((37, 6, 10))
    d = engine.upload_resident(A, torch.float32, DEV, chunk_bytes=5 * 6 * 10 * 4)
    assert torch.equal(d.cpu(), torch.from_numpy(A).float())
    B = torch.from_numpy(A).float().pin_memory()
    assert torch.equal(engine.upload_resident(B, torch.float32, DEV, chunk_bytes=1 << 12).cpu(), B)

    class View:                       # anything with .shape and [lo:hi]
        shape = A.shape

        def __getitem__(self, sl):
            return A[sl]
    assert torch.equal(engine.upload_resident(View(), torch.float64, DEV, chunk_bytes=999).cpu(), torch.from_numpy(A))


def test_mn_out_of_core_equals_resident():
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    N, dims, C, R = 333, (6, 5, 8), 3, 4
    X, y, _ = O.synth_mn(N, dims, R, C, 32)
    B0 = O.init_mn(list(dims) + [C], R, [False] * 4, scale=0.5)
    cw = np.array([1.0, 2.0, 0.5], dtype=np.float32)
    a = MTR.CP_logistic_regression(X, y, rank=R, Bcp_init=[b.clone() for b in B0], device=DEV)
    a.fit_Adam(lambda_L2=0.01, max_iter=8, tol=1e-50, patience=100, weights=cw, Adam_kwargs=ADAM)
    b = MTR.CP_logistic_regression(X.numpy(), y.numpy(), rank=R, Bcp_init=[b.clone() for b in B0], device=DEV,
                                   out_of_core=True, chunk_samples=50)
    assert not isinstance(b.X, torch.Tensor)
    b.fit_Adam(lambda_L2=0.01, max_iter=8, tol=1e-50, patience=100, weights=cw, Adam_kwargs=ADAM)
    assert rel(b.loss_running, a.loss_running) < 1e-6
    for i in range(4):
        assert rel(b.Bcp[i], a.Bcp[i]) < 1e-5
    pa, _ = a.predict()
    pb, _ = b.predict()
    assert rel(pb, pa) < 1e-5
    c = STR_out_of_core_check()
    assert c < 1e-6


def STR_out_of_core_check():
    """standard model: out_of_core fit_Adam (streamed from a numpy array) == resident fit_Adam."""
    from tensor_regression_b200 import standard_tensor_regression as STR
    X, y, _ = O.synth_std(257, (8, 6, 4), 3, 33)
    B0 = O.init_std((8, 6, 4), 3, [False] * 4)
    a = STR.CP_linear_regression(X.shape, rank=3, Bcp_init=[b.clone() for b in B0], device=DEV)
    a.fit_Adam(X.to(DEV), y.to(DEV), max_iter=6, tol=1e-50, patience=100, Adam_kwargs=ADAM)
    b = STR.CP_linear_regression(X.shape, rank=3, Bcp_init=[b.clone() for b in B0], device=DEV)
    b.fit_Adam(X.numpy(), y.numpy(), max_iter=6, tol=1e-50, patience=100, Adam_kwargs=ADAM, out_of_core=True,
               chunk_samples=40)
    return rel(b.loss_running, a.loss_running)


def test_notebook_known_answer_through_the_cuda_path():
    """The reference's only known-answer vector (demo_TensorRegression.ipynb cells 5 + 8: fp64, seeds 321,
    rank 10, L-BFGS strong-Wolfe, saved log 560125.5196947237, 1699.8925874402807, 0.041904340578888165 x 11,
    'Convergence reached'), reproduced with the notebook's own calls on the CUDA path at the notebook's full
    size (X is 2000 x 500 x 500 float64 = 4 GB)."""
    import scipy.signal
    from oracle.tensorly_standin import cp_to_tensor, inner
    from tensor_regression_b200 import standard_tensor_regression as STR
    torch.manual_seed(321)
    np.random.seed(321)
    dims = [2000, 500, 500]
    Xcp = [torch.rand(dims[0], 4) - 0.5,
           torch.vstack([torch.sin(torch.linspace(0, 140, dims[1])),
                         torch.cos(torch.linspace(2, 19, dims[1])),
                         torch.linspace(0, 1, dims[1]),
                         torch.cos(torch.linspace(0, 17, dims[1])) > 0]).T,
           torch.tensor(scipy.signal.savgol_filter(np.random.rand(dims[2], 4), 15, 3, axis=0)) - 0.5]
    X_fake = cp_to_tensor((np.ones(4), Xcp))
    y = inner(X_fake + torch.rand(dims) / 100, cp_to_tensor((np.ones(4), Xcp[1:])), n_modes=2)
    X = X_fake - X_fake.mean(0)
    del X_fake
    assert X.dtype == torch.float64
    # cell 8, verbatim apart from the device
    cpmlr = STR.CP_linear_regression(X.shape, dtype=X.dtype, rank=10, non_negative=[False, False], weights=None,
                                     Bcp_init=None, Bcp_init_scale=0.005, device=DEV, softplus_kwargs={'beta': 50, 'threshold': 1})
    conv = cpmlr.fit(X, y, lambda_L2=1e-5, max_iter=200, tol=1e-50, patience=10, verbose=0,
                     running_loss_logging_interval=1, LBFGS_kwargs=LBFGS)
    L = cpmlr.loss_running
    assert conv is True and len(L) == 13, (conv, len(L), L)
    assert abs(L[0] - 560125.5196947237) / 560125.5196947237 < 1e-3      # (the noise term of y is not seeded identically)
    assert abs(L[-1] - 0.041904340578888165) / 0.041904340578888165 < 1e-5
    assert max(abs(v - L[2]) for v in L[2:]) < 1e-9 * L[2]
    y_hat = cpmlr.predict(X[:64])
    assert y_hat.shape == (64,) and np.all(np.isfinite(y_hat))


def test_randomised_geometry_sweep_vs_oracle():
    """40 seeded random geometries per model (1-5 feature modes, odd sizes, ranks that need channel padding,
    softplus masks, non-unit rank / class weights, N from 1 to a few hundred): packed sums through the C ABI
    against the closed-form oracle in float64, and the fp32 kernels against the same at 1e-5."""
    rng = np.random.default_rng(20261018)
    for case in range(40):
        k = int(rng.integers(1, 6))
        dims = tuple(int(rng.integers(2, 9 if k > 2 else 24)) for _ in range(k))
        R = int(rng.integers(1, 17))
        N = int(rng.choice([1, 2, 3, 7, 33, 130, 257]))
        dt = torch.float64 if case % 4 == 0 else torch.float32
        tol = TOL[dt]
        nn = [bool(rng.integers(0, 2)) for _ in range(k + 1)]
        mask = sum(1 << i for i in range(k) if nn[i])
        # ---- standard
        X, y, _ = O.synth_std(N, dims, R, 9000 + case, dtype=dt)
        y = y.reshape(-1)
        B0 = O.init_std(dims, R, nn, dtype=dt)
        bias = torch.tensor([float(rng.normal())], dtype=dt)
        w = torch.tensor(rng.uniform(0.5, 1.5, R), dtype=dt)
        eng = engine_for(dims, R, 0, dt)
        gs = eng.fwd_grad_std(dev(X), dev(y), dev(O.pack(B0, bias)), dev(w), mask, 50.0, 1.0)
        cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn)
        assert rel(gs, cf['gradsum']) < tol, ('std', case, dims, R, N, dt, rel(gs, cf['gradsum']))
        # the single-pass cluster kernel and the dataflow kernel on the same geometry, where they are eligible
        from tensor_regression_b200 import engine as _E
        for opt in ('fused', 'flow'):
            eng.set_option('fused', 0)
            try:
                eng.set_option(opt, 1)
                alt = eng.fwd_grad_std(dev(X), dev(y), dev(O.pack(B0, bias)), dev(w), mask, 50.0, 1.0)
            except _E.TRError:
                alt = None                      # geometry / alignment not eligible: loud refusal, not a fallback
            eng.set_option(opt, 0)
            if alt is not None:
                assert rel(alt, cf['gradsum']) < tol, (opt, case, dims, R, N, dt, rel(alt, cf['gradsum']))
        # ---- multinomial (class factor last; its softplus flag is nn[k])
        C = int(rng.integers(2, 12))
        Xm, ym, _ = O.synth_mn(N, dims, R, C, 9500 + case)
        Xm = Xm.to(dt)
        Bm = [b.to(dt) for b in O.init_mn(list(dims) + [C], R, nn, scale=0.3)]
        cw = torch.tensor(rng.uniform(0.5, 2.0, C), dtype=dt)
        mask_mn = sum(1 << i for i in range(k + 1) if nn[i])
        engm = engine_for(dims, R, C, dt)
        P = torch.empty((N, C), dtype=dt, device=DEV)
        gm = engm.fwd_grad_mn(dev(Xm), dev(ym), dev(cw), dev(O.pack(Bm)), dev(w), mask_mn, 50.0, 1.0, P=P)
        cm = O.closed_form_mn(Xm.double(), ym, [b.double() for b in Bm], w.double(), nn, cw.double().numpy())
        assert rel(P, cm['P']) < tol, ('mn P', case, dims, R, C, N, dt, rel(P, cm['P']))
        assert rel(gm, cm['gradsum']) < tol, ('mn', case, dims, R, C, N, dt, rel(gm, cm['gradsum']))
        try:
            engm.set_option('flow', 1)
            gf = engm.fwd_grad_mn(dev(Xm), dev(ym), dev(cw), dev(O.pack(Bm)), dev(w), mask_mn, 50.0, 1.0)
        except _E.TRError:
            gf = None
        if gf is not None:
            assert rel(gf, cm['gradsum']) < tol, ('mn flow', case, dims, R, C, N, dt, rel(gf, cm['gradsum']))


def test_sample_larger_than_one_wave_of_warp_tiles():
    """D so large that a sample has more warp tiles than the grid has warps (the persistent kernels then
    loop over tiles): (1300, 1000) = 1.3 M elements per sample, both models, all eligible paths."""
    from tensor_regression_b200 import engine as _E
    dims, R, N = (1300, 1000), 2, 6
    X, y, _ = O.synth_std(N, dims, R, 77)
    y = y.reshape(-1)
    nn = [False, False, False]
    B0 = O.init_std(dims, R, nn)
    bias, w = torch.tensor([0.3]), torch.ones(R)
    eng = engine_for(dims, R, 0, torch.float32)
    cf = O.closed_form_std(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn)
    for opt in (None, 'fused', 'flow'):
        eng.set_option('fused', 0)
        eng.set_option('flow', 0)
        try:
            if opt:
                eng.set_option(opt, 1)
            yh = torch.empty(N, device=DEV)
            gs = eng.fwd_grad_std(dev(X), dev(y), dev(O.pack(B0, bias)), dev(w), 0, 50.0, 1.0, yhat=yh)
        except _E.TRError:
            assert opt is not None           # the plain two-pass path must take any geometry
            continue
        assert rel(yh, cf['y_hat']) < 1e-5 and rel(gs, cf['gradsum']) < 1e-5, (opt, eng.launch_info())
    C = 3
    Xm, ym, _ = O.synth_mn(N, dims, R, C, 78)
    Bm = O.init_mn(list(dims) + [C], R, nn + [False], scale=0.05)
    engm = engine_for(dims, R, C, torch.float32)
    cm = O.closed_form_mn(Xm.double(), ym, [b.double() for b in Bm], w.double(), nn + [False], np.ones(C))
    gm = engm.fwd_grad_mn(dev(Xm), dev(ym), dev(torch.ones(C)), dev(O.pack(Bm)), dev(w), 0, 50.0, 1.0)
    assert rel(gm, cm['gradsum']) < 1e-5, engm.launch_info()
