"""GPU parity tests of the single-pass multinomial cluster kernel (k_fused_mn, csrc/tr_fused_mn.cuh), forced with
``fused=1`` through the C ABI:

* against the closed-form CPU oracle (fp32 at 1e-5, fp64 at 1e-10) over geometries that exercise every index path:
  ragged row split over the cluster, fewer rows than threads, more than one row per thread, every cluster size,
  fewer samples than clusters, one sample, softplus factors, non-unit rank and class weights, rank padded to an
  even channel count;
* against the two-pass kernels and the reference's own outputs (golden fixtures: kernels and 20-iteration fits);
* determinism (second launch bit-identical), shard additivity, chunked long sums;
* at full size on one GPU's shard of BASELINE configs[2] (25 GB).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import tr_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ADAM = {'lr': 0.01, 'amsgrad': True}
DEV = 'cuda:0'
TOL = {torch.float32: 1e-5, torch.float64: 1e-10}


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dev(t, dtype=None):
    t = torch.as_tensor(t)
    return t.to(device=DEV, dtype=dtype or t.dtype).contiguous()


def run_case(N, dims, R, C, dt, cl, seed=0, nn=None, rank_w=None, bal=True):
    from tensor_regression_b200 import engine
    k = len(dims)
    nn = nn or [False] * (k + 1)
    X, y, _ = O.synth_mn(N, dims, R, C, 700 + seed)
    X = X.to(dt)
    B = [b.to(dt) for b in O.init_mn(list(dims) + [C], R, nn, scale=0.4, seed=321 + seed)]
    w = torch.ones(R, dtype=dt) if rank_w is None else torch.tensor(rank_w, dtype=dt)
    cw = torch.tensor(np.random.default_rng(seed).uniform(0.5, 2.0, C) if bal else np.ones(C), dtype=dt)
    want = O.closed_form_mn(X.double(), y, [b.double() for b in B], w.double(), nn, cw.double().numpy())
    eng = engine.Engine(dims, R, C, dt, DEV)
    mask = sum(1 << i for i in range(k + 1) if nn[i])
    Xd, yd, cwd, th, wd = dev(X), dev(y), dev(cw), dev(O.pack(B)), dev(w)
    eng.set_option('fused', 0)
    P2 = torch.empty((N, C), dtype=dt, device=DEV)
    two = eng.fwd_grad_mn(Xd, yd, cwd, th, wd, mask, 50.0, 1.0, P=P2).clone()
    eng.set_option('fused', 1)
    eng.set_option('fused_cl', cl)
    P1 = torch.full((N, C), float('nan'), dtype=dt, device=DEV)
    one = eng.fwd_grad_mn(Xd, yd, cwd, th, wd, mask, 50.0, 1.0, P=P1).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-pass'), info
    if cl:
        assert info['cluster_size'] == cl, info
    tol = TOL[dt]
    assert rel(P1, want['P']) < tol, ('P', rel(P1, want['P']))
    assert rel(one, want['gradsum']) < tol, ('gradsum', rel(one, want['gradsum']), info)
    assert rel(one, two) < tol and rel(P1, P2) < tol
    again = eng.fwd_grad_mn(Xd, yd, cwd, th, wd, mask, 50.0, 1.0)
    assert torch.equal(again, one), 'second launch differs'
    eng.close()
    return info


CASES = [
    # name, N, dims, R, C, dtype, forced cluster size
    ('cfg3_shape_cl16', 40, (100, 50, 20), 6, 10, torch.float32, 16),
    ('cfg5_shape_cl16', 37, (100, 50, 20), 4, 4, torch.float32, 16),
    ('tiny_cl1', 64, (6, 5, 8), 3, 4, torch.float32, 1),
    ('tiny_cl2_ragged', 33, (7, 5, 8), 3, 4, torch.float32, 2),
    ('tiny_cl4_ragged', 50, (7, 3, 12), 5, 7, torch.float32, 4),
    ('rows_gt_threads_cl1', 21, (30, 11, 8), 2, 3, torch.float32, 1),       # 330 rows on one CTA: 3 rows per thread
    ('rows_cl8', 29, (40, 33, 16), 4, 5, torch.float32, 8),
    ('two_modes_cl4', 45, (203, 28), 3, 9, torch.float32, 4),                # k = 2: rows = first mode
    ('five_modes_cl2', 19, (3, 4, 2, 5, 24), 6, 6, torch.float32, 2),
    ('one_sample_cl16', 1, (100, 50, 20), 6, 10, torch.float32, 16),
    ('fewer_samples_than_clusters', 3, (16, 16, 8), 2, 2, torch.float32, 1),
    ('auto_cluster', 70, (64, 32, 20), 6, 10, torch.float32, 0),
    ('f64_cl1', 48, (6, 5, 8), 3, 4, torch.float64, 1),
    ('f64_cl4_ragged', 31, (9, 7, 10), 4, 5, torch.float64, 4),
    ('f64_cl16', 17, (50, 40, 12), 4, 3, torch.float64, 16),
]


@pytest.mark.parametrize('name,N,dims,R,C,dt,cl', CASES, ids=[c[0] for c in CASES])
def test_fused_mn_vs_oracle_and_two_pass(name, N, dims, R, C, dt, cl):
    run_case(N, dims, R, C, dt, cl)


def test_fused_mn_softplus_and_weights():
    run_case(44, (12, 9, 16), 5, 6, torch.float32, 2, seed=3, nn=[True, False, True, True], rank_w=[0.5, 2.0, 1.0, -0.7, 1.3])
    run_case(23, (12, 9, 8), 3, 4, torch.float64, 4, seed=4, nn=[False, True, False, True], rank_w=[1.5, 0.25, -1.0])


def test_fused_mn_not_eligible_fails_loudly_and_auto_falls_back():
    from tensor_regression_b200 import engine
    N, dims, R, C = 20, (7, 5, 6), 3, 4                      # last mode = 6 floats: rows are not 16-byte multiples
    X, y, _ = O.synth_mn(N, dims, R, C, 5)
    B = O.init_mn(list(dims) + [C], R, [False] * 4)
    eng = engine.Engine(dims, R, C, torch.float32, DEV)
    args = (dev(X), dev(y), dev(torch.ones(C)), dev(O.pack(B)), dev(torch.ones(R)), 0, 50.0, 1.0)
    eng.set_option('fused', 1)
    with pytest.raises(engine.TRError):
        eng.fwd_grad_mn(*args)
    eng.set_option('fused', -1)
    eng.fwd_grad_mn(*args)
    assert eng.launch_info()['path'] == 'two-pass'
    eng.close()


def test_fused_mn_shards_add_up_and_long_sums_are_chunked():
    """cnt per cluster > 2048 -> several flush chunks; shard sums == whole (what the all-reduce relies on)."""
    from tensor_regression_b200 import engine
    N, dims, R, C = 700000, (8, 8, 8), 4, 3
    g = torch.Generator(device=DEV).manual_seed(11)
    X = torch.randn((N, *dims), device=DEV, generator=g)
    gc = torch.Generator().manual_seed(12)
    Fs = [0.5 * torch.randn(d, R, generator=gc) for d in list(dims) + [C]]
    B0 = O.init_mn(list(dims) + [C], R, [False] * 4, scale=0.3)
    eng = engine.Engine(dims, R, C, torch.float32, DEV)
    w, cw = dev(torch.ones(R)), dev(torch.tensor([1.0, 0.6, 1.7]))
    _, y = eng.forward_mn(X, dev(O.pack(Fs)), w, 0, 50.0, 1.0)
    th = dev(O.pack(B0))
    eng.set_option('fused', 0)
    two = eng.fwd_grad_mn(X, y, cw, th, w, 0, 50.0, 1.0).clone()
    eng.set_option('fused', 1)
    one = eng.fwd_grad_mn(X, y, cw, th, w, 0, 50.0, 1.0).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-pass') and info['chunks'] >= 2, info
    assert rel(one, two) < 1e-5
    parts = torch.zeros_like(one)
    for r in range(3):
        lo, hi = engine.shard_bounds(N, r, 3)
        parts += eng.fwd_grad_mn(X[lo:hi], y[lo:hi], cw, th, w, 0, 50.0, 1.0)
    assert rel(parts, one) < 1e-5
    # a sub-range against the oracle
    sub = slice(1000, 1200)
    want = O.closed_form_mn(X[sub].cpu().double(), y[sub].cpu(), [b.double() for b in B0], torch.ones(R, dtype=torch.float64),
                            [False] * 4, cw.cpu().double().numpy())['gradsum']
    assert rel(eng.fwd_grad_mn(X[sub], y[sub], cw, th, w, 0, 50.0, 1.0), want) < 1e-5
    eng.close()


ELIGIBLE = ['mn_3mode_bal_nn', 'mn_4mode']


@pytest.mark.parametrize('name', ELIGIBLE)
def test_fused_mn_api_fit_vs_reference_golden(name):
    """20 fit_Adam iterations with every closure evaluation on k_fused_mn == what the UNMODIFIED reference produced."""
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    k = len([n for n in z.files if n.startswith('Bcp_init_')])
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    nn = [bool(v) for v in z['non_negative']]
    wts = None if np.all(z['weights'] == 1) else z['weights']
    m = MTR.CP_logistic_regression(X, y, rank=int(z['R']), non_negative=nn, weights=wts, Bcp_init=[b.clone() for b in B0],
                                   device=DEV)
    m._engine().set_option('fused', 1)
    m.fit_Adam(lambda_L2=float(z['lambda_L2']), max_iter=20, tol=1e-50, patience=100, weights=z['class_weights'],
               Adam_kwargs=ADAM)
    assert m._engine().launch_info()['path'].startswith('single-pass')
    assert rel(m.loss_running, z['adam_loss_running']) < 1e-4
    for i in range(k):
        assert rel(m.Bcp[i], z[f'adam_Bcp_{i}']) < 1e-4
    prob, pred = m.predict()
    assert rel(prob, z['adam_prob']) < 1e-4


def test_fused_mn_hierarchical_fit_vs_reference_golden():
    from tensor_regression_b200 import multinomial_tensor_regression_hierarchical as HTR
    z = np.load(os.path.join(GOLDEN, 'hier_2mode.npz'))
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(3)]
    m = HTR.CP_logistic_regression(X, y, rank=int(z['R']), non_negative=[bool(v) for v in z['non_negative']],
                                   Bcp_init=[b.clone() for b in B0], device=DEV)
    m._engine().set_option('fused', 1)
    m.fit_Adam(lambda_L2=float(z['lambda_L2']), max_iter=20, tol=1e-50, patience=100, Adam_kwargs=ADAM)
    assert m._engine().launch_info()['path'].startswith('single-pass')
    assert rel(m.loss_running, z['adam_loss_running']) < 1e-4
    for i in range(3):
        assert rel(m.Bcp[i], z[f'adam_Bcp_{i}']) < 1e-4


def test_fused_mn_cfg3_full_size_properties():
    """One GPU's shard of configs[2] (62 500 x (100, 50, 20) fp32 = 25 GB, 10 classes, rank 6) on the single-pass
    kernel: == two-pass kernels, == oracle on sub-ranges, deterministic, shard-additive, loss bounds."""
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
    eng = engine.Engine(dims, R, C, torch.float32, DEV)
    w = dev(torch.ones(R))
    theta_star, theta = dev(O.pack(Fs)), dev(O.pack(B0))
    _, y = eng.forward_mn(X, theta_star, w, 0, 50.0, 1.0)
    counts = torch.bincount(y, minlength=C).double()
    cw = dev((N / (C * counts.clamp(min=1))).float())
    eng.set_option('fused', 0)
    P2 = torch.empty((N, C), device=DEV)
    two = eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0, P=P2).clone()
    eng.set_option('fused', 1)
    P1 = torch.empty((N, C), device=DEV)
    one = eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0, P=P1).clone()
    info = eng.launch_info()
    assert info['path'].startswith('single-pass'), info
    assert rel(one, two) < 1e-5 and rel(P1, P2) < 1e-5
    assert torch.equal(eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0), one)
    for sl in (slice(0, 150), slice(N - 150, N)):
        want = O.closed_form_mn(X[sl].cpu().double(), y[sl].cpu(), [b.double() for b in B0],
                                torch.ones(R, dtype=torch.float64), nn, cw.cpu().double().numpy())
        assert rel(P1[sl], want['P']) < 1e-5
        assert rel(eng.fwd_grad_mn(X[sl], y[sl], cw, theta, w, 0, 50.0, 1.0), want['gradsum']) < 1e-5
    parts = torch.zeros_like(one)
    for r in range(3):
        lo, hi = engine.shard_bounds(N, r, 3)
        parts += eng.fwd_grad_mn(X[lo:hi], y[lo:hi], cw, theta, w, 0, 50.0, 1.0)
    assert rel(parts, one) < 1e-5
    ones = dev(torch.ones(C))
    for th in (theta, theta_star):
        ce = float(eng.fwd_grad_mn(X, y, ones, th, w, 0, 50.0, 1.0)[-1]) / N
        assert np.log(1 + (C - 1) / np.e) - 1e-6 <= ce <= np.log(C) + 1e-6, ce
    B_zero = [b.clone() for b in B0]
    B_zero[-1].zero_()
    gz = eng.fwd_grad_mn(X, y, ones, dev(O.pack(B_zero)), w, 0, 50.0, 1.0)
    # (the single-pass kernel's epilogue runs in fp32 like the reference's own arithmetic: log C to fp32 accuracy)
    assert abs(float(gz[-1]) / N - np.log(C)) < 2e-6 and float(gz[:eng.P - C * R].abs().max()) == 0.0
    eng.close()
    del X
    torch.cuda.empty_cache()


def test_fused_mn_automatic_selection_by_row_pitch():
    """Automatic selection (X >= 3 x L2): last modes of an odd number of 16-byte chunks, or six (two-way bank conflicts, still
    ahead of the two-pass kernels), run single-pass; 4 chunks (four-way conflicts) stay on the two-pass kernels."""
    from tensor_regression_b200 import engine
    for dims, single in (((100, 50, 24), True), ((100, 50, 20), True), ((100, 50, 16), False)):
        N = 3000
        X = torch.randn((N, *dims), device=DEV)
        eng = engine.Engine(dims, 6, 10, torch.float32, DEV)
        th = 0.2 * torch.rand(eng.P, device=DEV) - 0.1
        y = torch.randint(0, 10, (N,), device=DEV)
        args = (X, y, torch.ones(10, device=DEV), th, torch.ones(6, device=DEV), 0, 50.0, 1.0)
        one = eng.fwd_grad_mn(*args).clone()
        assert eng.launch_info()['path'].startswith('single-pass') == single, (dims, eng.launch_info())
        eng.set_option('fused', 0)
        two = eng.fwd_grad_mn(*args)
        assert rel(one, two) < 1e-5
        eng.close()
        del X
