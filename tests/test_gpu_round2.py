"""GPU parity tests added in round 2 (all through the C ABI / the reference-shaped API):

* K-iteration fits on the single-pass cluster kernel (the kernel bench.py times) against the reference's own
  fit outputs (golden fixtures) — fit_Adam (20 iterations) and fit (L-BFGS, fp64);
* two engines that use the same kernel instantiation with different shared-memory sizes, alternated;
* per-group learning rates (tr_adam_step_groups) against torch.optim.Adam with parameter groups;
* tr_allreduce / tr_comm_* on a one-rank communicator;
* label range check, out-of-core L-BFGS == resident L-BFGS;
* full-size properties for BASELINE configs[3] (fp64, 105 GB) and one GPU's shard of configs[4] (125 GB).
"""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from oracle import tr_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STD = sorted(glob.glob(os.path.join(GOLDEN, 'std_*.npz')))
ADAM = {'lr': 0.01, 'amsgrad': True}
LBFGS = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
         'history_size': 100, 'line_search_fn': 'strong_wolfe'}
DEV = 'cuda:0'
TOL = {torch.float32: 1e-5, torch.float64: 1e-10}


def rel(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dev(t, dtype=None):
    t = torch.as_tensor(t)
    return t.to(device=DEV, dtype=dtype or t.dtype).contiguous()


def load(path):
    z = np.load(path)
    k = len([n for n in z.files if n.startswith('Bcp_init_')])
    return z, k


# ------------------------------------------------------------------------------------------
# fits on the single-pass kernel vs the reference's fit outputs
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', STD, ids=[os.path.basename(p)[:-4] for p in STD])
def test_std_api_fit_on_single_pass_kernel_vs_reference_golden(path):
    """fit_Adam (20 iterations) and fit (L-BFGS) with fused=1 forced on the model's engine: every closure
    evaluation runs k_fused_std, the result must match what the UNMODIFIED reference produced."""
    from tensor_regression_b200 import standard_tensor_regression as STR
    z, k = load(path)
    X, y = torch.from_numpy(z['X']), torch.from_numpy(z['y'])
    dt = X.dtype
    D = int(np.prod(X.shape[1:]))
    if (D * X.element_size()) % 16 != 0:
        pytest.skip('sample rows are not 16-byte aligned: not eligible for the single-pass kernel')
    tol = 10 * TOL[dt]
    B0 = [torch.from_numpy(z[f'Bcp_init_{i}']) for i in range(k)]
    nn = [bool(v) for v in z['non_negative']]
    wts = None if np.all(z['weights'] == 1) else z['weights']
    lam = float(z['lambda_L2'])

    def make():
        m = STR.CP_linear_regression(X.shape, dtype=dt, rank=int(z['R']), non_negative=nn, weights=wts,
                                     Bcp_init=[b.clone() for b in B0], bias_init=float(z['bias_init']), device=DEV)
        m._engine().set_option('fused', 1)
        return m

    m = make()
    m.fit_Adam(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, Adam_kwargs=ADAM)
    assert m._engine().launch_info()['path'].startswith('single-pass')
    assert rel(m.loss_running, z['adam_loss_running']) < tol
    for i in range(k):
        assert rel(m.Bcp[i], z[f'adam_Bcp_{i}']) < tol
    assert rel(m.bias, z['adam_bias']) < tol
    if 'lbfgs_loss_running' in z.files:
        m2 = make()
        m2.fit(X.to(DEV), y.to(DEV), lambda_L2=lam, max_iter=6, tol=1e-50, patience=10,
               running_loss_logging_interval=1, LBFGS_kwargs=LBFGS)
        assert rel(m2.loss_running, z['lbfgs_loss_running']) < 1e-7
        for i in range(k):
            assert rel(m2.Bcp[i], z[f'lbfgs_Bcp_{i}']) < 1e-6


def test_two_engines_sharing_a_kernel_with_different_shared_memory():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per (device, kernel).  Same dims, two ranks ->
    same k_fused_std / k_fwd instantiation, different dynamic shared memory.  Alternating the engines must
    keep working (the library only ever raises the attribute)."""
    from tensor_regression_b200 import engine
    dims, N = (64, 64, 32), 24
    X = torch.randn((N, *dims), device=DEV)
    y = torch.randn(N, device=DEV)
    engs, thetas, ws, wants = [], [], [], []
    for R in (48, 2, 100):          # factor rows + rank weights: 31 KB, 1.3 KB, 64 KB of the dynamic shared memory
        eng = engine.Engine(dims, R, 0, torch.float32, DEV)
        g = torch.Generator().manual_seed(R)
        B = [0.1 * torch.randn(d, R, generator=g) for d in dims]
        theta = dev(O.pack(B, torch.tensor([0.05])))
        w = dev(torch.ones(R))
        want = O.closed_form_std(X.cpu().double(), y.cpu().double(), [b.double() for b in B],
                                 torch.tensor([0.05], dtype=torch.float64), torch.ones(R, dtype=torch.float64),
                                 [False] * 4)['gradsum']
        engs.append(eng), thetas.append(theta), ws.append(w), wants.append(want)
    for rnd in range(3):
        for fused in (1, 0):
            for eng, theta, w, want in zip(engs, thetas, ws, wants):
                eng.set_option('fused', fused)
                gs = eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0)
                assert rel(gs, want) < 1e-5, (rnd, fused, eng.rank)
    for eng in engs:
        eng.close()


def test_adam_step_groups_matches_torch_parameter_groups():
    from tensor_regression_b200 import engine
    dims, R, C = (6, 5), 3, 4
    eng = engine.Engine(dims, R, C, torch.float32, DEV)
    sizes, offs = engine.factor_offsets(dims, R, C)
    torch.manual_seed(5)
    theta0 = torch.randn(eng.P)
    lrs = [0.01, 0.003, 0.05]
    params = [theta0[offs[i]:offs[i + 1]].clone().requires_grad_(True) for i in range(3)]
    opt = torch.optim.Adam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], amsgrad=True)
    theta = dev(theta0.clone())
    m, v, vm = torch.zeros_like(theta), torch.zeros_like(theta), torch.zeros_like(theta)
    for step in range(1, 6):
        g = torch.randn(eng.P, generator=torch.Generator().manual_seed(100 + step))
        for i, p in enumerate(params):
            p.grad = g[offs[i]:offs[i + 1]].clone()
        opt.step()
        eng.adam_step(theta, dev(g), m, v, vm, step, lr_groups=lrs)
    want = torch.cat([p.detach() for p in params])
    assert rel(theta, want) < 1e-6
    with pytest.raises(engine.TRError):
        eng.adam_step(theta, dev(g), m, v, vm, 6, lr_groups=[0.1, 0.1])          # wrong number of groups
    # standard model: k factors + bias
    eng2 = engine.Engine(dims, R, 0, torch.float64, DEV)
    th = dev(torch.randn(eng2.P, dtype=torch.float64))
    before = th.clone()
    g = dev(torch.ones(eng2.P, dtype=torch.float64))
    z = torch.zeros_like(th)
    eng2.adam_step(th, g, z.clone(), z.clone(), None, 1, lr_groups=[0.1, 0.2, 0.4])
    step = (before - th).cpu()
    assert torch.allclose(step[:18], torch.full((18,), 0.1, dtype=torch.float64), rtol=1e-6)
    assert torch.allclose(step[18:33], torch.full((15,), 0.2, dtype=torch.float64), rtol=1e-6)
    assert abs(float(step[33]) - 0.4) < 1e-6


def test_hierarchical_per_factor_learning_rates():
    """hier:436-440 with DIFFERENT rates per group == torch.optim.Adam with three parameter groups on the oracle."""
    from tensor_regression_b200 import multinomial_tensor_regression_hierarchical as HTR
    N, dims, C, R = 64, (6, 5), 3, 2
    X, y, _ = O.synth_mn(N, dims, R, C, 77)
    nn = [False] * 3
    B0 = O.init_mn(list(dims) + [C], R, nn)
    lrs = [0.02, 0.005, 0.01]
    B = [b.clone().requires_grad_(True) for b in B0]
    opt = torch.optim.Adam([{'params': [b], 'lr': lr} for b, lr in zip(B, lrs)], lr=0.01, amsgrad=True)
    lossf = torch.nn.CrossEntropyLoss()
    want = []
    for _ in range(8):
        opt.zero_grad()
        loss = lossf(O.mn_model(X, B, torch.ones(R), nn), y) + 0.01 * O.L2_penalty(B)
        loss.backward()
        opt.step()
        want.append(loss.item())
    m = HTR.CP_logistic_regression(X, y, rank=R, Bcp_init=[b.clone() for b in B0], device=DEV)
    m.lr_groups = lrs
    m.fit_Adam(lambda_L2=0.01, max_iter=8, tol=1e-50, patience=100, Adam_kwargs={'lr': 0.01, 'amsgrad': True})
    assert rel(m.loss_running, want) < 1e-5
    for i in range(3):
        assert rel(m.Bcp[i], B[i]) < 1e-4


def test_allreduce_through_the_c_abi_one_rank_communicator():
    """tr_comm_unique_id / tr_comm_create / tr_allreduce / tr_comm_destroy with world = 1: the sum over one rank
    is the identity; exercises the run-time NCCL binding a non-Python host would use."""
    from tensor_regression_b200 import _lib, engine
    eng = engine.Engine((4, 5), 2, 0, torch.float64, DEV)
    uid = (ctypes.c_char * 128)()
    rc = _lib.lib.tr_comm_unique_id(uid)
    assert rc == 0, _lib.lib.tr_last_error(None)
    comm = ctypes.c_void_p()
    rc = _lib.lib.tr_comm_create(ctypes.byref(comm), uid, 0, 1, 0)
    assert rc == 0, _lib.lib.tr_last_error(None)
    buf = dev(torch.arange(eng.n_gradsum, dtype=torch.float64) * 0.5 - 3)
    want = buf.clone()
    eng.allreduce(buf, comm.value)
    torch.cuda.synchronize()
    assert torch.equal(buf, want)
    assert _lib.lib.tr_comm_destroy(comm) == 0
    with pytest.raises(engine.TRError):
        eng.allreduce(buf, 0)


@pytest.mark.parametrize('R', [20, 32])
def test_mn_rank_above_16_runs_on_the_wide_channel_kernels(R):
    """multinomial rank 17..32 (24- and 32-channel streaming kernels, one block per SM)."""
    from tensor_regression_b200 import engine
    N, dims, C = 70, (9, 6, 8), 5
    X, y, _ = O.synth_mn(N, dims, R, C, 55)
    nn = [False, True, False, False]
    B = O.init_mn(list(dims) + [C], R, nn, scale=0.4)
    cw = torch.tensor([1.0, 0.5, 2.0, 1.5, 0.8])
    for dt in (torch.float32, torch.float64):
        want = O.closed_form_mn(X.double(), y, [b.double() for b in B], torch.ones(R, dtype=torch.float64), nn,
                                cw.double().numpy())
        eng = engine.Engine(dims, R, C, dt, DEV)
        P = torch.empty((N, C), dtype=dt, device=DEV)
        gs = eng.fwd_grad_mn(dev(X.to(dt)), dev(y), dev(cw.to(dt)), dev(O.pack([b.to(dt) for b in B])),
                             dev(torch.ones(R, dtype=dt)), 2, 50.0, 1.0, P=P)
        assert eng.launch_info()['channels'] in (24, 32)
        assert rel(P, want['P']) < TOL[dt] and rel(gs, want['gradsum']) < TOL[dt]
        eng.close()


def test_labels_outside_the_class_range_raise():
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    X = np.random.default_rng(0).standard_normal((12, 4, 3)).astype(np.float32)
    with pytest.raises(ValueError):
        MTR.CP_logistic_regression(X, np.array([1, 2, 3] * 4), rank=2, device=DEV)           # not consecutive from 0
    with pytest.raises(ValueError):
        MTR.CP_logistic_regression(X, np.array([0, 1, -1] * 4), rank=2, device=DEV, n_classes=2)
    MTR.CP_logistic_regression(X, MTR.squeeze_integers(np.array([1, 2, 3] * 4)), rank=2, device=DEV)


def test_std_out_of_core_lbfgs_equals_resident():
    from tensor_regression_b200 import standard_tensor_regression as STR
    N, dims, R = 90, (6, 4, 8), 3
    X, y, _ = O.synth_std(N, dims, R, 5, dtype=torch.float64)
    nn = [False, True, False, False]
    B0 = O.init_std(dims, R, nn, dtype=torch.float64)

    def run(**kw):
        m = STR.CP_linear_regression(X.shape, dtype=torch.float64, rank=R, non_negative=nn,
                                     Bcp_init=[b.clone() for b in B0], device=DEV)
        m.fit(X.numpy() if kw else X.to(DEV), y, lambda_L2=0.01, max_iter=5, tol=1e-50, patience=10,
              running_loss_logging_interval=1, LBFGS_kwargs=LBFGS, **kw)
        return m

    a, b = run(), run(out_of_core=True, chunk_samples=32)
    assert rel(b.loss_running, a.loss_running) < 1e-9
    for i in range(3):
        assert rel(b.Bcp[i], a.Bcp[i]) < 1e-8
    assert b.h2d_bytes_per_iteration == N * 6 * 4 * 8 * 8


# ------------------------------------------------------------------------------------------
# BASELINE configs[3] (fp64, 5-mode: N=100000 x (16,16,16,32), rank 12, 104.9 GB) at FULL size
# ------------------------------------------------------------------------------------------
def test_cfg4_full_size_properties():
    from tensor_regression_b200 import engine
    free, _ = torch.cuda.mem_get_info()
    N, dims, R = 100000, (16, 16, 16, 32), 12
    D = int(np.prod(dims))
    if free < N * D * 8 + (8 << 30):
        pytest.skip('needs ~113 GB of free HBM')
    dt = torch.float64
    g = torch.Generator(device=DEV).manual_seed(2026)
    X = torch.empty((N, *dims), dtype=dt, device=DEV)
    for lo in range(0, N, 1024):
        X[lo:lo + 1024].normal_(generator=g)
    nn = [False] * 5
    B0 = O.init_std(dims, R, nn, dtype=dt)
    Fs = [0.3 * torch.randn(d, R, generator=torch.Generator().manual_seed(9), dtype=dt) for d in dims]
    eng = engine.Engine(dims, R, 0, dt, DEV)
    w = dev(torch.ones(R, dtype=dt))
    theta_star = dev(O.pack(Fs, torch.tensor([0.1], dtype=dt)))
    theta = dev(O.pack(B0, torch.tensor([0.0], dtype=dt)))
    y = eng.forward_std(X, theta_star, w, 0, 50.0, 1.0)
    # (1) forward at full size == oracle on the first / last 100 samples, at the fp64 tolerance
    for sl in (slice(0, 100), slice(N - 100, N)):
        want = O.lin_model(X[sl].cpu(), Fs, torch.ones(R, dtype=dt), nn, torch.tensor([0.1], dtype=dt))
        assert rel(y[sl], want) < 1e-10
    # (2) gradient sums of a sub-range == closed-form oracle (1e-10), on both kernel paths
    sub = slice(N // 2, N // 2 + 64)
    want = O.closed_form_std(X[sub].cpu(), y[sub].cpu(), B0, torch.tensor([0.0], dtype=dt), torch.ones(R, dtype=dt),
                             nn)['gradsum']
    for fused in (0, 1):
        eng.set_option('fused', fused)
        assert rel(eng.fwd_grad_std(X[sub], y[sub], theta, w, 0, 50.0, 1.0), want) < 1e-10
    # (3) single-pass (16-CTA clusters) == two-pass at full size to fp64 round-off; second launch bit-identical
    eng.set_option('fused', 1)
    one = eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0).clone()
    assert eng.launch_info()['path'].startswith('single-pass')
    assert torch.equal(eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0), one)
    eng.set_option('fused', 0)
    two = eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0).clone()
    assert rel(one, two) < 1e-10
    # (4) shard sums add up
    parts = torch.zeros_like(two)
    for r in range(3):
        lo, hi = engine.shard_bounds(N, r, 3)
        eng.set_option('fused', r % 2)
        parts += eng.fwd_grad_std(X[lo:hi], y[lo:hi], theta, w, 0, 50.0, 1.0)
    assert rel(parts, two) < 1e-10
    # (5) the gradient at the generating factors vanishes (y was generated from them without noise)
    g_star = eng.fwd_grad_std(X, y, theta_star, w, 0, 50.0, 1.0)
    assert float(g_star[:-1].abs().max()) < 1e-9 * float(two[:-1].abs().max())
    eng.close()
    del X
    torch.cuda.empty_cache()


def test_cfg5_full_size_properties():
    """One GPU's shard of configs[4] (312 500 x (100, 50, 20) fp32 = 125 GB, 4 classes, rank 4)."""
    from tensor_regression_b200 import engine
    free, _ = torch.cuda.mem_get_info()
    N, dims, C, R = 312500, (100, 50, 20), 4, 4
    D = int(np.prod(dims))
    if free < N * D * 4 + (8 << 30):
        pytest.skip('needs ~133 GB of free HBM')
    g = torch.Generator(device=DEV).manual_seed(2027)
    X = torch.empty((N, *dims), dtype=torch.float32, device=DEV)
    for lo in range(0, N, 2048):
        X[lo:lo + 2048].normal_(generator=g)
    nn = [False] * 4
    gc = torch.Generator().manual_seed(13)
    Fs = [0.3 * torch.randn(d, R, generator=gc) for d in list(dims) + [C]]
    B0 = O.init_mn(list(dims) + [C], R, nn, scale=0.2)
    eng = engine.Engine(dims, R, C, torch.float32, DEV)
    w = dev(torch.ones(R))
    theta_star, theta = dev(O.pack(Fs)), dev(O.pack(B0))
    P, y = eng.forward_mn(X, theta_star, w, 0, 50.0, 1.0)
    paths = set()
    for sl in (slice(0, 200), slice(N - 200, N)):
        want = O.mn_model(X[sl].cpu().double(), [f.double() for f in Fs], torch.ones(R, dtype=torch.float64), nn)
        assert rel(P[sl], want) < 1e-5
    assert float((P.sum(1) - 1).abs().max()) < 1e-5 and torch.equal(P.argmax(1), y)
    counts = torch.bincount(y, minlength=C).double()
    cw = dev((N / (C * counts.clamp(min=1))).float())
    # gradient sums of a sub-range == closed-form oracle
    sub = slice(N // 3, N // 3 + 96)
    want = O.closed_form_mn(X[sub].cpu().double(), y[sub].cpu(), [b.double() for b in B0],
                            torch.ones(R, dtype=torch.float64), nn, cw.cpu().double().numpy())['gradsum']
    assert rel(eng.fwd_grad_mn(X[sub], y[sub], cw, theta, w, 0, 50.0, 1.0), want) < 1e-5
    fulls = {}
    for fused in (0, 1):
        eng.set_option('fused', fused)
        full = eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0).clone()
        paths.add(eng.launch_info()['path'])
        assert torch.equal(eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0), full)      # deterministic
        parts = torch.zeros_like(full)
        for r in range(3):
            lo, hi = engine.shard_bounds(N, r, 3)
            parts += eng.fwd_grad_mn(X[lo:hi], y[lo:hi], cw, theta, w, 0, 50.0, 1.0)
        assert rel(parts, full) < 1e-5
        fulls[fused] = full
    assert rel(fulls[1], fulls[0]) < 1e-5 and len(paths) == 2, paths         # single-pass == two-pass at 125 GB
    eng.set_option('fused', -1)
    ones = dev(torch.ones(C))
    gs = eng.fwd_grad_mn(X, y, ones, theta, w, 0, 50.0, 1.0)
    ce = float(gs[-1]) / N
    assert np.log(1 + (C - 1) / np.e) - 1e-6 <= ce <= np.log(C) + 1e-6, ce
    eng.close()
    del X
    torch.cuda.empty_cache()
