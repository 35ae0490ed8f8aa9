"""CPU-side checks: the C-ABI library loads and exports every symbol include/tr_b200.h declares,
fails loudly without a device, and the host logic (masks, offsets, shard split, gloo all-reduce
of the packed gradient sums, convergence rule, kwargs validation) behaves like the reference."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tensor_regression_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'tr_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|const char\*)\s+(tr_[a-z_0-9]+)\s*\(', hdr, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.tr_version() == 210


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-device failure mode')
def test_no_cpu_fallback():
    from tensor_regression_b200 import _lib, engine
    from tensor_regression_b200 import standard_tensor_regression as STR
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    h = ctypes.c_void_p()
    dims = (ctypes.c_int64 * 2)(4, 5)
    rc = _lib.lib.tr_create(ctypes.byref(h), 0, 2, dims, 3, 0, 0)
    assert rc != 0 and b'no CPU path' in _lib.lib.tr_last_error(None)
    with pytest.raises(engine.TRError):
        engine.Engine([4, 5], 3)
    with pytest.raises(engine.TRError):
        STR.CP_linear_regression((10, 4, 5), device='cpu')
    with pytest.raises(engine.TRError):
        MTR.CP_logistic_regression(np.zeros((4, 3, 2), np.float32), np.array([0, 1, 0, 1]), device='cpu')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'tensor_regression_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src.replace('no CPU', ''), fn


def test_spectral_signatures_match_reference():
    """spectral_tensor_regression.py: names, positional order and defaults of everything the drop-in re-exposes —
    hard-coded from the reference (spectral:17, 62, 118, 168, 284, 339, 393, 425-437, 541-550, 652-659, 895, 966)
    and, where /root/reference is present, compared with the reference module itself."""
    import inspect
    from tensor_regression_b200 import spectral_tensor_regression as SPR

    def names(f):
        return [p.name for p in inspect.signature(f).parameters.values() if p.kind != p.KEYWORD_ONLY]

    model_sig = ['X', 'Bcp', 'weights', 'non_negative', 'bias', 'softplus_kwargs']
    assert names(SPR.make_BcpInit) == ['B_dims', 'rank', 'non_negative', 'complex_dims', 'scale', 'device', 'dtype']
    for fn in (SPR.lin_model, SPR.spectral_model, SPR.stepwise_latents_model, SPR.stepwise_spectral_model):
        assert names(fn) == model_sig
    C = SPR.CP_linear_regression
    assert names(C.__init__)[1:] == ['X_shape', 'y_shape', 'dtype', 'rank_normal', 'rank_spectral', 'non_negative', 'weights',
                                     'Bcp_init', 'Bcp_init_scale', 'n_complex_dim', 'bias_init', 'device', 'softplus_kwargs']
    assert names(C.fit)[1:] == ['X', 'y', 'lambda_L2', 'max_iter', 'tol', 'patience', 'verbose',
                                'running_loss_logging_interval', 'LBFGS_kwargs']
    assert names(C.fit_Adam)[1:] == ['X', 'y', 'lambda_L2', 'max_iter', 'tol', 'patience', 'verbose', 'plotting_interval',
                                     'Adam_kwargs']
    assert names(C.predict)[1:] == ['X', 'Bcp', 'device', 'plot_pref']
    assert names(C.predict_latents)[1:] == ['X', 'Bcp', 'device', 'plot_pref']
    d = inspect.signature(C.__init__).parameters
    assert d['rank_normal'].default == 1 and d['rank_spectral'].default == 1 and d['n_complex_dim'].default == 0
    d = inspect.signature(C.fit_Adam).parameters
    assert d['lambda_L2'].default == 0.01 and d['max_iter'].default == 1000 and d['plotting_interval'].default == 100
    for name in ['return_Bcp_final', 'detach_Bcp', 'get_params', 'set_params', 'display_params', 'plot_outputs']:
        assert hasattr(C, name)
    from oracle import ref_loader
    if ref_loader.available():
        R = ref_loader.spectral()
        for fn in ['make_BcpInit', 'non_neg_fn', 'lin_model', 'spectral_model', 'stepwise_latents_model',
                   'stepwise_spectral_model', 'L2_penalty']:
            assert names(getattr(SPR, fn)) == names(getattr(R, fn)), fn
        for meth in ['__init__', 'fit', 'fit_Adam', 'predict', 'predict_latents']:
            ours, ref = getattr(C, meth), getattr(R.CP_linear_regression, meth)
            assert names(ours) == names(ref), meth
            po, pr = inspect.signature(ours).parameters, inspect.signature(ref).parameters
            for k in pr:
                if k not in ('self', 'device') and pr[k].default is not inspect.Parameter.empty:
                    assert po[k].default == pr[k].default or (po[k].default is pr[k].default), (meth, k)


def test_signatures_match_reference():
    import inspect
    from tensor_regression_b200 import standard_tensor_regression as STR
    from tensor_regression_b200 import multinomial_tensor_regression as MTR

    def names(f):
        return [p.name for p in inspect.signature(f).parameters.values() if p.kind != p.KEYWORD_ONLY]

    assert names(STR.make_BcpInit) == ['B_dims', 'rank', 'non_negative', 'scale', 'device', 'dtype']
    assert names(STR.lin_model) == ['X', 'Bcp', 'weights', 'non_negative', 'bias', 'softplus_kwargs']
    assert names(STR.CP_linear_regression.__init__)[1:] == ['X_shape', 'dtype', 'rank', 'non_negative', 'weights',
                                                            'Bcp_init', 'Bcp_init_scale', 'bias_init', 'device',
                                                            'softplus_kwargs']
    assert names(STR.CP_linear_regression.fit)[1:] == ['X', 'y', 'lambda_L2', 'max_iter', 'tol', 'patience', 'verbose',
                                                       'running_loss_logging_interval', 'LBFGS_kwargs']
    assert names(STR.CP_linear_regression.fit_Adam)[1:] == ['X', 'y', 'lambda_L2', 'max_iter', 'tol', 'patience',
                                                            'verbose', 'Adam_kwargs']
    assert names(STR.CP_linear_regression.predict)[1:] == ['X', 'Bcp', 'device', 'plot_pref']
    assert names(MTR.model) == ['X', 'Bcp', 'weights', 'non_negative', 'softplus_kwargs']
    assert names(MTR.CP_logistic_regression.__init__)[1:] == ['X', 'y', 'rank', 'non_negative', 'weights', 'Bcp_init',
                                                              'Bcp_init_scale', 'device', 'softplus_kwargs']
    assert names(MTR.CP_logistic_regression.fit)[1:] == ['lambda_L2', 'max_iter', 'tol', 'patience', 'weights',
                                                         'verbose', 'running_loss_logging_interval', 'LBFGS_kwargs']
    assert names(MTR.CP_logistic_regression.fit_Adam)[1:] == ['lambda_L2', 'max_iter', 'tol', 'patience', 'weights',
                                                              'verbose', 'Adam_kwargs']
    assert names(MTR.CP_logistic_regression.predict)[1:] == ['X', 'y_true', 'Bcp', 'device']
    d = inspect.signature(STR.CP_linear_regression.fit).parameters
    assert d['lambda_L2'].default == 0.01 and d['max_iter'].default == 1000 and d['tol'].default == 1e-5
    assert d['patience'].default == 10 and d['running_loss_logging_interval'].default == 10
    for name in ['squeeze_integers', 'confusion_matrix', 'idx_to_oneHot', 'make_BcpInit', 'non_neg_fn', 'model',
                 'L2_penalty', 'CP_logistic_regression']:
        assert hasattr(MTR, name)
    # multinomial_tensor_regression_hierarchical.py: no class weights, predict(plot_pref) (hier:291-299, 385-392, 473)
    from tensor_regression_b200 import multinomial_tensor_regression_hierarchical as HTR
    assert names(HTR.CP_logistic_regression.fit)[1:] == ['lambda_L2', 'max_iter', 'tol', 'patience', 'verbose',
                                                         'running_loss_logging_interval', 'LBFGS_kwargs']
    assert names(HTR.CP_logistic_regression.fit_Adam)[1:] == ['lambda_L2', 'max_iter', 'tol', 'patience', 'verbose',
                                                              'Adam_kwargs']
    assert names(HTR.CP_logistic_regression.predict)[1:] == ['X', 'y_true', 'Bcp', 'device', 'plot_pref']
    assert names(HTR.CP_logistic_regression.__init__)[1:] == names(MTR.CP_logistic_regression.__init__)[1:]
    for name in ['squeeze_integers', 'confusion_matrix', 'idx_to_oneHot', 'make_BcpInit', 'non_neg_fn', 'model',
                 'L2_penalty']:
        assert hasattr(HTR, name)


def test_masks_offsets_shards():
    from tensor_regression_b200 import engine
    assert engine.nn_mask_of([True, False, True, False], 3) == 0b101
    assert engine.nn_mask_of([False, False, True], 2) == 0            # std's unused trailing entry
    sizes, offs = engine.factor_offsets([4, 5, 6], 3, 2)
    assert sizes == [12, 15, 18, 6] and list(offs) == [0, 12, 27, 45, 51]
    for n, w in [(10, 3), (7, 8), (500000, 8), (0, 2)]:
        b = [engine.shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_label_helpers_match_reference_semantics():
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    assert list(MTR.squeeze_integers(np.array([7, 2, 7, 4, 1]))) == [3, 1, 3, 2, 0]
    oh = MTR.idx_to_oneHot(np.array([0, 2, 1]), 3)
    assert oh.shape == (3, 3) and oh[1, 2] == 1 and oh.sum() == 3
    cm = MTR.confusion_matrix(np.array([0, 1, 1, 2]), np.array([0, 1, 2, 2]))
    assert np.allclose(cm.sum(axis=0), 1.0) and cm[1, 2] == 0.5


def test_adam_kwargs_validation():
    from tensor_regression_b200.standard_tensor_regression import _adam_hyper
    h = _adam_hyper({'lr': 0.01, 'amsgrad': True})
    assert h['lr'] == 0.01 and h['amsgrad'] is True and h['betas'] == (0.9, 0.999) and h['eps'] == 1e-8
    with pytest.raises(TypeError):
        _adam_hyper({'lr': 0.01, 'maximize': True})
    with pytest.raises(TypeError):
        _adam_hyper({'bogus': 1})


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from oracle import tr_oracle as O
from tensor_regression_b200 import engine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
d = torch.float64
# standard model: every rank computes the unnormalised local sums of ITS slice with the oracle
X, y, _ = O.synth_std(37, (4, 3, 5), 2, 5, dtype=d)
nn = [False, True, False, False]
B0 = O.init_std((4, 3, 5), 2, nn, dtype=d)
bias = torch.tensor([0.2], dtype=d); w = torch.ones(2, dtype=d)
lo, hi = engine.shard_bounds(X.shape[0], rank, world)
sh = engine.ShardedSum()
assert sh.enabled and sh.world == world
gs = O.closed_form_std(X[lo:hi], y[lo:hi], B0, bias, w, nn)["gradsum"].clone()
sh.sum_(gs)
n_total = sh.total(hi - lo, "cpu")
full = O.closed_form_std(X, y, B0, bias, w, nn)["gradsum"]
assert n_total == X.shape[0]
assert torch.allclose(gs, full, rtol=1e-12, atol=1e-12), (gs - full).abs().max()
grad, ld, lt = O.finish(gs, B0, nn, 0.01, 2.0 / n_total, 1.0 / n_total, True)
ref = O.std_loss_grad(X, y, B0, bias, w, nn, 0.01)
want = torch.cat([g.reshape(-1) for g in ref["grads"]] + [ref["dbias"].reshape(-1)])
assert torch.allclose(grad, want, rtol=1e-10, atol=1e-12)
# multinomial: the normaliser W = sum_n omega[y_n] must be the GLOBAL sum
Xm, ym, _ = O.synth_mn(41, (3, 4), 2, 3, 6)
Xm = Xm.to(d); nnm = [False, False, False]
Bm = [b.to(d) for b in O.init_mn([3, 4, 3], 2, nnm)]
cw = torch.tensor([0.5, 1.0, 2.0], dtype=d)
lo, hi = engine.shard_bounds(Xm.shape[0], rank, world)
cf = O.closed_form_mn(Xm[lo:hi], ym[lo:hi], Bm, w, nnm, cw)
gsm = cf["gradsum"].clone(); sh.sum_(gsm)
W = sh.total(cf["W"].item(), "cpu")
gradm, _, ltm = O.finish(gsm, Bm, nnm, 0.01, 1.0 / W, 1.0 / W, False)
refm = O.mn_loss_grad(Xm, ym, Bm, w, nnm, cw, 0.01)
wantm = torch.cat([g.reshape(-1) for g in refm["grads"]])
assert torch.allclose(gradm, wantm, rtol=1e-10, atol=1e-12)
assert abs(ltm.item() - refm["loss"].item()) < 1e-12
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_sum_gloo_world2(tmp_path):
    """N-sharding host logic on 2 CPU ranks: local unnormalised sums (oracle) + one all-reduce of
    the packed vector + global normaliser == the single-process result."""
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29531', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_create_rejects_unsupported_geometry_loudly():
    """Limits of the kernels are explicit errors with a message, never a silent fallback."""
    from tensor_regression_b200 import _lib
    h = ctypes.c_void_p()

    def create(dtype, dims, R, C):
        arr = (ctypes.c_int64 * len(dims))(*dims)
        rc = _lib.lib.tr_create(ctypes.byref(h), dtype, len(dims), arr, R, C, 0)
        return rc, _lib.lib.tr_last_error(None).decode()

    rc, msg = create(0, [2] * 9, 2, 0)
    assert rc == 3 and 'feature modes' in msg
    rc, msg = create(0, [4, 4], 2, 129)
    assert rc == 3 and 'n_classes' in msg
    rc, msg = create(0, [4, 4], 33, 3)
    assert rc == 3 and 'multinomial rank' in msg
    rc, msg = create(7, [4, 4], 2, 0)
    assert rc == 1 and 'dtype' in msg
    assert _lib.lib.tr_set_option(None, b'fused', 0) != 0
