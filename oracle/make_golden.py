"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Each fixture stores the seeded inputs and what the reference's own code returned for them:
one closure evaluation (prediction, loss, factor gradients — std:368-373 / mn:357-362) and a
fixed-iteration ``fit_Adam`` / ``fit`` run through the reference's estimator classes
(std:400-476, std:305-398, mn:389-471).  /root/reference cannot travel to the GPU box; these
small files can.  Test infrastructure only.
"""
import copy
import os
import sys

import numpy as np
import torch

from . import ref_loader
from . import tr_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

ADAM = {'lr': 0.01, 'amsgrad': True}
LBFGS = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
         'history_size': 100, 'line_search_fn': 'strong_wolfe'}

STD_CASES = [
    # name, N, dims, R, dtype, non_negative (len k+1 as std:281-284), weights, lambda, seed
    ('std_3mode_f32', 64, (6, 5, 8), 3, torch.float32, [False, False, False, False], None, 0.01, 11),
    ('std_3mode_f32_nn', 48, (7, 3, 8), 4, torch.float32, [True, False, True, False], [1.0, 0.5, 2.0, 1.5], 0.02, 12),
    ('std_4mode_f64_nn', 40, (4, 3, 5, 6), 5, torch.float64, [True, False, False, True, False], [0.7, 1.3, 1.0, 2.0, 0.4], 0.01, 13),
    ('std_2mode_f64', 50, (9, 10), 2, torch.float64, [False, False, False], None, 0.001, 14),
    ('std_1mode_f32', 30, (12,), 3, torch.float32, [False, False], None, 0.01, 15),
    ('std_odd_f32', 33, (5, 7, 3), 2, torch.float32, [False, False, False, False], None, 0.01, 16),   # D=105: no 16-byte rows
]

MN_CASES = [
    # name, N, dims, C, R, non_negative (len k+1), rank weights, class weights mode, lambda, seed
    ('mn_3mode', 96, (5, 4, 6), 4, 3, [False, False, False, False], None, 'ones', 0.01, 21),
    ('mn_3mode_bal_nn', 80, (4, 6, 8), 3, 4, [True, False, False, True], [1.0, 0.5, 2.0, 1.5], 'balanced', 0.02, 22),
    ('mn_2mode', 60, (7, 5), 5, 2, [False, False, False], None, 'balanced', 0.01, 23),
    ('mn_4mode', 40, (3, 4, 2, 8), 6, 6, [False, False, False, False, False], None, 'ones', 0.01, 24),
]


def np_list(ts):
    return {f'{i}': t.detach().cpu().numpy() for i, t in enumerate(ts)}


def gen_std(STR):
    for name, N, dims, R, dtype, nn, w, lam, seed in STD_CASES:
        X, y, _ = O.synth_std(N, dims, R, 1234 + seed, dtype=dtype)
        init = O.init_std(dims, R, nn, scale=1.0, dtype=dtype, seed=321)
        bias0 = 0.05
        weights = torch.ones(R, dtype=dtype) if w is None else torch.tensor(w, dtype=dtype)
        # one closure evaluation with the reference's own module-level functions
        B = [b.clone().requires_grad_(True) for b in init]
        bias = torch.tensor([bias0], dtype=dtype, requires_grad=True)
        y_hat = STR.lin_model(X, B, weights, nn, bias)
        mse = torch.nn.MSELoss()(y_hat, y)
        loss = mse + lam * STR.L2_penalty(B)
        loss.backward()
        out = {'X': X.numpy(), 'y': y.numpy(), 'weights': weights.numpy(), 'non_negative': np.array(nn),
               'lambda_L2': lam, 'bias_init': bias0, 'R': R,
               'y_hat': y_hat.detach().numpy(), 'loss_data': mse.item(), 'loss': loss.item(),
               'dbias': bias.grad.numpy()}
        for i, b in enumerate(init):
            out[f'Bcp_init_{i}'] = b.numpy()
            out[f'grad_{i}'] = B[i].grad.numpy()
        # fixed-iteration fit_Adam through the reference estimator
        m = STR.CP_linear_regression(X.shape, dtype=dtype, rank=R, non_negative=nn,
                                     weights=None if w is None else np.array(w),
                                     Bcp_init=[b.clone().requires_grad_(True) for b in init],
                                     bias_init=bias0, device='cpu')
        m.fit_Adam(X, y, lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, verbose=False, Adam_kwargs=ADAM)
        out['adam_loss_running'] = np.array(m.loss_running)
        out['adam_bias'] = m.bias.detach().numpy()
        for i, b in enumerate(m.Bcp):
            out[f'adam_Bcp_{i}'] = b.detach().numpy()
        out['adam_predict'] = m.predict(X)
        # L-BFGS (fp64 cases only: fp32 strong-Wolfe branches are not stable)
        if dtype == torch.float64:
            m2 = STR.CP_linear_regression(X.shape, dtype=dtype, rank=R, non_negative=nn,
                                          weights=None if w is None else np.array(w),
                                          Bcp_init=[b.clone().requires_grad_(True) for b in init],
                                          bias_init=bias0, device='cpu')
            conv = m2.fit(X, y, lambda_L2=lam, max_iter=6, tol=1e-50, patience=10, verbose=False,
                          running_loss_logging_interval=1, LBFGS_kwargs=LBFGS)
            out['lbfgs_loss_running'] = np.array(m2.loss_running)
            out['lbfgs_bias'] = m2.bias.detach().numpy()
            out['lbfgs_converged'] = conv
            for i, b in enumerate(m2.Bcp):
                out[f'lbfgs_Bcp_{i}'] = b.detach().numpy()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print('wrote', name, 'loss', out['loss'])


def gen_mn(MTR):
    for name, N, dims, C, R, nn, w, cwmode, lam, seed in MN_CASES:
        X, y, _ = O.synth_mn(N, dims, R, C, 1234 + seed)
        init = O.init_mn(list(dims) + [C], R, nn, scale=1.0, seed=321)
        weights = torch.ones(R) if w is None else torch.tensor(w, dtype=torch.float32)
        counts = np.bincount(y.numpy(), minlength=C).astype(np.float64)
        cw = np.ones(C, dtype=np.float32) if cwmode == 'ones' else (N / (C * counts)).astype(np.float32)
        B = [b.clone().requires_grad_(True) for b in init]
        P = MTR.model(X, B, weights, nn)
        loss_fn = torch.nn.CrossEntropyLoss(weight=torch.as_tensor(cw, dtype=torch.float32))
        ce = loss_fn(P, y)
        loss = ce + lam * MTR.L2_penalty(B)
        loss.backward()
        out = {'X': X.numpy(), 'y': y.numpy(), 'weights': weights.numpy(), 'non_negative': np.array(nn),
               'class_weights': cw, 'lambda_L2': lam, 'R': R, 'C': C,
               'P': P.detach().numpy(), 'loss_data': ce.item(), 'loss': loss.item()}
        for i, b in enumerate(init):
            out[f'Bcp_init_{i}'] = b.numpy()
            out[f'grad_{i}'] = B[i].grad.numpy()
        m = MTR.CP_logistic_regression(X, y, rank=R, non_negative=nn,
                                       weights=None if w is None else w,
                                       Bcp_init=[b.clone().requires_grad_(True) for b in init], device='cpu')
        m.fit_Adam(lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, weights=cw, verbose=False, Adam_kwargs=ADAM)
        out['adam_loss_running'] = np.array(m.loss_running)
        for i, b in enumerate(m.Bcp):
            out[f'adam_Bcp_{i}'] = b.detach().numpy()
        prob, pred = m.predict()
        out['adam_prob'] = prob
        out['adam_pred'] = pred
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print('wrote', name, 'loss', out['loss'])


def gen_hier(HTR):
    """multinomial_tensor_regression_hierarchical.py: unweighted CE, three Adam parameter groups (3-D X)."""
    name, N, dims, C, R, lam, seed = 'hier_2mode', 96, (10, 8), 3, 3, 0.01, 21
    X, y, _ = O.synth_mn(N, dims, R, C, 1234 + seed)
    nn = [False, True, False]
    init = O.init_mn(list(dims) + [C], R, nn, scale=1.0, seed=321)
    out = {'X': X.numpy(), 'y': y.numpy(), 'non_negative': np.array(nn), 'lambda_L2': lam, 'R': R, 'C': C}
    for i, b in enumerate(init):
        out[f'Bcp_init_{i}'] = b.numpy()
    m = HTR.CP_logistic_regression(X, y, rank=R, non_negative=nn,
                                   Bcp_init=[b.clone().requires_grad_(True) for b in init], device='cpu')
    m.fit_Adam(lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, verbose=False, Adam_kwargs=ADAM)
    out['adam_loss_running'] = np.array(m.loss_running)
    for i, b in enumerate(m.Bcp):
        out[f'adam_Bcp_{i}'] = b.detach().numpy()
    prob, pred = m.predict()
    out['adam_prob'] = prob
    out['adam_pred'] = pred
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'final loss', out['adam_loss_running'][-1])


SPEC_CASES = [
    # name, T, W, D, n_out, rank_normal, rank_spectral, n_complex_dim, dtype, non_negative, weights, lambda, seed
    ('spec_f32', 48, 10, 12, 3, 2, 2, 1, torch.float32, [False, False, False], None, 0.01, 31),
    ('spec_f64_nn', 40, 9, 8, 2, 1, 3, 2, torch.float64, [True, False, False], [0.7, 1.0, 1.0, 1.0], 0.02, 32),
    ('spec_only_f64', 36, 7, 5, 4, 0, 2, 1, torch.float64, [False, False, False], None, 0.01, 33),
    ('spec_normal_only_f32', 40, 8, 16, 2, 3, 0, 0, torch.float32, [False, True, False], [1.0, 0.5, 2.0], 0.01, 34),
]


def gen_spec(SPR):
    """spectral_tensor_regression.py: one closure evaluation with the module-level functions, fit_Adam, fit (fp64),
    predict / predict_latents through the reference's estimator class."""
    from . import tr_oracle_spectral as OS
    for name, T, W, D, NO, rn, rs, ncd, dtype, nn, w, lam, seed in SPEC_CASES:
        cc = ncd + 1
        X, y = OS.synth(T, W, D, NO, max(rn, 1), max(rs, 1), cc, 1234 + seed, dtype=dtype)
        Bn0, Bc0 = OS.init(W, D, NO, rn, rs, cc, dtype=dtype, seed=321 + seed)
        weights = torch.ones(rn + rs, dtype=dtype) if w is None else torch.tensor(w, dtype=dtype)
        Bn = [b.clone().requires_grad_(True) for b in Bn0]
        Bc = [b.clone().requires_grad_(True) for b in Bc0]
        bias = torch.zeros(NO, dtype=dtype, requires_grad=True)
        y_hat = SPR.lin_model(X, Bn, weights[:rn], nn, bias) + SPR.stepwise_spectral_model(X, Bc, weights[rn:], nn, bias)
        mse = torch.nn.MSELoss()(y_hat, y)
        loss = mse + lam * (SPR.L2_penalty(Bn) + SPR.L2_penalty(Bc))
        loss.backward()
        out = {'X': X.numpy(), 'y': y.numpy(), 'weights': weights.numpy(), 'non_negative': np.array(nn), 'lambda_L2': lam,
               'rank_normal': rn, 'rank_spectral': rs, 'n_complex_dim': ncd,
               'y_hat': y_hat.detach().numpy(), 'loss_data': mse.item(), 'loss': loss.item(), 'dbias': bias.grad.numpy()}
        for i in range(3):
            out[f'Bn_init_{i}'] = Bn0[i].numpy()
            out[f'Bc_init_{i}'] = Bc0[i].numpy()
            out[f'grad_n_{i}'] = (Bn[i].grad if Bn[i].grad is not None else torch.zeros_like(Bn[i])).numpy()
            out[f'grad_c_{i}'] = (Bc[i].grad if Bc[i].grad is not None else torch.zeros_like(Bc[i])).numpy()

        def fresh():
            return SPR.CP_linear_regression(X.shape, y.shape, dtype=dtype, rank_normal=rn, rank_spectral=rs,
                                            non_negative=nn, weights=None if w is None else np.array(w),
                                            Bcp_init=[[b.clone().requires_grad_(True) for b in Bn0],
                                                      [b.clone().requires_grad_(True) for b in Bc0]],
                                            n_complex_dim=ncd, device='cpu')
        m = fresh()
        m.fit_Adam(X, y, lambda_L2=lam, max_iter=20, tol=1e-50, patience=100, verbose=False, Adam_kwargs=ADAM)
        out['adam_loss_running'] = np.array(m.loss_running)
        out['adam_bias'] = m.bias.detach().numpy()
        for i in range(3):
            out[f'adam_Bn_{i}'] = m.Bcp_n[i].detach().numpy()
            out[f'adam_Bc_{i}'] = m.Bcp_c[i].detach().numpy()
        # the reference's predict adds spectral_model's (T, rank_spectral) output to the (T, n_out) normal part:
        # only defined when torch can broadcast the two (stored when it can)
        try:
            out['adam_predict'] = m.predict(X).numpy()
        except RuntimeError:
            pass
        if rn > 0:
            out['adam_latents'] = np.asarray(m.predict_latents(X))
        if dtype == torch.float64:
            m2 = fresh()
            conv = m2.fit(X, y, lambda_L2=lam, max_iter=6, tol=1e-50, patience=10, verbose=False,
                          running_loss_logging_interval=1, LBFGS_kwargs=LBFGS)
            out['lbfgs_loss_running'] = np.array(m2.loss_running)
            out['lbfgs_bias'] = m2.bias.detach().numpy()
            out['lbfgs_converged'] = conv
            for i in range(3):
                out[f'lbfgs_Bn_{i}'] = m2.Bcp_n[i].detach().numpy()
                out[f'lbfgs_Bc_{i}'] = m2.Bcp_c[i].detach().numpy()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print('wrote', name, 'loss', out['loss'], 'adam', out['adam_loss_running'][[0, -1]])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'spec':        # only the spectral fixtures (the others stay as committed)
        if not ref_loader.available():
            sys.exit('reference not found at ' + ref_loader.REFERENCE_DIR)
        torch.set_num_threads(1)
        gen_spec(ref_loader.spectral())
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'hier':        # only the hierarchical fixture (the others stay as committed)
        if not ref_loader.available():
            sys.exit('reference not found at ' + ref_loader.REFERENCE_DIR)
        torch.set_num_threads(1)
        gen_hier(ref_loader.hierarchical())
        return
    if not ref_loader.available():
        sys.exit('reference not found at ' + ref_loader.REFERENCE_DIR)
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)            # fixed summation order for the fixtures
    gen_std(ref_loader.standard())
    gen_mn(ref_loader.multinomial())
    gen_hier(ref_loader.hierarchical())
    gen_spec(ref_loader.spectral())


if __name__ == '__main__':
    main()
