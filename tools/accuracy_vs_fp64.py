"""error of the fp32 CUDA path against the float64 truth, next to the error of the reference's own
fp32 arithmetic (oracle port: dense B + one matmul + autograd) against the same truth."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import tr_oracle as O
from tensor_regression_b200 import engine


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


dev = 'cuda:0'
torch.set_num_threads(8)
for name, N, dims, R in [('cfg2 shape', 1024, (64, 64, 32), 8), ('cfg1', 2000, (20, 30, 40), 5)]:
    X, y, _ = O.synth_std(N, dims, R, 1236)
    y = y.reshape(-1)
    nn = [False] * (len(dims) + 1)
    B0 = O.init_std(dims, R, nn)
    bias, w = torch.tensor([0.05]), torch.ones(R)
    t64 = O.std_loss_grad(X.double(), y.double(), [b.double() for b in B0], bias.double(), w.double(), nn, 0.01)
    r32 = O.std_loss_grad(X, y, B0, bias, w, nn, 0.01)
    eng = engine.Engine(dims, R, 0, torch.float32, dev)
    theta = O.pack(B0, bias).to(dev)
    yh = torch.empty(N, device=dev)
    for fused in (0, 1):
        eng.set_option('fused', fused)
        gs = eng.fwd_grad_std(X.to(dev), y.to(dev), theta, w.to(dev), 0, 50.0, 1.0, yhat=yh)
        grad, loss = eng.finish(gs, 2.0 / N, 1.0 / N, theta, 0.01, 0, 50.0, 1.0)
        want = torch.cat([g.reshape(-1) for g in t64['grads']] + [t64['dbias'].reshape(-1)])
        ref = torch.cat([g.reshape(-1) for g in r32['grads']] + [r32['dbias'].reshape(-1)])
        print(f"{name:11s} std  {'single-pass' if fused else 'two-pass   '}: y_hat ours {rel(yh.cpu(), t64['y_hat']):.2e} reference-fp32 {rel(r32['y_hat'], t64['y_hat']):.2e} | "
              f"grad ours {rel(grad.cpu(), want):.2e} reference-fp32 {rel(ref, want):.2e} | loss ours {abs(loss[1].item() - t64['loss'].item()) / t64['loss'].item():.2e} "
              f"reference-fp32 {abs(r32['loss'].item() - t64['loss'].item()) / t64['loss'].item():.2e}")
for name, N, dims, C, R in [('cfg3 shape', 1024, (100, 50, 20), 10, 6)]:
    X, y, _ = O.synth_mn(N, dims, R, C, 1237)
    nn = [False] * (len(dims) + 1)
    B0 = O.init_mn(list(dims) + [C], R, nn, scale=0.2)
    w, cw = torch.ones(R), np.ones(C, dtype=np.float32)
    t64 = O.mn_loss_grad(X.double(), y, [b.double() for b in B0], w.double(), nn, cw.astype(np.float64), 0.01)
    r32 = O.mn_loss_grad(X, y, B0, w, nn, cw, 0.01)
    eng = engine.Engine(dims, R, C, torch.float32, dev)
    theta = O.pack(B0).to(dev)
    P = torch.empty((N, C), device=dev)
    gs = eng.fwd_grad_mn(X.to(dev), y.to(dev), torch.ones(C, device=dev), theta, w.to(dev), 0, 50.0, 1.0, P=P)
    grad, loss = eng.finish(gs, 1.0 / N, 1.0 / N, theta, 0.01, 0, 50.0, 1.0)
    want = torch.cat([g.reshape(-1) for g in t64['grads']])
    ref = torch.cat([g.reshape(-1) for g in r32['grads']])
    print(f"{name:11s} mn   two-pass   : P ours {rel(P.cpu(), t64['P']):.2e} reference-fp32 {rel(r32['P'], t64['P']):.2e} | "
          f"grad ours {rel(grad.cpu(), want):.2e} reference-fp32 {rel(ref, want):.2e} | loss ours {abs(loss[1].item() - t64['loss'].item()) / t64['loss'].item():.2e} "
          f"reference-fp32 {abs(r32['loss'].item() - t64['loss'].item()) / t64['loss'].item():.2e}")
