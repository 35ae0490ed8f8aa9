"""Cost of one L-BFGS closure evaluation through CP_linear_regression.fit (the reference's default entry point,
std:305-398) next to the bare forward+gradient kernels: what the device-resident optimiser adds (direction /
trial-point kernels, 4 doubles read back per evaluation)."""
import json, sys, time
import torch
sys.path.insert(0, '.')
from tensor_regression_b200 import standard_tensor_regression as STR
from tensor_regression_b200 import lbfgs as L

dev = 'cuda:0'
N, dims, R = 40000, (64, 64, 32), 8
g = torch.Generator(device=dev).manual_seed(5)
X = torch.empty((N, *dims), device=dev)
for lo in range(0, N, 4096):
    X[lo:lo + 4096].normal_(generator=g)
y = torch.randn(N, device=dev, generator=g)
kw = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
      'history_size': 100, 'line_search_fn': 'strong_wolfe'}
evals = [0]
orig = L.LBFGS._evaluate


def counting(self, closure, with_d):
    evals[0] += 1
    return orig(self, closure, with_d)


L.LBFGS._evaluate = counting
m = STR.CP_linear_regression(X.shape, rank=R, device=dev)
m.fit(X, y, max_iter=1, tol=1e-50, running_loss_logging_interval=1, LBFGS_kwargs=kw)      # warm-up
torch.cuda.synchronize()
evals[0] = 0
t0 = time.perf_counter()
outer = 4
m.fit(X, y, max_iter=outer, tol=1e-50, running_loss_logging_interval=1, LBFGS_kwargs=kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
eng = m._engine()
gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=dev)
w = torch.ones(R, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    eng.fwd_grad_std(X, y, m.theta, w, 0, 50.0, 1.0, gradsum=gs)
e0.record()
for _ in range(10):
    eng.fwd_grad_std(X, y, m.theta, w, 0, 50.0, 1.0, gradsum=gs)
e1.record()
torch.cuda.synchronize()
bare = e0.elapsed_time(e1) / 10
# per outer iteration fit also runs one logging forward (std:380-382): count it as work
print(json.dumps({'workload': f'standard, X ({N}, 64,64,32) fp32 rank 8 ({X.numel() * 4 / 1e9:.1f} GB)', 'outer_iterations': outer,
                  'closure_evaluations': evals[0], 'logging_forwards': outer, 'fit_seconds': dt,
                  'ms_per_closure_incl_everything': dt * 1e3 / evals[0], 'ms_bare_fwd_grad_kernels': bare,
                  'final_losses': m.loss_running[-outer:]}))
