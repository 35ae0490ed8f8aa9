"""Per-iteration wall time of fit_Adam / fit on a problem so small that the GPU work is negligible: what the
host-side loop (Python wrappers, ctypes calls, launches, the per-iteration loss read) costs."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from tensor_regression_b200 import standard_tensor_regression as STR
from tensor_regression_b200 import multinomial_tensor_regression as MTR

dev = 'cuda:0'
g = torch.Generator().manual_seed(1)
X = torch.randn((512, 8, 6, 16), generator=g).to(dev)
y = torch.randn(512, generator=g).to(dev)
yc = torch.randint(0, 4, (512,), generator=g)
adam = {'lr': 0.01, 'amsgrad': True}
for name in ('std', 'mn'):
    for it in (50, 400):
        if name == 'std':
            m = STR.CP_linear_regression(X.shape, rank=3, device=dev)
            run = lambda k: m.fit_Adam(X, y, max_iter=k, tol=0.0, patience=10 ** 9, Adam_kwargs=adam)
        else:
            m = MTR.CP_logistic_regression(X, yc, rank=3, device=dev)
            run = lambda k: m.fit_Adam(max_iter=k, tol=0.0, patience=10 ** 9, weights=np.ones(4, dtype=np.float32), Adam_kwargs=adam)
        run(5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(it)
        torch.cuda.synchronize()
        print(f'{name} fit_Adam: {(time.perf_counter() - t0) / it * 1e6:8.1f} us / iteration ({it} iterations)')


# L-BFGS (fit): wall time per closure evaluation on the same tiny problem
from tensor_regression_b200 import lbfgs as L
kw = {'lr': 1, 'max_iter': 20, 'max_eval': None, 'tolerance_grad': 1e-07, 'tolerance_change': 1e-09,
      'history_size': 100, 'line_search_fn': 'strong_wolfe'}
evals = [0]
orig = L.LBFGS._evaluate


def counting(self, closure, with_d):
    evals[0] += 1
    return orig(self, closure, with_d)


L.LBFGS._evaluate = counting
m = STR.CP_linear_regression(X.shape, rank=3, device=dev)
m.fit(X, y, max_iter=2, tol=0.0, patience=10 ** 9, running_loss_logging_interval=1, LBFGS_kwargs=kw)
torch.cuda.synchronize()
evals[0] = 0
t0 = time.perf_counter()
m.fit(X, y, max_iter=20, tol=0.0, patience=10 ** 9, running_loss_logging_interval=1, LBFGS_kwargs=kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'std fit (L-BFGS): {dt / max(1, evals[0]) * 1e6:8.1f} us / closure evaluation ({evals[0]} evaluations, 20 outer iterations)')
