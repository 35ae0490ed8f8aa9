import sys, torch, numpy as np
sys.path.insert(0, '.')
from oracle import tr_oracle as O
from tensor_regression_b200 import engine
N, dims, R, dt = 300, (64, 64, 32), 8, torch.float32
X, y, _ = O.synth_std(N, dims, R, 1241, dtype=dt)
y = y.reshape(-1)
nn = [True] + [False] * len(dims)
B0 = O.init_std(dims, R, nn, dtype=dt)
bias = torch.tensor([0.07], dtype=dt)
w = torch.linspace(0.5, 1.5, R, dtype=dt)
eng = engine.Engine(dims, R, 0, dt, 'cuda:0')
theta = O.pack(B0, bias).cuda()
eng.set_option('fused', 0); eng.set_option('flow', 1)
yh = torch.empty(N, device='cuda')
gs = eng.fwd_grad_std(X.cuda(), y.cuda(), theta, w.cuda(), 1, 50.0, 1.0, yhat=yh)
torch.cuda.synchronize()
print(eng.launch_info(), gs[:4])
