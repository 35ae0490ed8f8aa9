#!/usr/bin/env python
"""Spectral fit iteration (tr_spec_fwd_grad) over the channel count Q = rank_normal + rank_spectral * complex columns and the
dtype: two-pass path against the single-pass kernel on 60 000 samples of 64 x 128 (fp32) / 64 x 64 (fp64) — DESIGN 4e.
    python tools/spec_channels.py"""
import sys, torch
sys.path.insert(0, '.')
from tensor_regression_b200 import engine
dev='cuda:0'
def run(N, W, D, NO, rn, rs, cc, dtype, single, reps=10):
    X = torch.randn((N, W, D), device=dev, dtype=dtype); y = torch.randn((N, NO), device=dev, dtype=dtype)
    eng = engine.SpectralEngine(W, D, NO, rn, rs, cc, dtype, dev)
    th = (0.2 * torch.rand(eng.P, device=dev) - 0.1).to(dtype); w = torch.ones(rn + rs, device=dev, dtype=dtype)
    eng.set_option('spec_single', single)
    for _ in range(3): eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (rn, rs, cc) in ((1,0,1),(2,0,1),(2,1,2),(4,0,1),(1,2,2),(2,2,2),(3,2,2),(4,2,2),(8,0,1),(0,2,4)):
    for dtype, D in ((torch.float32,128),(torch.float64,64)):
        a = run(60000, 64, D, 4, rn, rs, cc, dtype, 0); b = run(60000, 64, D, 4, rn, rs, cc, dtype, 1)
        print(f'Q={rn+rs*cc} (rn={rn}, rs={rs}, cc={cc}) {str(dtype)[6:]}: two-pass {a:.3f} ms, single-pass {b:.3f} ms, ratio {a/b:.2f}')
