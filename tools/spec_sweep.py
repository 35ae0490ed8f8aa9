#!/usr/bin/env python
"""Spectral fit iteration (tr_spec_fwd_grad): single-pass kernel against the two-pass path over the number of samples
(where does the automatic selection have to switch?) and for fp64.   python tools/spec_sweep.py"""
import sys
import time

import torch

sys.path.insert(0, '.')
from tensor_regression_b200 import engine  # noqa: E402

dev = 'cuda:0'


def run(N, W, D, NO, rn, rs, cc, dtype, single, reps=20):
    X = torch.randn((N, W, D), device=dev, dtype=dtype)
    y = torch.randn((N, NO), device=dev, dtype=dtype)
    eng = engine.SpectralEngine(W, D, NO, rn, rs, cc, dtype, dev)
    th = (0.2 * torch.rand(eng.P, device=dev) - 0.1).to(dtype)
    w = torch.ones(rn + rs, device=dev, dtype=dtype)
    eng.set_option('spec_single', single)
    for _ in range(3):
        eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, eng.launch_info()['path'][:11]


import os
for ns in ([int(v) for v in os.environ['SWEEP_NS'].split(',')] if os.environ.get('SWEEP_NS') else []):
    # stage-count sweep on a 16 KB sample (W = 32) and on the 32 KB sample of spec1: is the ring the limit?
    for W, D in ((32, 128), (64, 128)):
        Xs = torch.randn((100000, W, D), device=dev)
        ys = torch.randn((100000, 4), device=dev)
        eng = engine.SpectralEngine(W, D, 4, 2, 2, 2, torch.float32, dev)
        th = 0.2 * torch.rand(eng.P, device=dev) - 0.1
        w = torch.ones(4, device=dev)
        eng.set_option('spec_single', 1)
        eng.set_option('spec_single_ns', ns)
        for _ in range(3):
            eng.fwd_grad(Xs, ys, th, w, 0, 50.0, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.fwd_grad(Xs, ys, th, w, 0, 50.0, 1.0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'stages<={ns:2d} W={W} D={D} N=100000: {ms:7.4f} ms, {Xs.numel() * 4 / ms / 1e6:7.1f} GB/s of X, stages {eng.launch_info()["stages"]}')
        del Xs, ys
if os.environ.get('SWEEP_NS'):
    sys.exit(0)
for dtype, W, D in ((torch.float32, 64, 128), (torch.float64, 64, 64)):
    for N in (300, 1200, 2400, 4800, 9600, 40000, 160000 if dtype == torch.float32 else 80000):
        a, pa = run(N, W, D, 4, 2, 2, 2, dtype, 0)
        b, pb = run(N, W, D, 4, 2, 2, 2, dtype, 1)
        print(f'{str(dtype)[6:]} N={N:7d} W={W} D={D}: two-pass {a:8.4f} ms ({pa}) | single-pass {b:8.4f} ms ({pb}) | ratio {a / b:5.2f}')
