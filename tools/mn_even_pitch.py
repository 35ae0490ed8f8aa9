#!/usr/bin/env python
"""Single-pass multinomial kernel on last modes with an EVEN number of 16-byte chunks (bank conflicts on the row-per-thread
shared-memory reads: 2-way for 2 / 6 chunks, 4-way for 4, 8-way for 8) against the two-pass kernels.
    python tools/mn_even_pitch.py"""
import sys

import torch

sys.path.insert(0, '.')
from tensor_regression_b200 import engine  # noqa: E402

dev = 'cuda:0'


def run(N, dims, R, C, fused, reps=5):
    X = torch.randn((N, *dims), device=dev)
    eng = engine.Engine(dims, R, C, torch.float32, dev)
    th = 0.2 * torch.rand(eng.P, device=dev) - 0.1
    w = torch.ones(R, device=dev)
    y = torch.randint(0, C, (N,), device=dev)
    cw = torch.ones(C, device=dev)
    eng.set_option('fused', fused)
    try:
        for _ in range(2):
            eng.fwd_grad_mn(X, y, cw, th, w, 0, 50.0, 1.0)
    except engine.TRError as e:
        return None, str(e)[:60]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.fwd_grad_mn(X, y, cw, th, w, 0, 50.0, 1.0)
    e1.record()
    torch.cuda.synchronize()
    info = eng.launch_info()
    return e0.elapsed_time(e1) / reps, f"{info['path'][:11]} CL {info.get('cluster_size')} NS {info.get('stages')}"


for dims in ((100, 50, 20), (100, 50, 24), (100, 60, 8), (100, 50, 16), (100, 40, 32), (64, 64, 32), (100, 50, 12), (100, 50, 28)):
    D = dims[0] * dims[1] * dims[2]
    N = int(16e9 / (4 * D))
    a, ia = run(N, dims, 6, 10, 0)
    b, ib = run(N, dims, 6, 10, 1)
    c, ic = run(N, dims, 6, 10, -1)
    gb = N * D * 4 / 1e9
    print(f'dims {dims} chunks {dims[2] // 4} N {N} ({gb:.1f} GB): two-pass {a:.3f} ms | forced single-pass '
          f'{(f"{b:.3f} ms" if b else "n/a")} ({ib}) | auto {c:.3f} ms ({ic})')
