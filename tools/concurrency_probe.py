#!/usr/bin/env python
"""Does the single-pass cluster kernel (16-CTA clusters, 112 of 148 SMs) run concurrently with other kernels?
Times a single-pass fwd_grad (engine A, stream A), a two-pass fwd_grad on a smaller sample set (engine B, stream B),
and both queued together.  usage: python tools/concurrency_probe.py [mn|std]"""
import sys
import torch
sys.path.insert(0, '.')
from tensor_regression_b200 import engine  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else 'mn'
dev = 'cuda:0'
if kind == 'mn':
    dims, R, C, NA, NB = (100, 50, 20), 6, 10, 30000, 4000
else:
    dims, R, C, NA, NB = (64, 64, 32), 8, 0, 20000, 3000
XA = torch.randn((NA, *dims), device=dev)
XB = torch.randn((NB, *dims), device=dev)
A = engine.Engine(dims, R, C, torch.float32, dev)
B = engine.Engine(dims, R, C, torch.float32, dev)
th = 0.2 * torch.rand(A.P, device=dev) - 0.1
w = torch.ones(R, device=dev)
A.set_option('fused', 1); A.set_option('hybrid', 0)
B.set_option('fused', 0)
if C:
    _, yA = A.forward_mn(XA, th, w, 0, 50.0, 1.0)
    _, yB = B.forward_mn(XB, th, w, 0, 50.0, 1.0)
    cw = torch.ones(C, device=dev)
    fa = lambda: A.fwd_grad_mn(XA, yA, cw, th, w, 0, 50.0, 1.0)      # noqa: E731
    fb = lambda: B.fwd_grad_mn(XB, yB, cw, th, w, 0, 50.0, 1.0)      # noqa: E731
else:
    yA = torch.randn(NA, device=dev); yB = torch.randn(NB, device=dev)
    fa = lambda: A.fwd_grad_std(XA, yA, th, w, 0, 50.0, 1.0)         # noqa: E731
    fb = lambda: B.fwd_grad_std(XB, yB, th, w, 0, 50.0, 1.0)         # noqa: E731
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def only_a():
    with torch.cuda.stream(sA):
        fa()
    torch.cuda.current_stream().wait_stream(sA)


def only_b():
    with torch.cuda.stream(sB):
        fb()
    torch.cuda.current_stream().wait_stream(sB)


def both():
    sA.wait_stream(torch.cuda.current_stream()); sB.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(sA):
        fa()
    with torch.cuda.stream(sB):
        fb()
    torch.cuda.current_stream().wait_stream(sA); torch.cuda.current_stream().wait_stream(sB)


for f in (only_a, only_b, both):
    f()
print(kind, 'A (single-pass, %d samples): %.3f ms | B (two-pass, %d samples): %.3f ms | both queued together: %.3f ms'
      % (NA, timed(only_a), NB, timed(only_b), timed(both)), A.launch_info()['path'], '|', B.launch_info()['path'])
