#!/usr/bin/env python
"""Per-sample timeline of the single-pass multinomial kernel (cluster 0, CTA rank 0), from a library built with
-DTRM_TRACE:

    make -C tensor_regression_b200/csrc OUT=../libtrb200_trace.so OBJDIR=build_trace TR_NVCC_EXTRA=-DTRM_TRACE -j8
    TR_B200_LIB=tensor_regression_b200/libtrb200_trace.so python tools/fused_mn_trace.py [cfg3|cfg5] [N]

Events per sample i (cycles, relative to the TMA issue of the first traced sample):
  0 TMA issued | 1 forward start (data landed) | 2 forward end | 3 reducer saw the 4 warp partials |
  4 epilogue warp ready (owned samples: i % CL == 0) | 5 all CL partials arrived | 6 u summed | 7 epilogue done, v sent |
  8/9 gradient A start / end | 10/11 gradient B start / end
"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tensor_regression_b200 import _lib, engine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # 1: skip gradient work, 2: skip forward work, 4: skip epilogue math
dims, R, C = ((100, 50, 20), 6, 10) if wl == 'cfg3' else ((100, 50, 20), 4, 4)
dev = 'cuda:0'
X = torch.randn((N, *dims), device=dev)
eng = engine.Engine(dims, R, C, torch.float32, dev)
th = 0.2 * torch.rand(eng.P, device=dev) - 0.1
w = torch.ones(R, device=dev)
_, y = eng.forward_mn(X, th, w, 0, 50.0, 1.0)
cw = torch.ones(C, device=dev)
eng.set_option('fused', 1)
eng.set_option('flow_debug', dbg)
for _ in range(3):
    eng.fwd_grad_mn(X, y, cw, th, w, 0, 50.0, 1.0)
torch.cuda.synchronize()
print('dbg', dbg, eng.launch_info())
NS, EV = 48, 16
out = (ctypes.c_longlong * (NS * EV))()
_lib.lib.tr_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
rc = _lib.lib.tr_debug_trace(eng._h, out, NS * EV)
assert rc == 0, rc
t = np.array(out, dtype=np.int64).reshape(NS, EV)
t0 = t[0, 0]
names = ['tma', 'fwd0', 'fwd1', 'red', 'epi_rdy', 'parts', 'usum', 'epi_end', 'gA0', 'gA1', 'gB0', 'gB1']
print('sample ' + ' '.join(f'{n:>8s}' for n in names))
for i in range(NS):
    print(f'{i:6d} ' + ' '.join(f'{(t[i, e] - t0) if t[i, e] else 0:8d}' for e in range(12)))
d = lambda a, b: np.median((t[:, b] - t[:, a])[(t[:, a] > 0) & (t[:, b] > 0)])  # noqa: E731
print('median cycles: tma->fwd0 %d | fwd %d | fwd1->red %d | red->gA0 %d | gradA %d | gradB %d | tma->gB1 (stage residency) %d'
      % (d(0, 1), d(1, 2), d(2, 3), d(3, 8), d(8, 9), d(10, 11), d(0, 11)))
print('TMA issue -> bytes landed (observer warp) %d | landed -> forward start %d' % (d(0, 14), d(14, 1)))
print('forward: compute %d | warp reduce + pA %d' % (d(1, 13), d(13, 2)))
own = t[:, 5] > 0
if own.any():
    print('owned samples: epi_rdy->parts %d | parts->usum %d | epilogue math %d | red(own)->parts %d'
          % (np.median((t[own, 5] - t[own, 4])), np.median(t[own, 6] - t[own, 5]), np.median(t[own, 7] - t[own, 6]),
             np.median(t[own, 5] - t[own, 3])))
print('period (cycles per sample): %.0f' % np.median(np.diff(t[:, 0])))
