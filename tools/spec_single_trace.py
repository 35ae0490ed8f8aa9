#!/usr/bin/env python
"""Per-sample timeline of the single-pass spectral kernel (block 0), from a library built with -DTRSS_TRACE:

    make -C tensor_regression_b200/csrc OUT=../libtrb200_trace.so OBJDIR=build_trace TR_NVCC_EXTRA=-DTRSS_TRACE -j8
    TR_B200_LIB=tensor_regression_b200/libtrb200_trace.so python tools/spec_single_trace.py [N]

Events per sample j of the block (cycles relative to the first traced TMA issue):
  0 TMA issued | 1 forward warp starts waiting | 2 bytes landed (forward start) | 3 window loop done |
  4 epilogue done, da released (9 norms done, 10 s reduced, 11 ds reduced) | 5 gradient warp 0 starts waiting | 6 gradient start | 7 gradient warp 0 done | 8 gradient warp 7 done
"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tensor_regression_b200 import _lib, engine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
W, D, NO, rn, rs, cc = 64, 128, 4, 2, 2, 2
dev = 'cuda:0'
X = torch.randn((N, W, D), device=dev)
y = torch.randn((N, NO), device=dev)
eng = engine.SpectralEngine(W, D, NO, rn, rs, cc, torch.float32, dev)
th = 0.2 * torch.rand(eng.P, device=dev) - 0.1
w = torch.ones(rn + rs, device=dev)
eng.set_option('spec_single', 1)
for _ in range(3):
    eng.fwd_grad(X, y, th, w, 0, 50.0, 1.0)
torch.cuda.synchronize()
print(eng.launch_info())
NSM, EV = 48, 16
out = (ctypes.c_longlong * (NSM * EV))()
_lib.lib.tr_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
rc = _lib.lib.tr_debug_trace(eng._h, out, NSM * EV)
assert rc == 0, rc
t = np.array(out, dtype=np.int64).reshape(NSM, EV)
t0 = t[0, 0]
names = ['tma', 'fwait', 'fwd0', 'loop1', 'ready', 'gwait', 'g0', 'g1', 'g1w7', 'norm', 'sred', 'dsred']
print('sample ' + ' '.join(f'{n:>8s}' for n in names))
for i in range(NSM):
    print(f'{i:6d} ' + ' '.join(f'{(t[i, e] - t0) if t[i, e] else 0:8d}' for e in range(len(names))))
d = lambda a, b: np.median((t[:, b] - t[:, a])[(t[:, a] > 0) & (t[:, b] > 0)])  # noqa: E731
print('median cycles: tma->landed(fwd0) %d | fwd wait %d | window loop %d | epilogue %d | ready->g0 %d | gradient w0 %d | tma->g1 (stage residency) %d'
      % (d(0, 2), d(1, 2), d(2, 3), d(3, 4), d(4, 6), d(6, 7), d(0, 8)))
print('epilogue: store a + norms %d | second contraction + reduce %d | outputs + ds reduce %d | da + release %d' % (d(3, 9), d(9, 10), d(10, 11), d(11, 4)))
print('period (cycles per sample): %.0f' % np.median(np.diff(t[:, 0])))
