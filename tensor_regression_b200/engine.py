"""Host-side driver of the C-ABI kernels: one ``Engine`` per model geometry.

PyTorch is used only for device memory, streams and (optionally) ``torch.distributed``; every
number on the fit path comes from libtrb200.so.  ``Engine`` has no CPU mode.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import TRError

_DT = {torch.float32: _lib.TR_F32, torch.float64: _lib.TR_F64}


def nn_mask_of(non_negative, n_factors):
    """Bit m set <=> softplus on list position m (non_neg_fn, std:81-85 / mn:142-146).
    Entries beyond the factor list (std's unused trailing entry, std:281-284) are ignored."""
    mask = 0
    for m in range(min(len(non_negative), n_factors)):
        if bool(non_negative[m]):
            mask |= 1 << m
    return mask


def factor_offsets(dims, rank, n_classes):
    sizes = [int(d) * rank for d in dims] + ([n_classes * rank] if n_classes > 0 else [])
    return sizes, np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_bounds(n_total, rank, world):
    """Contiguous, balanced split of the sample axis: rank r owns [lo, hi)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Engine:
    """Plan + workspace for one (dims, rank, n_classes, dtype, device)."""

    def __init__(self, dims, rank, n_classes=0, dtype=torch.float32, device='cuda'):
        device = torch.device(device)
        if device.type != 'cuda':
            raise TRError(f"device '{device}': tensor_regression_b200 runs on CUDA devices only (no CPU path)")
        if not torch.cuda.is_available():
            raise TRError('no CUDA device available (tensor_regression_b200 has no CPU path)')
        if dtype not in _DT:
            raise TRError(f'dtype {dtype} not supported (float32 / float64)')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device, self.dtype = device, dtype
        self.dims = [int(d) for d in dims]
        self.rank, self.n_classes = int(rank), int(n_classes)
        self.D = int(np.prod(self.dims))
        h = ctypes.c_void_p()
        arr = (ctypes.c_int64 * len(self.dims))(*self.dims)
        rc = _lib.lib.tr_create(ctypes.byref(h), _DT[dtype], len(self.dims), arr, self.rank, self.n_classes,
                                device.index)
        if rc != 0:
            raise TRError(f'tr_create failed ({rc}): {_lib.lib.tr_last_error(None).decode()}')
        self._h = h
        P, Pf, ngs = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._ck(_lib.lib.tr_param_count(h, ctypes.byref(P), ctypes.byref(Pf)))
        self._ck(_lib.lib.tr_gradsum_count(h, ctypes.byref(ngs)))
        self.P, self.Pf, self.n_gradsum = P.value, Pf.value, ngs.value
        self.launches = 0            # kernels launched through this engine (bench: gpu_launches)

    # -- plumbing -----------------------------------------------------------------------------
    # (no torch.cuda.device(...) context around the calls: every entry point of the library selects its
    # handle's device itself, and the context manager costs more host time than a whole small-problem launch)
    def _ck(self, rc):
        if rc != 0:
            raise TRError(f'libtrb200 error {rc}: {_lib.lib.tr_last_error(self._h).decode()}')

    def close(self):
        if getattr(self, '_h', None):
            _lib.lib.tr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _x(self, X):
        if not isinstance(X, torch.Tensor) or X.device != self.device or X.dtype != self.dtype:
            raise TRError(f'X must be a {self.dtype} tensor on {self.device}')
        if list(X.shape[1:]) != self.dims:
            raise TRError(f'X.shape[1:]={list(X.shape[1:])} does not match the model dims {self.dims}')
        return X if X.is_contiguous() else X.contiguous()

    def _vec(self, t, n, name, dtype=None):
        dtype = dtype or self.dtype
        if not isinstance(t, torch.Tensor) or t.device != self.device or t.dtype != dtype or t.numel() != n \
                or not t.is_contiguous():
            raise TRError(f'{name} must be a contiguous {dtype} tensor with {n} elements on {self.device}')
        return t

    def _count(self):
        info = (ctypes.c_int64 * 8)()
        _lib.lib.tr_last_launch_info(self._h, info)
        self.launches += int(info[0])
        return list(info)

    def reserve(self, n_samples):
        self._ck(_lib.lib.tr_reserve(self._h, int(n_samples)))

    def launch_info(self):
        info = (ctypes.c_int64 * 8)()
        _lib.lib.tr_last_launch_info(self._h, info)
        if int(info[7]) < 0 and int(info[2]) == 0:
            keys = ['launches', 'grid', '_', 'lag_samples', 'groups', 'chunks', 'channels', 'vector_width']
            d = {k: int(v) for k, v in zip(keys, info) if k != '_'}
            d['vector_width'] = -d['vector_width']
            d['path'] = 'single-launch dataflow (second read of X from L2)'
            return d
        if int(info[7]) < 0:
            keys = ['launches', 'grid', 'cluster_size', 'stages', 'clusters', 'chunks', 'channels', 'vector_width']
            d = dict(zip(keys, [int(v) for v in info]))
            d['vector_width'] = -d['vector_width']
            d['path'] = 'single-pass (cluster-resident sample)'
            return d
        keys = ['launches', 'grid_fwd', 'grid_grad', 'tiles_per_sample', 'groups_fwd', 'groups_grad', 'channels',
                'vector_width']
        d = dict(zip(keys, [int(v) for v in info]))
        d['path'] = 'two-pass'
        return d

    def profile(self, enable=True):
        """Start (and reset) / stop CUDA-event timing of the two streaming kernels."""
        self._ck(_lib.lib.tr_profile_enable(self._h, 1 if enable else 0))

    def profile_read(self):
        out = (ctypes.c_double * 6)()
        self._ck(_lib.lib.tr_profile_read(self._h, out))
        return {'fwd_ms': out[0], 'fwd_launches': int(out[1]), 'grad_ms': out[2], 'grad_launches': int(out[3]),
                'fused_ms': out[4], 'fused_launches': int(out[5])}

    def set_option(self, name, value):
        """'fused': -1 auto, 0 never, 1 always the single-pass cluster kernel (standard model);
        'flow': -1 auto, 0 never, 1 always the single-launch dataflow kernel; 'flow_window_mb': bytes of X
        the forward warps may lead the gradient warps by (the part of X that must stay in L2)."""
        self._ck(_lib.lib.tr_set_option(self._h, name.encode(), int(value)))

    # -- the hot path --------------------------------------------------------------------------
    def forward_std(self, X, theta, w, nn_mask, beta, thr):
        X = self._x(X)
        N = X.shape[0]
        yhat = torch.empty(N, dtype=self.dtype, device=self.device)
        self._ck(_lib.lib.tr_forward_std(self._h, X.data_ptr(), N, self._vec(theta, self.P, 'theta').data_ptr(),
                                         self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                         yhat.data_ptr(), self._stream()))
        self._count()
        return yhat

    def forward_mn(self, X, theta, w, nn_mask, beta, thr, want_pred=True):
        X = self._x(X)
        N = X.shape[0]
        P = torch.empty((N, self.n_classes), dtype=self.dtype, device=self.device)
        pred = torch.empty(N, dtype=torch.int64, device=self.device) if want_pred else None
        self._ck(_lib.lib.tr_forward_mn(self._h, X.data_ptr(), N, self._vec(theta, self.P, 'theta').data_ptr(),
                                        self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                        P.data_ptr(), pred.data_ptr() if want_pred else None, self._stream()))
        self._count()
        return P, pred

    def fwd_grad_std(self, X, y, theta, w, nn_mask, beta, thr, gradsum=None, yhat=None):
        X = self._x(X)
        N = X.shape[0]
        if gradsum is None:
            gradsum = torch.empty(self.n_gradsum, dtype=torch.float64, device=self.device)
        self._ck(_lib.lib.tr_fwd_grad_std(self._h, X.data_ptr(), self._vec(y, N, 'y').data_ptr(), N,
                                          self._vec(theta, self.P, 'theta').data_ptr(),
                                          self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                          self._vec(gradsum, self.n_gradsum, 'gradsum', torch.float64).data_ptr(),
                                          yhat.data_ptr() if yhat is not None else None, self._stream()))
        self._count()
        return gradsum

    def backward_std(self, X, dyhat, theta, w, nn_mask, beta, thr, gradsum=None):
        X = self._x(X)
        N = X.shape[0]
        if gradsum is None:
            gradsum = torch.empty(self.n_gradsum, dtype=torch.float64, device=self.device)
        self._ck(_lib.lib.tr_backward_std(self._h, X.data_ptr(), self._vec(dyhat, N, 'dyhat').data_ptr(), N,
                                          self._vec(theta, self.P, 'theta').data_ptr(),
                                          self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                          gradsum.data_ptr(), self._stream()))
        self._count()
        return gradsum

    def backward_mn(self, X, dP, theta, w, nn_mask, beta, thr, gradsum=None):
        X = self._x(X)
        N = X.shape[0]
        if gradsum is None:
            gradsum = torch.empty(self.n_gradsum, dtype=torch.float64, device=self.device)
        self._ck(_lib.lib.tr_backward_mn(self._h, X.data_ptr(), self._vec(dP, N * self.n_classes, 'dP').data_ptr(), N,
                                         self._vec(theta, self.P, 'theta').data_ptr(),
                                         self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                         gradsum.data_ptr(), self._stream()))
        self._count()
        return gradsum

    def fwd_grad_mn(self, X, y, class_w, theta, w, nn_mask, beta, thr, gradsum=None, P=None):
        X = self._x(X)
        N = X.shape[0]
        if gradsum is None:
            gradsum = torch.empty(self.n_gradsum, dtype=torch.float64, device=self.device)
        self._ck(_lib.lib.tr_fwd_grad_mn(self._h, X.data_ptr(), self._vec(y, N, 'y', torch.int64).data_ptr(),
                                         self._vec(class_w, self.n_classes, 'class weights').data_ptr(), N,
                                         self._vec(theta, self.P, 'theta').data_ptr(),
                                         self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                         self._vec(gradsum, self.n_gradsum, 'gradsum', torch.float64).data_ptr(),
                                         P.data_ptr() if P is not None else None, self._stream()))
        self._count()
        return gradsum

    def finish(self, gradsum, grad_scale, loss_scale, theta, lambda_L2, nn_mask, beta, thr, grad=None, loss=None):
        if grad is None:
            grad = torch.empty(self.P, dtype=self.dtype, device=self.device)
        if loss is None:
            loss = torch.empty(2, dtype=torch.float64, device=self.device)
        self._ck(_lib.lib.tr_finish_grad(self._h, gradsum.data_ptr(), float(grad_scale), float(loss_scale),
                                         self._vec(theta, self.P, 'theta').data_ptr(), float(lambda_L2), nn_mask,
                                         beta, thr, grad.data_ptr(), loss.data_ptr(), self._stream()))
        self.launches += 1
        return grad, loss

    def adam_step(self, theta, grad, m, v, vmax, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                  lr_groups=None):
        """``lr_groups``: one learning rate per parameter group (the factors of theta in order, plus the bias
        for the standard model) — torch.optim.Adam with a list of parameter groups (hier:436-440)."""
        if lr_groups is not None:
            arr = (ctypes.c_double * len(lr_groups))(*[float(x) for x in lr_groups])
            self._ck(_lib.lib.tr_adam_step_groups(self._h, theta.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                                  v.data_ptr(), vmax.data_ptr() if vmax is not None else None,
                                                  int(step), arr, len(lr_groups), float(betas[0]), float(betas[1]),
                                                  float(eps), float(weight_decay), self._stream()))
        else:
            self._ck(_lib.lib.tr_adam_step(self._h, theta.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                                           vmax.data_ptr() if vmax is not None else None, int(step), float(lr),
                                           float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                           self._stream()))
        self.launches += 1

    def allreduce(self, buf, nccl_comm):
        """In-place sum of a float64 device vector over the ranks of ``nccl_comm`` (an ncclComm_t as an int),
        enqueued on the current stream by the library (tr_allreduce)."""
        self._ck(_lib.lib.tr_allreduce(self._h, self._vec(buf, buf.numel(), 'buf', torch.float64).data_ptr(),
                                       buf.numel(), ctypes.c_void_p(int(nccl_comm)), self._stream()))


    # -- L-BFGS vector kernels (see lbfgs.py) ------------------------------------------------
    def lbfgs_direction(self, g, prev_g, d, t, first, S, Y, lstate, history, scal4):
        self._ck(_lib.lib.tr_lbfgs_direction(self._h, g.data_ptr(), prev_g.data_ptr(), d.data_ptr(), float(t),
                                             1 if first else 0, S.data_ptr(), Y.data_ptr(), lstate.data_ptr(),
                                             int(history), scal4.data_ptr(), self._stream()))
        self.launches += 1

    def lbfgs_point(self, out, x, t, d):
        self._ck(_lib.lib.tr_lbfgs_point(self._h, out.data_ptr(), x.data_ptr(), float(t), d.data_ptr(),
                                         self._stream()))
        self.launches += 1

    def lbfgs_gtd(self, g, d, scal2):
        self._ck(_lib.lib.tr_lbfgs_gtd(self._h, g.data_ptr(), d.data_ptr() if d is not None else None,
                                       scal2.data_ptr(), self._stream()))
        self.launches += 1


class SpectralEngine(Engine):
    """Plan + workspace for the model of spectral_tensor_regression.py: X (T, W, D), y (T, n_out), ``rank_normal``
    ordinary CP components and ``rank_spectral`` components whose first-mode factor has ``complex_dim`` columns
    (tr_spec_* entry points).  The optimizer / all-reduce methods of ``Engine`` work on this handle unchanged."""

    def __init__(self, W, D, n_out, rank_normal, rank_spectral, complex_dim, dtype=torch.float32, device='cuda'):
        device = torch.device(device)
        if device.type != 'cuda':
            raise TRError(f"device '{device}': tensor_regression_b200 runs on CUDA devices only (no CPU path)")
        if not torch.cuda.is_available():
            raise TRError('no CUDA device available (tensor_regression_b200 has no CPU path)')
        if dtype not in _DT:
            raise TRError(f'dtype {dtype} not supported (float32 / float64)')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device, self.dtype = device, dtype
        self.W, self.Dm, self.n_out = int(W), int(D), int(n_out)
        self.rank_normal, self.rank_spectral, self.complex_dim = int(rank_normal), int(rank_spectral), int(complex_dim)
        self.dims = [self.W, self.Dm]
        self.rank = self.rank_normal + self.rank_spectral
        self.n_classes = 0
        h = ctypes.c_void_p()
        rc = _lib.lib.tr_spec_create(ctypes.byref(h), _DT[dtype], self.W, self.Dm, self.n_out, self.rank_normal,
                                     self.rank_spectral, self.complex_dim, device.index)
        if rc != 0:
            raise TRError(f'tr_spec_create failed ({rc}): {_lib.lib.tr_last_error(None).decode()}')
        self._h = h
        P, Pf, ngs = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._ck(_lib.lib.tr_param_count(h, ctypes.byref(P), ctypes.byref(Pf)))
        self._ck(_lib.lib.tr_gradsum_count(h, ctypes.byref(ngs)))
        self.P, self.Pf, self.n_gradsum = P.value, Pf.value, ngs.value
        self.launches = 0

    def block_shapes(self):
        """Shapes of the six factor blocks of theta in order (Bcp_n then Bcp_c, spectral:518-523)."""
        rn, rs, cc = self.rank_normal, self.rank_spectral, self.complex_dim
        return [(self.W, rn, 1), (self.Dm, rn, 1), (self.n_out, rn, 1),
                (self.W, rs, cc), (self.Dm, rs, 1), (self.n_out, rs, 1)]

    def fwd_grad(self, X, y, theta, w, nn_mask, beta, thr, gradsum=None, yhat=None):
        X = self._x(X)
        N = X.shape[0]
        if gradsum is None:
            gradsum = torch.empty(self.n_gradsum, dtype=torch.float64, device=self.device)
        self._ck(_lib.lib.tr_spec_fwd_grad(self._h, X.data_ptr(), self._vec(y, N * self.n_out, 'y').data_ptr(), N,
                                           self._vec(theta, self.P, 'theta').data_ptr(),
                                           self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                           self._vec(gradsum, self.n_gradsum, 'gradsum', torch.float64).data_ptr(),
                                           yhat.data_ptr() if yhat is not None else None, self._stream()))
        self._count()
        return gradsum

    def forward(self, X, theta, w, nn_mask, beta, thr, want=('yhat',)):
        """Returns a dict with the requested outputs among 'yhat' (the model of the fit), 'yhat_lin' (lin_model),
        'spec_pred' (spectral_model, (N, n_out)) and 'latents' ((N, rank_normal))."""
        X = self._x(X)
        N = X.shape[0]
        shapes = {'yhat': (N, self.n_out), 'yhat_lin': (N, self.n_out), 'spec_pred': (N, self.n_out),
                  'latents': (N, self.rank_normal)}
        out = {k: torch.empty(shapes[k], dtype=self.dtype, device=self.device) for k in want}
        ptr = lambda k: out[k].data_ptr() if k in out and out[k].numel() > 0 else None   # noqa: E731
        self._ck(_lib.lib.tr_spec_forward(self._h, X.data_ptr(), N, self._vec(theta, self.P, 'theta').data_ptr(),
                                          self._vec(w, self.rank, 'weights').data_ptr(), nn_mask, beta, thr,
                                          ptr('yhat'), ptr('yhat_lin'), ptr('spec_pred'), ptr('latents'), self._stream()))
        self._count()
        return out

    def launch_info(self):
        info = (ctypes.c_int64 * 8)()
        _lib.lib.tr_last_launch_info(self._h, info)
        keys = ['launches', 'grid_fwd', 'grid_grad', 'tiles_per_sample', 'df1_slabs', 'groups_grad', 'channels_per_pass',
                'vector_width']
        d = dict(zip(keys, [int(v) for v in info]))
        if d['df1_slabs'] < 0:
            d['stages'] = d.pop('groups_grad')
            d['path'] = 'single-pass (ring of whole samples in shared memory: window contraction, norm epilogue and both factor gradients in one kernel)'
        else:
            d['path'] = ('two-pass, epilogue fused into the first (window contraction + norm epilogue + second-mode gradient in one kernel)'
                         if d['df1_slabs'] == 0 else 'two-pass (window contraction, epilogue and second-mode gradient as separate kernels)')
        return d


class ShardedSum:
    """Sums the packed ``gradsum`` vector over the ranks that each hold a slice of the sample
    axis: one all-reduce of P+2 doubles per closure evaluation.  With
    ``group=None`` and no initialised default group this is the identity (single GPU)."""

    def __init__(self, group=None, enabled=None, engine=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.enabled = bool(enabled)
        self.engine = engine          # with an Engine and an NCCL group the sum goes through the C ABI (tr_allreduce)
        self._comm = None
        self.via = 'none' if not self.enabled else 'torch.distributed'

    @property
    def world(self):
        return self.dist.get_world_size(self.group) if self.enabled else 1

    def _nccl_comm(self, device):
        """ncclComm_t of the group's NCCL backend (ProcessGroupNCCL._comm_ptr), or None (gloo groups, a
        communicator that has not been created yet, older torch)."""
        if self._comm is not None:
            return self._comm or None
        try:
            pg = self.group if self.group is not None else self.dist.distributed_c10d._get_default_group()
            ptr = int(pg._get_backend(torch.device(device))._comm_ptr())
            if ptr:
                self._comm = ptr
                return ptr
        except Exception:
            pass
        return None

    def sum_(self, t):
        if self.enabled:
            comm = self._nccl_comm(t.device) if (self.engine is not None and t.is_cuda
                                                 and t.dtype == torch.float64) else None
            if comm:
                self.engine.allreduce(t, comm)
                self.via = 'tr_allreduce (C ABI, NCCL communicator of the process group)'
            else:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def total(self, value, device):
        """Global sum of a host scalar (N_total, sum of class weights) — done once, at fit start."""
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
        return float(self.sum_(t).item())


def host_pointer(X, dtype):
    """(address, keepalive) of a host array that tr_upload can read in place — a C-contiguous numpy array /
    np.memmap / CPU tensor whose element type already is ``dtype`` — else None."""
    if isinstance(X, torch.Tensor):
        if X.device.type == 'cpu' and X.dtype == dtype and X.is_contiguous() and X.numel() > 0:
            return X.data_ptr(), X
        return None
    if isinstance(X, np.ndarray) and X.flags['C_CONTIGUOUS'] and X.size > 0:
        want = torch.empty((), dtype=dtype).numpy().dtype
        if X.dtype == want:
            return int(X.ctypes.data), X
    return None


def upload_threads():
    """Host threads one upload may use: all CPUs of this process's affinity mask, shared between the ranks of a
    multi-GPU launch (LOCAL_WORLD_SIZE); 0 = let the library decide."""
    import os
    env = os.environ.get('TR_B200_UPLOAD_THREADS')
    if env:
        return int(env)
    local_world = int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1)
    if local_world <= 1:
        return 0
    try:
        cpus = len(os.sched_getaffinity(0))
    except Exception:
        cpus = os.cpu_count() or 1
    return max(1, min(16, cpus // local_world))


def upload_into(dst, ptr, nbytes, stream, chunk_bytes=0):
    """tr_upload: pageable (or pinned) host bytes at ``ptr`` -> the device tensor ``dst`` on ``stream``."""
    rc = _lib.lib.tr_upload(dst.data_ptr(), ctypes.c_void_p(ptr), nbytes, dst.device.index, upload_threads(),
                            chunk_bytes, ctypes.c_void_p(stream.cuda_stream))
    if rc != 0:
        raise TRError(f'tr_upload failed ({rc}): {_lib.lib.tr_host_last_error().decode()}')


def upload_stats():
    out = (ctypes.c_double * 4)()
    _lib.lib.tr_upload_stats(out)
    return {'seconds': out[0], 'host_fill_seconds': out[1], 'threads': int(out[2]), 'staged': bool(out[3])}


class HostStreamer:
    """Double-buffered host -> device streaming of a sample-major array that stays in host memory
    (numpy array, np.memmap, CPU tensor, or anything with ``.shape`` and ``[lo:hi]`` slicing).

    ``for lo, hi, xd in streamer.chunks(): ...`` yields device views valid until the next-but-one
    iteration; the copy of chunk i+1 (pinned staging buffer -> device, on a side stream) overlaps
    the kernels the caller enqueues for chunk i on the current stream.  Chunks that are already
    pinned tensors are copied from in place (no staging copy)."""

    def __init__(self, X, dtype, device, chunk_samples=None, chunk_bytes=1 << 30):
        self.X, self.dtype, self.device = X, dtype, torch.device(device)
        self.N = int(X.shape[0])
        self.sample_shape = tuple(int(d) for d in X.shape[1:])
        row = max(1, int(np.prod(self.sample_shape))) * torch.empty((), dtype=dtype).element_size()
        if chunk_samples is None:
            chunk_samples = max(1, chunk_bytes // row)
        self.step = int(max(1, min(max(self.N, 1), chunk_samples)))
        self.bytes_per_pass = self.N * row
        self._copy = torch.cuda.Stream(device=self.device)
        self._dev = [torch.empty((self.step, *self.sample_shape), dtype=dtype, device=self.device) for _ in range(2)]
        self._pin = [None, None]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._done = [torch.cuda.Event() for _ in range(2)]
        self._used = [False, False]
        self._row = row
        self._ptr = host_pointer(X, dtype)      # in-place source for tr_upload (parallel staging in the library)

    def _host_chunk(self, lo, hi, s):
        c = self.X[lo:hi]
        if not isinstance(c, torch.Tensor):
            c = torch.as_tensor(np.ascontiguousarray(c))
        if c.dtype == self.dtype and c.is_contiguous() and c.is_pinned():
            return c
        if self._pin[s] is None:
            self._pin[s] = torch.empty((self.step, *self.sample_shape), dtype=self.dtype, pin_memory=True)
        self._pin[s][:hi - lo].copy_(c)
        return self._pin[s][:hi - lo]

    def chunks(self):
        main = torch.cuda.current_stream(self.device)
        for i, lo in enumerate(range(0, self.N, self.step)):
            hi = min(self.N, lo + self.step)
            s = i % 2
            if self._used[s]:
                self._done[s].synchronize()       # kernels reading dev[s] finished; staging buffer reusable
            if self._ptr is not None:
                # the library stages the chunk through its pinned ring with several host threads and returns
                # when the DMA is done; the kernels of the previous chunk keep the GPU busy meanwhile
                upload_into(self._dev[s][:hi - lo], self._ptr[0] + lo * self._row, (hi - lo) * self._row, self._copy)
                self._ready[s].record(self._copy)
            else:
                src = self._host_chunk(lo, hi, s)
                self._copy.wait_event(self._done[s]) if self._used[s] else None
                with torch.cuda.stream(self._copy):
                    self._dev[s][:hi - lo].copy_(src, non_blocking=True)
                    self._ready[s].record(self._copy)
            main.wait_event(self._ready[s])
            yield lo, hi, self._dev[s][:hi - lo]
            self._done[s].record(main)
            self._used[s] = True


def upload_resident(X, dtype, device, chunk_bytes=256 << 20):
    """Host array (numpy / memmap / CPU tensor / anything with ``.shape`` and ``[lo:hi]``) -> ONE resident
    device tensor.  A contiguous array of the model dtype goes through the library's tr_upload (parallel
    staging, see csrc/tr_host.cu).  Anything else (dtype conversion needed, array-like objects) is copied chunk
    by chunk on a side stream straight into its place; chunks that are not
    already pinned go through two pinned staging buffers, so the host-side staging copy of chunk i+1
    overlaps the DMA of chunk i.  (``torch.as_tensor(X).to(device)`` does one synchronous pageable copy
    and needs X materialised as a single host tensor first.)"""
    device = torch.device(device)
    N = int(X.shape[0])
    shape = tuple(int(d) for d in X.shape[1:])
    row = max(1, int(np.prod(shape))) * torch.empty((), dtype=dtype).element_size()
    step = int(max(1, min(N, chunk_bytes // row)))
    out = torch.empty((N, *shape), dtype=dtype, device=device)
    main = torch.cuda.current_stream(device)
    hp = host_pointer(X, dtype)
    if hp is not None:
        # one contiguous host array of the right type (the reference's case: a numpy array / CPU tensor):
        # tr_upload stages it through pinned buffers with several host threads, or DMAs in place if it is pinned
        upload_into(out, hp[0], N * row, main)
        return out
    copy = torch.cuda.Stream(device=device)
    copy.wait_stream(main)
    pin, busy = [None, None], [None, None]
    for i, lo in enumerate(range(0, N, step)):
        hi = min(N, lo + step)
        s = i % 2
        c = X[lo:hi]
        if not isinstance(c, torch.Tensor):
            c = torch.as_tensor(np.ascontiguousarray(c))
        if not (c.dtype == dtype and c.is_contiguous() and c.is_pinned()):
            if busy[s] is not None:
                busy[s].synchronize()                      # the DMA that read this staging buffer is done
            if pin[s] is None:
                pin[s] = torch.empty((step, *shape), dtype=dtype, pin_memory=True)
            pin[s][:hi - lo].copy_(c)
            c = pin[s][:hi - lo]
        with torch.cuda.stream(copy):
            out[lo:hi].copy_(c, non_blocking=True)
            busy[s] = torch.cuda.Event()
            busy[s].record(copy)
    main.wait_stream(copy)
    return out
