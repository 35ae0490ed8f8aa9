"""Drop-in for the reference's ``standard_tensor_regression.py`` (CP linear regression):
same names, positional order, defaults and Kruskal-list layout; the compute runs in the
sm_100a kernels of libtrb200.so.

Reference lines each item mirrors are cited as ``std:<lines>``
(= /root/reference/standard_tensor_regression.py).

Deliberate differences (all on the host side):
  * ``device`` must be a CUDA device (default 'cuda'); there is no CPU path.
  * ``self.Bcp`` entries and ``self.bias`` are views into one flat device parameter vector
    ``self.theta`` (so the optimizer kernel and the all-reduce see one contiguous buffer); a
    user-supplied ``Bcp_init`` list is copied into it instead of being aliased.
  * matplotlib is imported lazily inside ``plot_outputs``.
  * keyword-only ``shard_group`` (constructor): when given (or when torch.distributed is
    initialised with world_size > 1 and ``shard_group='world'``), X / y passed to ``fit*`` are
    this rank's slice of the sample axis and the gradient sums are all-reduced.
"""
import numpy as np
import torch

from . import engine as _engine
from . import lbfgs as _lbfgs
from .engine import nn_mask_of

_DEFAULT_SOFTPLUS = {'beta': 50, 'threshold': 1}

####################################
######## Helper functions ##########
####################################


def make_BcpInit(B_dims, rank, non_negative, scale=1, device='cpu', dtype=torch.float32):
    """std:18-51 — orthogonal init drawn on the CPU (same RNG stream as the reference), then moved."""
    Bcp_init = [torch.nn.init.orthogonal_(torch.empty(B_dims[ii], rank, dtype=dtype), gain=scale).to(device)
                for ii in range(len(B_dims))]
    Bcp_init = [(Bcp_init[ii] + torch.std(Bcp_init[ii]) * 2 * non_negative[ii]) / ((non_negative[ii] + 1))
                if Bcp_init[0].shape[0] > 1 else Bcp_init[ii] for ii in range(len(Bcp_init))]
    return Bcp_init


def non_neg_fn(B_cp, non_negative, softplus_kwargs=None):
    """std:53-85 — generator applying softplus to the flagged list positions (tiny; plain torch)."""
    if softplus_kwargs is None:
        softplus_kwargs = _DEFAULT_SOFTPLUS
    for ii in range(len(B_cp)):
        if non_negative[ii]:
            yield torch.nn.functional.softplus(B_cp[ii], **softplus_kwargs)
        else:
            yield B_cp[ii]


# Engines (library handle + its cudaMalloc'ed workspace) used by the MODULE-LEVEL functions lin_model / model:
# a small LRU keyed by geometry; an evicted engine is closed (its workspace freed).  Estimator objects own their
# engine (one handle per model: no scratch buffers shared between models) and release it with close() / on
# deletion.  A handle's workspace is single-stream: do not drive one model from two CUDA streams at once.
_ENGINES = {}
_ENGINES_MAX = 4


def _engine_for(dims, rank, n_classes, dtype, device):
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    key = (tuple(int(d) for d in dims), int(rank), int(n_classes), dtype, str(device))
    eng = _ENGINES.pop(key, None)
    if eng is None:
        eng = _engine.Engine(dims, rank, n_classes, dtype, device)
        while len(_ENGINES) >= _ENGINES_MAX:
            _ENGINES.pop(next(iter(_ENGINES))).close()       # least recently used first (dicts keep insertion order)
    _ENGINES[key] = eng
    return eng


def release_engines():
    """Free the workspaces of the engines cached for lin_model / model."""
    while _ENGINES:
        _ENGINES.popitem()[1].close()


def _flatten(Bcp, bias, dtype, device):
    parts = [torch.as_tensor(b).detach().to(device=device, dtype=dtype).reshape(-1) for b in Bcp]
    if bias is not None:
        parts.append(torch.as_tensor(bias).detach().to(device=device, dtype=dtype).reshape(-1))
    return torch.cat(parts).contiguous()


class _LinModelFn(torch.autograd.Function):
    """lin_model with the reference's differentiability (grad wrt factors and bias, not X):
    forward = tr_forward_std, backward = tr_backward_std + tr_finish_grad (std:123-130, 372)."""

    @staticmethod
    def forward(ctx, X, weights, nn_mask, beta, thr, eng, bias, *Bcp):
        theta = _flatten(Bcp, bias, eng.dtype, eng.device)
        ctx.save_for_backward(X, theta, weights)
        ctx.meta = (nn_mask, beta, thr, eng, [tuple(b.shape) for b in Bcp])
        return eng.forward_std(X, theta, weights, nn_mask, beta, thr)

    @staticmethod
    def backward(ctx, dy):
        X, theta, weights = ctx.saved_tensors
        nn_mask, beta, thr, eng, shapes = ctx.meta
        gs = eng.backward_std(X, dy.contiguous().to(eng.dtype), theta, weights, nn_mask, beta, thr)
        grad, _ = eng.finish(gs, 1.0, 0.0, theta, 0.0, nn_mask, beta, thr)
        outs, off = [], 0
        for shp in shapes:
            n = shp[0] * shp[1]
            outs.append(grad[off:off + n].reshape(shp))
            off += n
        return (None, None, None, None, None, None, grad[off:off + 1].clone(), *outs)


def lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs=None):
    """std:87-130 — y_hat = inner(X, outer(softplus(Bcp))) + bias, shape (N,).

    X: (N, I_1..I_k) CUDA tensor; Bcp: list of (I_m, rank) tensors; differentiable with respect
    to Bcp and bias like the reference's expression (through the CUDA backward kernels)."""
    if softplus_kwargs is None:
        softplus_kwargs = _DEFAULT_SOFTPLUS
    if not isinstance(X, torch.Tensor):
        raise TypeError('X must be a torch.Tensor')
    rank = Bcp[0].shape[1]
    eng = _engine_for(X.shape[1:], rank, 0, X.dtype, X.device)
    w = torch.as_tensor(weights).detach().to(device=X.device, dtype=X.dtype).contiguous()
    bias_t = bias if isinstance(bias, torch.Tensor) else torch.tensor([float(bias)], dtype=X.dtype, device=X.device)
    mask = nn_mask_of(non_negative, len(Bcp))
    return _LinModelFn.apply(X, w, mask, float(softplus_kwargs['beta']), float(softplus_kwargs['threshold']), eng,
                             bias_t, *Bcp)


def L2_penalty(B_cp):
    """std:180-196 — sum of the UN-squared Frobenius norms of the raw factors."""
    ii = 0
    for comp in B_cp:
        ii += torch.sqrt(torch.sum(comp ** 2))
    return ii


####################################
########### Main class #############
####################################

class CP_linear_regression():
    def __init__(self,
                 X_shape,
                 dtype=torch.float32,
                 rank=5,
                 non_negative=False,
                 weights=None,
                 Bcp_init=None,
                 Bcp_init_scale=1,
                 bias_init=0,
                 device='cuda',
                 softplus_kwargs=None,
                 *,
                 shard_group=None):
        """std:204-303.  ``X_shape`` includes the sample axis; ``B_dims = X_shape[1:]``."""
        self.dtype = dtype
        self.rank = rank
        self.device = device
        B_dims = list(X_shape[1:])
        self._dims = [int(d) for d in B_dims]
        self._eng = None
        self._shard_group = shard_group

        dev = self._torch_device()
        if weights is None:
            self.weights = torch.ones((rank), dtype=self.dtype, requires_grad=False, device=dev)
        else:
            self.weights = torch.tensor(weights, dtype=self.dtype, requires_grad=False, device=dev)

        if softplus_kwargs is None:
            self.softplus_kwargs = {'beta': 50, 'threshold': 1}
        else:
            self.softplus_kwargs = softplus_kwargs

        if non_negative == True:  # noqa: E712  (reference semantics, std:281-286)
            self.non_negative = [True] * (len(X_shape))
        elif non_negative == False:  # noqa: E712
            self.non_negative = [False] * (len(X_shape))
        else:
            self.non_negative = non_negative

        if Bcp_init is None:
            Bcp0 = make_BcpInit(B_dims, self.rank, self.non_negative, scale=Bcp_init_scale, device='cpu',
                                dtype=self.dtype)
        else:
            Bcp0 = Bcp_init
        self._set_theta(Bcp0, torch.tensor([bias_init], dtype=self.dtype))
        self.loss_running = []

    # ---- parameter storage ---------------------------------------------------------------
    def _torch_device(self):
        dev = torch.device(self.device)
        if dev.type != 'cuda':
            raise _engine.TRError(f"device='{self.device}': tensor_regression_b200 has no CPU path; pass a CUDA device")
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        return dev

    def _set_theta(self, Bcp, bias):
        dev = self._torch_device()
        if len(Bcp) != len(self._dims):
            raise ValueError(f'Bcp has {len(Bcp)} factors, X_shape[1:] has {len(self._dims)} modes')
        for b, d in zip(Bcp, self._dims):
            if tuple(b.shape) != (d, self.rank):
                raise ValueError(f'factor shape {tuple(b.shape)} != {(d, self.rank)}')
        self.theta = _flatten(Bcp, bias, self.dtype, dev)
        sizes, offs = _engine.factor_offsets(self._dims, self.rank, 0)
        self.Bcp = [self.theta[offs[m]:offs[m + 1]].view(self._dims[m], self.rank) for m in range(len(self._dims))]
        self.bias = self.theta[offs[-1]:offs[-1] + 1]

    def _engine(self):
        if self._eng is None:
            self._eng = _engine.Engine(self._dims, self.rank, 0, self.dtype, self._torch_device())
        return self._eng

    def close(self):
        """Release the library handle and its device workspace (also done when the object is deleted)."""
        if self._eng is not None:
            self._eng.close()
            self._eng = None

    def _mask(self):
        return nn_mask_of(self.non_negative, len(self._dims))

    def _sp(self):
        return float(self.softplus_kwargs['beta']), float(self.softplus_kwargs['threshold'])

    def _prep_xy(self, X, y):
        dev = self._torch_device()
        if not isinstance(y, torch.Tensor):
            y = torch.as_tensor(y)
        if isinstance(X, torch.Tensor) and X.is_cuda:
            X = X.to(device=dev, dtype=self.dtype).contiguous()    # made contiguous once, not per iteration
        elif hasattr(X, 'shape') and len(X.shape) >= 1 and int(X.shape[0]) > 0:
            X = _engine.upload_resident(X, self.dtype, dev)        # host data: pinned, double-buffered upload
        else:
            X = torch.as_tensor(np.asarray(X)).to(device=dev, dtype=self.dtype).contiguous()
        y = y.to(device=dev, dtype=self.dtype).reshape(-1).contiguous()
        if X.shape[0] != y.shape[0]:
            raise ValueError('X.shape[0] must match len(y)')
        return X, y

    def _sharder(self):
        g = self._shard_group
        if g is None:
            return _engine.ShardedSum(enabled=False)
        return _engine.ShardedSum(group=None if g == 'world' else g, engine=self._engine())

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_eng'] = None
        st['_shard_group'] = None
        st['theta'] = self.theta.detach().cpu()
        st['weights'] = self.weights.detach().cpu()
        st.pop('Bcp'), st.pop('bias')
        return st

    def __setstate__(self, st):
        theta = st.pop('theta')
        self.__dict__.update(st)
        dev = self._torch_device()
        self.weights = self.weights.to(dev)
        sizes, offs = _engine.factor_offsets(self._dims, self.rank, 0)
        self._set_theta([theta[offs[m]:offs[m + 1]].view(self._dims[m], self.rank) for m in range(len(self._dims))],
                        theta[offs[-1]:])

    # ---- one closure evaluation (std:368-373): forward, loss, gradient -------------------
    def _closure_eval(self, X, y, lambda_L2, sharder, n_total):
        eng = self._engine()
        beta, thr = self._sp()
        gs = eng.fwd_grad_std(X, y, self.theta, self.weights, self._mask(), beta, thr)
        sharder.sum_(gs)
        return eng.finish(gs, 2.0 / n_total, 1.0 / n_total, self.theta, lambda_L2, self._mask(), beta, thr)

    def fit(self,
            X,
            y,
            lambda_L2=0.01,
            max_iter=1000,
            tol=1e-5,
            patience=10,
            verbose=False,
            running_loss_logging_interval=10,
            LBFGS_kwargs=None,
            *,
            out_of_core=False,
            chunk_samples=None):
        """std:305-398 — L-BFGS with torch.optim.LBFGS's algorithm and keyword arguments; history,
        two-loop recursion and vector algebra are device-resident (lbfgs.py), every closure
        evaluation is one tr_fwd_grad_std + tr_finish_grad.

        ``out_of_core=True`` (keyword-only extension; the reference's abandoned fit_batch_LBFGS, std:478-620,
        done exactly): X stays in host memory and every closure evaluation streams it to the device in chunks
        of ``chunk_samples``; the gradient is the exact full-batch sum, so the optimizer takes the same steps
        as with a resident X."""
        if LBFGS_kwargs is None:
            # the reference's "default" dict (std:353-362) is a bare expression: None raises there too
            raise TypeError('LBFGS_kwargs must be a dict of torch.optim.LBFGS keyword arguments (got None)')
        streamer = None
        if out_of_core:
            dev = self._torch_device()
            y = torch.as_tensor(y).to(device=dev, dtype=self.dtype).reshape(-1).contiguous()
            if int(X.shape[0]) != y.shape[0]:
                raise ValueError('X.shape[0] must match len(y)')
            streamer = _engine.HostStreamer(X, self.dtype, dev, chunk_samples=chunk_samples)
            self.h2d_bytes_per_iteration = streamer.bytes_per_pass
            n_local = y.shape[0]
        else:
            X, y = self._prep_xy(X, y)
            n_local = X.shape[0]
        sharder = self._sharder()
        n_total = sharder.total(n_local, y.device)
        eng = self._engine()
        beta, thr = self._sp()

        # one flat vector == the reference's Bcp + [bias] parameter list in the same order; the optimizer
        # is torch.optim.LBFGS's algorithm with device-resident history / two-loop recursion (lbfgs.py)
        optimizer = _lbfgs.LBFGS(eng, self.theta, **LBFGS_kwargs)
        gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=self.theta.device)
        gs_chunk = torch.empty_like(gs) if streamer is not None else None

        def closure(grad_out, loss_out):
            if streamer is None:
                eng.fwd_grad_std(X, y, self.theta, self.weights, self._mask(), beta, thr, gradsum=gs)
            else:
                gs.zero_()
                for lo, hi, xd in streamer.chunks():
                    eng.fwd_grad_std(xd, y[lo:hi], self.theta, self.weights, self._mask(), beta, thr, gradsum=gs_chunk)
                    gs.add_(gs_chunk)
            sharder.sum_(gs)
            eng.finish(gs, 2.0 / n_total, 1.0 / n_total, self.theta, lambda_L2, self._mask(), beta, thr,
                       grad=grad_out, loss=loss_out)

        def logged_loss():
            # extra forward without the penalty (std:380-382): one pass over X
            if streamer is None:
                y_hat = eng.forward_std(X, self.theta, self.weights, self._mask(), beta, thr)
            else:
                y_hat = torch.cat([eng.forward_std(xd, self.theta, self.weights, self._mask(), beta, thr)
                                   for _, _, xd in streamer.chunks()])
            sq = torch.sum((y_hat - y).to(torch.float64) ** 2).reshape(1)
            return (sharder.sum_(sq) / n_total).item(), y_hat

        convergence_reached = False
        for ii in range(max_iter):
            if ii % running_loss_logging_interval == 0:
                val, y_hat = logged_loss()
                self.loss_running.append(val)
                if verbose == 2:
                    print(f'Iteration: {ii}, Loss: {self.loss_running[-1]}  ;  Variance ratio (y_hat / y_true): {torch.var(y_hat).item() / torch.var(y).item()}')

            if ii > patience:
                if np.sum(np.abs(np.diff(self.loss_running[ii - patience:]))) < tol:
                    convergence_reached = True
                    break

            optimizer.step(closure)
        if (verbose == True) or (verbose >= 1):  # noqa: E712
            if convergence_reached:
                print('Convergence reached')
            else:
                print('Reached maximum number of iterations without convergence')
        return convergence_reached

    def fit_Adam(self, X, y,
                 lambda_L2=0.01,
                 max_iter=1000,
                 tol=1e-5,
                 patience=10,
                 verbose=False,
                 Adam_kwargs=None,
                 *,
                 out_of_core=False,
                 chunk_samples=None):
        """std:400-476 — Adam; forward+gradient kernels, fused penalty/normalise kernel, fused
        Adam kernel; one scalar device->host read per iteration (std:464).

        ``out_of_core=True`` (keyword-only extension; the reference's abandoned fit_batch_Adam,
        std:478-620, done exactly): X stays in host memory (numpy / memmap / CPU tensor) and is
        streamed to the device in chunks of ``chunk_samples`` every iteration; residuals depend
        only on a chunk's own samples, so one host->device pass per iteration gives the exact
        full-batch gradient."""
        if Adam_kwargs is None:
            raise TypeError('Adam_kwargs must be a dict of torch.optim.Adam keyword arguments (got None)')
        hyper = _adam_hyper(Adam_kwargs)
        if out_of_core:
            return self._fit_Adam_out_of_core(X, y, lambda_L2, max_iter, tol, patience, verbose, hyper,
                                              chunk_samples)
        X, y = self._prep_xy(X, y)
        sharder = self._sharder()
        n_total = sharder.total(X.shape[0], X.device)
        eng = self._engine()
        beta, thr = self._sp()
        m = torch.zeros_like(self.theta)
        v = torch.zeros_like(self.theta)
        vmax = torch.zeros_like(self.theta) if hyper['amsgrad'] else None
        gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=self.theta.device)
        grad = torch.empty_like(self.theta)
        loss = torch.empty(2, dtype=torch.float64, device=self.theta.device)
        y_hat = torch.empty_like(y) if verbose == 2 else None

        convergence_reached = False
        for ii in range(max_iter):
            eng.fwd_grad_std(X, y, self.theta, self.weights, self._mask(), beta, thr, gradsum=gs, yhat=y_hat)
            sharder.sum_(gs)
            eng.finish(gs, 2.0 / n_total, 1.0 / n_total, self.theta, lambda_L2, self._mask(), beta, thr,
                       grad=grad, loss=loss)
            eng.adam_step(self.theta, grad, m, v, vmax, ii + 1, lr=hyper['lr'], betas=hyper['betas'],
                          eps=hyper['eps'], weight_decay=hyper['weight_decay'])
            self.loss_running.append(loss[1].item())
            if verbose == 2:
                print(f'Iteration: {ii}, Loss: {self.loss_running[-1]}  ;  Variance ratio (y_hat / y_true): {torch.var(y_hat).item() / torch.var(y).item()}')
            if ii > patience:
                if np.sum(np.abs(np.diff(self.loss_running[ii - patience:]))) < tol:
                    convergence_reached = True
                    break
        if (verbose == True) or (verbose >= 1):  # noqa: E712
            if convergence_reached:
                print('Convergence reached')
            else:
                print('Reached maximum number of iterations without convergence')
        return convergence_reached

    def _fit_Adam_out_of_core(self, X, y, lambda_L2, max_iter, tol, patience, verbose, hyper, chunk_samples):
        dev = self._torch_device()
        y = torch.as_tensor(y).to(device=dev, dtype=self.dtype).reshape(-1).contiguous()
        if int(X.shape[0]) != y.shape[0]:
            raise ValueError('X.shape[0] must match len(y)')
        streamer = _engine.HostStreamer(X, self.dtype, dev, chunk_samples=chunk_samples)
        self.h2d_bytes_per_iteration = streamer.bytes_per_pass
        sharder = self._sharder()
        n_total = sharder.total(y.shape[0], dev)
        eng = self._engine()
        beta, thr = self._sp()
        m = torch.zeros_like(self.theta)
        v = torch.zeros_like(self.theta)
        vmax = torch.zeros_like(self.theta) if hyper['amsgrad'] else None
        gs = torch.zeros(eng.n_gradsum, dtype=torch.float64, device=dev)
        gs_chunk = torch.empty_like(gs)
        grad = torch.empty_like(self.theta)
        loss = torch.empty(2, dtype=torch.float64, device=dev)
        convergence_reached = False
        for ii in range(max_iter):
            gs.zero_()
            for lo, hi, xd in streamer.chunks():
                eng.fwd_grad_std(xd, y[lo:hi], self.theta, self.weights, self._mask(), beta, thr, gradsum=gs_chunk)
                gs.add_(gs_chunk)
            sharder.sum_(gs)
            eng.finish(gs, 2.0 / n_total, 1.0 / n_total, self.theta, lambda_L2, self._mask(), beta, thr,
                       grad=grad, loss=loss)
            eng.adam_step(self.theta, grad, m, v, vmax, ii + 1, lr=hyper['lr'], betas=hyper['betas'],
                          eps=hyper['eps'], weight_decay=hyper['weight_decay'])
            self.loss_running.append(loss[1].item())
            if verbose == 2:
                print(f'Iteration: {ii}, Loss: {self.loss_running[-1]}')
            if ii > patience:
                if np.sum(np.abs(np.diff(self.loss_running[ii - patience:]))) < tol:
                    convergence_reached = True
                    break
        if (verbose == True) or (verbose >= 1):  # noqa: E712
            if convergence_reached:
                print('Convergence reached')
            else:
                print('Reached maximum number of iterations without convergence')
        return convergence_reached

    ####################################
    ############ POST-HOC ##############
    ####################################

    def predict(self, X, Bcp=None, device=None, plot_pref=False):
        """std:628-687 — forward only; returns a numpy array (N,).  Host arrays are streamed to
        the device in bounded chunks."""
        dev = self._torch_device() if device is None else torch.device(device)
        if dev.type != 'cuda':
            raise _engine.TRError('predict runs on CUDA devices only')
        eng = self._engine()
        beta, thr = self._sp()
        if Bcp is None:
            theta = self.theta
        else:
            theta = _flatten(Bcp, self.bias, self.dtype, self.theta.device)
        if isinstance(X, torch.Tensor) and X.is_cuda:
            X = X.to(device=self.theta.device, dtype=self.dtype)
            return eng.forward_std(X, theta, self.weights, self._mask(), beta, thr).cpu().numpy()
        return _predict_streamed(X, self.dtype, self.theta.device,
                                 lambda xb: eng.forward_std(xb, theta, self.weights, self._mask(), beta, thr))

    def return_Bcp_final(self):
        """std:690-703."""
        Bcp = list(non_neg_fn(self.Bcp, self.non_negative, softplus_kwargs=self.softplus_kwargs))
        return [Bcp[ii].detach().cpu().numpy() for ii in range(len(Bcp))]

    def detach_Bcp(self):
        """std:705-715."""
        return [Bcp.detach().cpu().numpy() for Bcp in self.Bcp]

    def get_params(self):
        """std:717-731."""
        return {'weights': self.weights.detach().cpu().numpy(),
                'Bcp': self.detach_Bcp(),
                'non_negative': self.non_negative,
                'softplus_kwargs': self.softplus_kwargs,
                'rank': self.rank,
                'device': self.device,
                'loss_running': self.loss_running}

    def set_params(self, params):
        """std:733-750 (factors are copied into the flat device buffer)."""
        self.weights = torch.as_tensor(params['weights']).to(device=self._torch_device(), dtype=self.dtype)
        self.non_negative = params['non_negative']
        self.softplus_kwargs = params['softplus_kwargs']
        self.rank = params['rank']
        self.device = params['device']
        self.loss_running = params['loss_running']
        self._eng = None
        self._set_theta([torch.as_tensor(b) for b in params['Bcp']], self.bias.detach().cpu())

    def display_params(self):
        """std:752-765."""
        print('weights:', self.weights)
        print('Bcp:', self.Bcp)
        print('non_negative:', self.non_negative)
        print('softplus_kwargs:', self.softplus_kwargs)
        print('rank:', self.rank)
        print('device:', self.device)
        print('loss_running:', self.loss_running)

    def plot_outputs(self):
        """std:767-783."""
        import matplotlib.pyplot as plt
        plt.figure()
        plt.plot(self.loss_running)
        plt.xlabel('logged iteration')
        plt.ylabel('loss')
        plt.title('loss')

        Bcp_final = self.return_Bcp_final()
        fig, axs = plt.subplots(len(Bcp_final))
        for ii, val in enumerate(Bcp_final):
            axs[ii].set_title(f'factor {ii}')
            axs[ii].plot(val)
        fig.suptitle('components')


def _adam_hyper(Adam_kwargs):
    """torch.optim.Adam's keyword arguments that the fused kernel implements."""
    allowed = {'lr': 1e-3, 'betas': (0.9, 0.999), 'eps': 1e-8, 'weight_decay': 0, 'amsgrad': False}
    extra = {k: v for k, v in Adam_kwargs.items() if k not in allowed}
    for k, v in extra.items():
        if k in ('foreach', 'fused', 'capturable', 'differentiable') and not (k == 'differentiable' and v):
            continue            # execution-strategy switches of torch.optim.Adam: no numerical meaning here
        if k == 'maximize' and not v:
            continue
        raise TypeError(f"Adam_kwargs['{k}']={v!r} is not supported by the fused Adam kernel")
    h = dict(allowed)
    h.update({k: v for k, v in Adam_kwargs.items() if k in allowed})
    h['amsgrad'] = bool(h['amsgrad'])
    return h


def _predict_streamed(X, dtype, device, fwd, chunk_bytes=256 << 20):
    """Forward over a host array in chunks: pinned staging buffers, copy stream overlapped with
    the kernels.  Returns numpy (N,) or (N, C)."""
    if not hasattr(X, 'shape'):
        X = np.asarray(X)
    if X.shape[0] == 0:
        return fwd(torch.as_tensor(np.asarray(X)).to(device=device, dtype=dtype)).cpu().numpy()
    st = _engine.HostStreamer(X, dtype, device, chunk_bytes=chunk_bytes)
    outs = [fwd(xd) for _, _, xd in st.chunks()]
    return torch.cat(outs).cpu().numpy()
