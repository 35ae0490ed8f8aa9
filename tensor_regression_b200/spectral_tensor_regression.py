"""Drop-in for the fit path of the reference's ``spectral_tensor_regression.py``: CP regression of a multi-output
target y (T, n_out) on X (T, W, D) with ``rank_normal`` ordinary components and ``rank_spectral`` components whose
first-mode (window) factor carries a "complex" axis — in the model the fits optimise the contraction over W is
followed by a norm over that axis before the second contraction (``stepwise_spectral_model``).  Same names, positional order, defaults and Kruskal-list
layout (``Bcp_n`` / ``Bcp_c`` lists of (I, rank, complex) tensors); the compute runs in the sm_100a kernels of
libtrb200.so (``tr_spec_*``, csrc/tr_spectral.cuh).

Reference lines are cited as ``spectral:<lines>`` (= /root/reference/spectral_tensor_regression.py).

Covered: ``make_BcpInit``, ``non_neg_fn``, ``L2_penalty``, the four model functions (``lin_model``,
``spectral_model``, ``stepwise_latents_model``, ``stepwise_spectral_model``: forward values through the CUDA path,
not differentiable — the reference differentiates them only inside ``fit`` / ``fit_Adam``, which use the fused
kernels here), ``CP_linear_regression`` with ``fit`` (L-BFGS), ``fit_Adam``, ``predict``, ``predict_latents``,
``return_Bcp_final``, ``detach_Bcp``, ``get_params`` / ``set_params`` / ``display_params``, ``plot_outputs``.
Not covered: ``edge_clamp`` (unused by the reference, its only call is commented out), ``stepwise_linear_model``
(never called), the interactive ``verbose=3`` plotting, the commented-out batch fits.

Deliberate differences (host side): ``device`` must be a CUDA device; ``Bcp_n`` / ``Bcp_c`` entries and ``bias``
are views into one flat device vector ``theta`` (a supplied ``Bcp_init`` is copied, not aliased); ``predict`` uses
the model dtype for host X (the reference forces float32); ``set_params`` accepts the ``Bcp_n`` / ``Bcp_c`` keys that
``get_params`` produces (the reference's reads a ``'Bcp'`` key its own ``get_params`` never writes, spectral:1096).
"""
import numpy as np
import torch

from . import engine as _engine
from . import lbfgs as _lbfgs
from .standard_tensor_regression import _adam_hyper, _predict_streamed

_DEFAULT_SOFTPLUS = {'beta': 50, 'threshold': 1}

####################################
######## Helper functions ##########
####################################


def make_BcpInit(B_dims, rank, non_negative, complex_dims=None, scale=1, device='cpu', dtype=torch.float32):
    """spectral:17-60 — orthogonal init of (B_dims[i], rank, complex_dims[i]) tensors drawn on the CPU (same RNG
    stream as the reference), then moved."""
    if complex_dims is None:
        complex_dims = list([1] * len(B_dims))
    Bcp_init = [torch.nn.init.orthogonal_(torch.empty(B_dims[ii], rank, complex_dims[ii], dtype=dtype), gain=scale).to(device)
                for ii in range(len(B_dims))]
    Bcp_init = [(Bcp_init[ii] + torch.std(Bcp_init[ii]) * 2 * non_negative[ii]) / ((non_negative[ii] + 1))
                if Bcp_init[0].shape[0] > 1 else Bcp_init[ii] for ii in range(len(Bcp_init))]
    return Bcp_init


def non_neg_fn(B_cp, non_negative, softplus_kwargs=None):
    """spectral:62-94 — generator applying softplus to the flagged list positions."""
    if softplus_kwargs is None:
        softplus_kwargs = _DEFAULT_SOFTPLUS
    for ii in range(len(B_cp)):
        if non_negative[ii]:
            yield torch.nn.functional.softplus(B_cp[ii], **softplus_kwargs)
        else:
            yield B_cp[ii]


def L2_penalty(B_cp):
    """spectral:393-416 — sum of the UN-squared Frobenius norms of the raw factors."""
    ii = 0
    for comp in B_cp:
        ii += torch.sqrt(torch.sum(comp ** 2))
    return ii


def spec_nn_mask(non_negative):
    """Bit i of the mask = softplus on block i of theta = [Bcp_n[0..2] | Bcp_c[0..2]]: the reference applies
    non_negative[m] to position m of both lists (spectral:158-161, 377)."""
    mask = 0
    for m in range(min(3, len(non_negative))):
        if bool(non_negative[m]):
            mask |= (1 << m) | (1 << (m + 3))
    return mask


_ENGINES = {}
_ENGINES_MAX = 4


def _engine_for(W, D, n_out, rn, rs, cc, dtype, device):
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    key = (int(W), int(D), int(n_out), int(rn), int(rs), int(cc), dtype, str(device))
    eng = _ENGINES.pop(key, None)
    if eng is None:
        eng = _engine.SpectralEngine(W, D, n_out, rn, rs, cc, dtype, device)
        while len(_ENGINES) >= _ENGINES_MAX:
            _ENGINES.pop(next(iter(_ENGINES))).close()
    _ENGINES[key] = eng
    return eng


def release_engines():
    """Free the workspaces of the engines cached for the module-level model functions."""
    while _ENGINES:
        _ENGINES.popitem()[1].close()


def _flat(parts, dtype, device):
    return torch.cat([torch.as_tensor(p).detach().to(device=device, dtype=dtype).reshape(-1) for p in parts]).contiguous()


def _module_forward(X, Bcp_n, Bcp_c, weights, non_negative, bias, softplus_kwargs, want, spectral_weights=None):
    """One forward through the CUDA path for the module-level model functions: the list that is not given is
    replaced by an empty one (rank 0)."""
    if softplus_kwargs is None:
        softplus_kwargs = _DEFAULT_SOFTPLUS
    if not isinstance(X, torch.Tensor) or X.ndim != 3:
        raise TypeError('X must be a 3-D torch.Tensor (T, W, D)')
    ref = Bcp_n if Bcp_n is not None else Bcp_c
    if len(ref) != 3:
        raise ValueError('the spectral model takes three factors (window, feature, output)')
    W, D, n_out = (int(b.shape[0]) for b in ref)
    rn = int(Bcp_n[0].shape[1]) if Bcp_n is not None else 0
    rs = int(Bcp_c[0].shape[1]) if Bcp_c is not None else 0
    cc = int(Bcp_c[0].shape[2]) if Bcp_c is not None else 1
    dt, dev = X.dtype, X.device
    eng = _engine_for(W, D, n_out, rn, rs, cc, dt, dev)
    zero = lambda *shp: torch.zeros(shp, dtype=dt, device=dev)      # noqa: E731
    bn = list(Bcp_n) if Bcp_n is not None else [zero(W, 0, 1), zero(D, 0, 1), zero(n_out, 0, 1)]
    bc = list(Bcp_c) if Bcp_c is not None else [zero(W, 0, 1), zero(D, 0, 1), zero(n_out, 0, 1)]
    b = bias if bias is not None else zero(n_out)
    if not isinstance(b, torch.Tensor):
        b = torch.as_tensor(b)
    theta = _flat(bn + bc + [b.reshape(-1).expand(n_out) if b.numel() == 1 else b], dt, dev)
    w = torch.ones(rn + rs, dtype=dt, device=dev)
    if rn > 0:
        w[:rn] = torch.as_tensor(weights).detach().to(device=dev, dtype=dt).reshape(-1)[:rn]
    if rs > 0 and spectral_weights is not None:
        w[rn:] = torch.as_tensor(spectral_weights).detach().to(device=dev, dtype=dt).reshape(-1)[:rs]
    sp = (float(softplus_kwargs['beta']), float(softplus_kwargs['threshold']))
    return eng.forward(X.contiguous(), theta, w, spec_nn_mask(non_negative), sp[0], sp[1], want=(want,))[want]


def lin_model(X, Bcp, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:118-165 — inner(X, cp_to_tensor((weights, softplus?(Bcp[i][:,:,0])))) over the last two modes of X,
    squeezed, plus bias: (T, n_out)."""
    if Bcp[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    return _module_forward(X, Bcp, None, weights, non_negative, bias, softplus_kwargs, 'yhat_lin').squeeze()


def spectral_model(X, Bcp, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:168-221 — for every slice ii of the complex axis the complete CP contraction
    inner(X, cp_to_tensor((weights, [Bcp[0][:,:,ii], Bcp[1][:,:,0], Bcp[2][:,:,0]]))), then the norm over ii, plus
    bias: (T, n_out).  (The model ``predict`` adds to lin_model; the fit uses stepwise_spectral_model.)"""
    if Bcp[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    return _module_forward(X, None, Bcp, None, non_negative, bias, softplus_kwargs, 'spec_pred',
                           spectral_weights=weights).squeeze()


def stepwise_latents_model(X, Bcp, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:284-337 — the same two contractions on a list whose complex axis has length 1: (T, rank)."""
    if Bcp[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    if Bcp[0].shape[2] != 1:
        raise ValueError('stepwise_latents_model on a factor list with a complex axis is not supported '
                         '(the reference only calls it on Bcp_n, spectral:1031)')
    return _module_forward(X, Bcp, None, torch.ones(Bcp[0].shape[1]), non_negative, None, softplus_kwargs, 'latents')


def stepwise_spectral_model(X, Bcp, weights, non_negative, bias, softplus_kwargs=None):
    """spectral:339-390 — norm over the complex axis between the first and the second contraction, third factor,
    plus bias: (T, n_out).  ``weights`` is unused, as in the reference."""
    if Bcp[0].shape[1] == 0:
        return torch.zeros(1).to(X.device)
    return _module_forward(X, None, Bcp, weights, non_negative, bias, softplus_kwargs, 'yhat')


####################################
########### Main class #############
####################################

class CP_linear_regression():
    def __init__(self,
                 X_shape,
                 y_shape,
                 dtype=torch.float32,
                 rank_normal=1,
                 rank_spectral=1,
                 non_negative=False,
                 weights=None,
                 Bcp_init=None,
                 Bcp_init_scale=1,
                 n_complex_dim=0,
                 bias_init=0,
                 device='cuda',
                 softplus_kwargs=None,
                 *,
                 shard_group=None):
        """spectral:425-539.  ``X_shape`` = (T, W, D) and ``y_shape`` = (T, n_out) include the sample axis."""
        self.dtype = dtype
        self.device = device
        self._eng = None
        self._shard_group = shard_group
        if len(X_shape) != 3 or len(y_shape) != 2:
            raise ValueError('the spectral model takes X of shape (T, W, D) and y of shape (T, n_out) '
                             '(stepwise_spectral_model hard-codes these ranks, spectral:385-388)')
        dev = self._torch_device()

        if weights is None:
            self.weights = torch.ones((rank_normal + rank_spectral), dtype=self.dtype, requires_grad=False, device=dev)
        else:
            self.weights = torch.tensor(weights, dtype=self.dtype, requires_grad=False, device=dev)

        if softplus_kwargs is None:
            self.softplus_kwargs = {'beta': 50, 'threshold': 1}
        else:
            self.softplus_kwargs = softplus_kwargs

        self.rank_normal = rank_normal
        self.rank_spectral = rank_spectral
        self.rank = rank_normal + rank_spectral

        if non_negative == True:  # noqa: E712  (reference semantics, spectral:508-513)
            self.non_negative = [True] * (len(X_shape))
        elif non_negative == False:  # noqa: E712
            self.non_negative = [False] * (len(X_shape))
        else:
            self.non_negative = non_negative

        self.y_shape = y_shape
        B_dims = list(X_shape[1:]) + list(y_shape[1:])
        self._B_dims = [int(d) for d in B_dims]
        complex_dims = list([n_complex_dim + 1] + [1] * (len(B_dims) - 1))
        self._complex_dim = int(n_complex_dim) + 1
        if Bcp_init is None:
            Bn = make_BcpInit(B_dims, self.rank_normal, self.non_negative, complex_dims=None, scale=Bcp_init_scale,
                              device='cpu', dtype=self.dtype)
            Bc = make_BcpInit(B_dims, self.rank_spectral, self.non_negative, complex_dims=complex_dims,
                              scale=Bcp_init_scale, device='cpu', dtype=self.dtype)
        else:
            Bn, Bc = Bcp_init[0], Bcp_init[1]
        # the reference's bias starts at zero whatever bias_init says (spectral:518)
        self._set_theta(Bn, Bc, torch.zeros(self._B_dims[2], dtype=self.dtype))
        self.loss_running = []

    # ---- parameter storage ---------------------------------------------------------------
    def _torch_device(self):
        dev = torch.device(self.device)
        if dev.type != 'cuda':
            raise _engine.TRError(f"device='{self.device}': tensor_regression_b200 has no CPU path; pass a CUDA device")
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        return dev

    def _shapes(self):
        W, D, NO = self._B_dims
        rn, rs, cc = self.rank_normal, self.rank_spectral, self._complex_dim
        return [(W, rn, 1), (D, rn, 1), (NO, rn, 1), (W, rs, cc), (D, rs, 1), (NO, rs, 1)]

    def _set_theta(self, Bn, Bc, bias):
        dev = self._torch_device()
        shapes = self._shapes()
        blocks = list(Bn) + list(Bc)
        if len(blocks) != 6:
            raise ValueError('Bcp_init must be [Bcp_n, Bcp_c] with three factors each')
        for b, shp in zip(blocks, shapes):
            if tuple(b.shape) != shp:
                raise ValueError(f'factor shape {tuple(b.shape)} != {shp}')
        self.theta = _flat(blocks + [bias], self.dtype, dev)
        offs = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in shapes])]).astype(np.int64)
        views = [self.theta[offs[i]:offs[i + 1]].view(shapes[i]) for i in range(6)]
        self.Bcp_n, self.Bcp_c = views[:3], views[3:]
        self.bias = self.theta[offs[6]:offs[6] + self._B_dims[2]]

    def _engine(self):
        if self._eng is None:
            W, D, NO = self._B_dims
            self._eng = _engine.SpectralEngine(W, D, NO, self.rank_normal, self.rank_spectral, self._complex_dim,
                                               self.dtype, self._torch_device())
        return self._eng

    def close(self):
        """Release the library handle and its device workspace (also done when the object is deleted)."""
        if self._eng is not None:
            self._eng.close()
            self._eng = None

    def _mask(self):
        return spec_nn_mask(self.non_negative)

    def _sp(self):
        return float(self.softplus_kwargs['beta']), float(self.softplus_kwargs['threshold'])

    def _prep_xy(self, X, y):
        dev = self._torch_device()
        if not isinstance(y, torch.Tensor):
            y = torch.as_tensor(np.asarray(y))
        if isinstance(X, torch.Tensor) and X.is_cuda:
            X = X.to(device=dev, dtype=self.dtype).contiguous()
        elif hasattr(X, 'shape') and len(X.shape) >= 1 and int(X.shape[0]) > 0:
            X = _engine.upload_resident(X, self.dtype, dev)
        else:
            X = torch.as_tensor(np.asarray(X)).to(device=dev, dtype=self.dtype).contiguous()
        if list(X.shape[1:]) != self._B_dims[:2]:
            raise ValueError(f'X.shape[1:]={list(X.shape[1:])} does not match the model ({self._B_dims[:2]})')
        y = y.to(device=dev, dtype=self.dtype).reshape(X.shape[0], -1).contiguous()
        if y.shape[1] != self._B_dims[2]:
            raise ValueError(f'y has {y.shape[1]} outputs, the model {self._B_dims[2]}')
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
        st.pop('Bcp_n'), st.pop('Bcp_c'), st.pop('bias')
        return st

    def __setstate__(self, st):
        theta = st.pop('theta')
        self.__dict__.update(st)
        self.weights = self.weights.to(self._torch_device())
        shapes = self._shapes()
        offs = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in shapes])]).astype(np.int64)
        blocks = [theta[offs[i]:offs[i + 1]].view(shapes[i]) for i in range(6)]
        self._set_theta(blocks[:3], blocks[3:], theta[offs[6]:])

    # ---- fits ----------------------------------------------------------------------------------
    def fit(self,
            X,
            y,
            lambda_L2=0.01,
            max_iter=1000,
            tol=1e-5,
            patience=10,
            verbose=False,
            running_loss_logging_interval=10,
            LBFGS_kwargs=None):
        """spectral:541-649 — L-BFGS over Bcp_n + Bcp_c + [bias] (torch.optim.LBFGS's algorithm, lbfgs.py); every
        closure evaluation is one tr_spec_fwd_grad + tr_finish_grad."""
        if LBFGS_kwargs is None:
            # the reference's "default" dict (spectral:590-599) is a bare expression: None raises there too
            raise TypeError('LBFGS_kwargs must be a dict of torch.optim.LBFGS keyword arguments (got None)')
        X, y = self._prep_xy(X, y)
        sharder = self._sharder()
        n_total = sharder.total(X.shape[0], X.device) * self._B_dims[2]          # MSELoss: mean over (T, n_out)
        eng = self._engine()
        beta, thr = self._sp()
        optimizer = _lbfgs.LBFGS(eng, self.theta, **LBFGS_kwargs)
        gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=self.theta.device)

        def closure(grad_out, loss_out):
            eng.fwd_grad(X, y, self.theta, self.weights, self._mask(), beta, thr, gradsum=gs)
            sharder.sum_(gs)
            eng.finish(gs, 2.0 / n_total, 1.0 / n_total, self.theta, lambda_L2, self._mask(), beta, thr,
                       grad=grad_out, loss=loss_out)

        def logged_loss():
            y_hat = eng.forward(X, self.theta, self.weights, self._mask(), beta, thr, want=('yhat',))['yhat']
            sq = torch.sum((y_hat - y).to(torch.float64) ** 2).reshape(1)
            return (sharder.sum_(sq) / n_total).item() + lambda_L2 * self._penalty(), y_hat

        convergence_reached = False
        for ii in range(max_iter):
            if ii % running_loss_logging_interval == 0:
                val, y_hat = logged_loss()
                self.loss_running.append(val)
                if verbose == 2 or verbose == 3:
                    print(f'Iteration: {ii}, Loss: {self.loss_running[-1]}  ;  Variance ratio (y_hat / y_true): {torch.var(y_hat).item() / torch.var(y).item()}')
            if len(self.loss_running) > patience:
                if np.sum(np.abs(np.diff(self.loss_running[-patience + 1:]))) < tol:
                    convergence_reached = True
                    break
            elif np.isnan(self.loss_running[-1]):
                convergence_reached = False
                print('Loss is NaN. Stopping.')
                break

            optimizer.step(closure)
        if (verbose == True) or (verbose >= 1):  # noqa: E712
            if convergence_reached:
                print('Convergence reached')
            else:
                print('Reached maximum number of iterations without convergence')
        return convergence_reached

    def _penalty(self):
        """L2_penalty(Bcp_n) + L2_penalty(Bcp_c) of the current parameters (the logged loss includes it, spectral:614)."""
        return float(sum(torch.sqrt(torch.sum(b.to(torch.float64) ** 2)) for b in self.Bcp_n + self.Bcp_c))

    def fit_Adam(self, X, y,
                 lambda_L2=0.01,
                 max_iter=1000,
                 tol=1e-5,
                 patience=10,
                 verbose=False,
                 plotting_interval=100,
                 Adam_kwargs=None):
        """spectral:652-762 — Adam over Bcp_n + Bcp_c + [bias]: forward + gradient kernels, fused penalty / normalise
        kernel, fused Adam kernel; one scalar device->host read per iteration."""
        if Adam_kwargs is None:
            raise TypeError('Adam_kwargs must be a dict of torch.optim.Adam keyword arguments (got None)')
        hyper = _adam_hyper(Adam_kwargs)
        X, y = self._prep_xy(X, y)
        sharder = self._sharder()
        n_total = sharder.total(X.shape[0], X.device) * self._B_dims[2]
        eng = self._engine()
        beta, thr = self._sp()
        m = torch.zeros_like(self.theta)
        v = torch.zeros_like(self.theta)
        vmax = torch.zeros_like(self.theta) if hyper['amsgrad'] else None
        gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=self.theta.device)
        grad = torch.empty_like(self.theta)
        loss = torch.empty(2, dtype=torch.float64, device=self.theta.device)
        y_hat = torch.empty_like(y) if verbose in (2, 3) else None

        convergence_reached = False
        for ii in range(max_iter):
            eng.fwd_grad(X, y, self.theta, self.weights, self._mask(), beta, thr, gradsum=gs, yhat=y_hat)
            sharder.sum_(gs)
            eng.finish(gs, 2.0 / n_total, 1.0 / n_total, self.theta, lambda_L2, self._mask(), beta, thr,
                       grad=grad, loss=loss)
            eng.adam_step(self.theta, grad, m, v, vmax, ii + 1, lr=hyper['lr'], betas=hyper['betas'],
                          eps=hyper['eps'], weight_decay=hyper['weight_decay'])
            self.loss_running.append(loss[1].item())
            if verbose == 2 or (verbose == 3 and ii % plotting_interval == 0):
                print(f'Iteration: {ii}, Loss: {self.loss_running[-1]}  ;  Variance ratio (y_hat / y_true): {torch.var(y_hat).item() / torch.var(y).item()}')
            if ii > patience:
                if np.sum(np.abs(np.diff(self.loss_running[ii - patience:]))) < tol:
                    convergence_reached = True
                    break
            elif np.isnan(self.loss_running[-1]):
                convergence_reached = False
                print('Loss is NaN. Stopping.')
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

    def _theta_of(self, Bcp):
        if Bcp is None:
            return self.theta
        return _flat(list(Bcp[0]) + list(Bcp[1]) + [self.bias], self.dtype, self.theta.device)

    def predict(self, X, Bcp=None, device=None, plot_pref=False):
        """spectral:895-963 — lin_model(X, Bcp_n, ...) + spectral_model(X, Bcp_c, ...): the reference's predict uses
        ``spectral_model`` (norm over the complex axis of the COMPLETE contraction, rank weights applied), not the
        ``stepwise_spectral_model`` the fit optimises, and both terms add the bias.  Returns a CPU tensor like the
        reference."""
        eng = self._engine()
        beta, thr = self._sp()
        theta = self._theta_of(Bcp)
        want = tuple(k for k, r in (('yhat_lin', self.rank_normal), ('spec_pred', self.rank_spectral)) if r > 0)

        def fwd(xb):
            o = eng.forward(xb, theta, self.weights, self._mask(), beta, thr, want=want)
            lin = o['yhat_lin'].squeeze() if 'yhat_lin' in o else torch.zeros(1, device=xb.device)
            spec = o['spec_pred'].squeeze() if 'spec_pred' in o else torch.zeros(1, device=xb.device)
            return lin + spec

        if isinstance(X, torch.Tensor) and X.is_cuda:
            return fwd(X.to(device=self.theta.device, dtype=self.dtype)).cpu().detach()
        return torch.as_tensor(_predict_streamed(X, self.dtype, self.theta.device, fwd))

    def predict_latents(self, X, Bcp=None, device=None, plot_pref=False):
        """spectral:966-1034 — stepwise_latents_model on the normal components: (T, rank_normal) numpy array."""
        eng = self._engine()
        beta, thr = self._sp()
        theta = self._theta_of(Bcp)
        if self.rank_normal == 0:
            return torch.zeros(1).numpy()

        def fwd(xb):
            return eng.forward(xb, theta, self.weights, self._mask(), beta, thr, want=('latents',))['latents']

        if isinstance(X, torch.Tensor) and X.is_cuda:
            return fwd(X.to(device=self.theta.device, dtype=self.dtype)).cpu().detach().numpy()
        return _predict_streamed(X, self.dtype, self.theta.device, fwd)

    def return_Bcp_final(self):
        """spectral:1038-1053."""
        Bcp_n = list(non_neg_fn(self.Bcp_n, self.non_negative, softplus_kwargs=self.softplus_kwargs))
        Bcp_c = list(non_neg_fn(self.Bcp_c, self.non_negative, softplus_kwargs=self.softplus_kwargs))
        return ([b.detach().cpu().numpy() for b in Bcp_n], [b.detach().cpu().numpy() for b in Bcp_c])

    def detach_Bcp(self):
        """spectral:1055-1066."""
        return ([b.detach().cpu().numpy() for b in self.Bcp_n], [b.detach().cpu().numpy() for b in self.Bcp_c])

    def get_params(self):
        """spectral:1068-1083."""
        return {'weights': self.weights.detach().cpu().numpy(),
                'Bcp_n': self.detach_Bcp()[0],
                'Bcp_c': self.detach_Bcp()[1],
                'non_negative': self.non_negative,
                'softplus_kwargs': self.softplus_kwargs,
                'rank': self.rank,
                'device': self.device,
                'loss_running': self.loss_running}

    def set_params(self, params):
        """spectral:1085-1102 (factors are copied into the flat device buffer; takes the keys get_params writes)."""
        self.weights = torch.as_tensor(params['weights']).to(device=self._torch_device(), dtype=self.dtype)
        self.non_negative = params['non_negative']
        self.softplus_kwargs = params['softplus_kwargs']
        self.rank = params['rank']
        self.device = params['device']
        self.loss_running = params['loss_running']
        self._eng = None
        self._set_theta([torch.as_tensor(b) for b in params['Bcp_n']], [torch.as_tensor(b) for b in params['Bcp_c']],
                        self.bias.detach().cpu())

    def display_params(self):
        """spectral:1104-1117."""
        print('weights:', self.weights)
        print('Bcp_n:', self.Bcp_n)
        print('Bcp_c:', self.Bcp_c)
        print('non_negative:', self.non_negative)
        print('softplus_kwargs:', self.softplus_kwargs)
        print('rank:', self.rank)
        print('device:', self.device)
        print('loss_running:', self.loss_running)

    def plot_outputs(self):
        """spectral:1119-1149."""
        import matplotlib.pyplot as plt
        plt.figure()
        plt.plot(self.loss_running)
        plt.xlabel('logged iteration')
        plt.ylabel('loss')
        plt.title('loss')
        Bcp_n_final, Bcp_c_final = self.return_Bcp_final()
        if self.rank_normal > 0:
            fig_n, axs = plt.subplots(len(Bcp_n_final))
            for ii, val in enumerate(Bcp_n_final):
                axs[ii].set_title(f'factor {ii+1}')
                axs[ii].plot(val.squeeze())
            fig_n.suptitle('Bcp_n components')
        if self.rank_spectral > 0:
            fig_c, axs = plt.subplots(len(Bcp_c_final[1:]) + Bcp_c_final[0].shape[1])
            jj = 0
            for ii, val in enumerate(Bcp_c_final):
                axs[ii + jj].set_title(f'factor {ii+1}')
                if ii == 0:
                    for jj in range(val.shape[1]):
                        axs[jj].plot(val[:, jj, :].squeeze())
                else:
                    axs[ii + jj].plot(val.squeeze())
            fig_c.suptitle('Bcp_c components')
