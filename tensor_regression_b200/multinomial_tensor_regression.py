"""Drop-in for the reference's ``multinomial_tensor_regression.py`` (CP multinomial logistic
regression): same names, positional order, defaults and Kruskal-list layout (k feature
factors + one (n_classes, rank) class factor LAST); compute in the sm_100a kernels of
libtrb200.so.  Reference lines are cited as ``mn:<lines>``.

Parity-critical quirks kept: the loss is CrossEntropyLoss applied to the
already-softmaxed output (softmax twice, mn:180,187 + 364-366); the penalty is the sum of
un-squared Frobenius norms (mn:201-204); there is no bias (mn:355,447); class ``weights`` are
required by ``fit`` / ``fit_Adam`` (None raises, mn:365,449).

Host-side differences: CUDA device required (default 'cuda'); ``Bcp`` entries are views into
one flat device vector ``theta``; lazy matplotlib; keyword-only ``shard_group`` (see
standard_tensor_regression.py).
"""
import numpy as np
import torch

from . import engine as _engine
from . import lbfgs as _lbfgs
from .engine import nn_mask_of
from .standard_tensor_regression import (_adam_hyper, _engine_for, _flatten, _predict_streamed)

_DEFAULT_SOFTPLUS = {'beta': 50, 'threshold': 1}

####################################
######## Useful functions ##########
####################################


def squeeze_integers(intVec):
    """mn:18-37 — make integers consecutive from 0: [7,2,7,4,1] -> [3,2,3,1,0]."""
    uniques = np.unique(intVec)
    unique_positions = np.arange(len(uniques))
    return unique_positions[np.array([np.where(intVec[ii] == uniques)[0] for ii in range(len(intVec))]).squeeze()]


def confusion_matrix(y_hat, y_true):
    """mn:45-65 — columns normalised by the true-class counts."""
    n_classes = np.max(y_true) + 1
    if y_hat.ndim == 1:
        y_hat = idx_to_oneHot(y_hat, n_classes)
    cmat = y_hat.T @ idx_to_oneHot(y_true, n_classes)
    return cmat / np.sum(cmat, axis=0)[None, :]


def idx_to_oneHot(arr, n_classes=None):
    """mn:67-86."""
    if n_classes is None:
        n_classes = np.max(arr) + 1
    oneHot = np.zeros((arr.size, n_classes))
    oneHot[np.arange(arr.size), arr] = 1
    return oneHot


def make_BcpInit(B_dims, rank, non_negative, scale=1, device='cpu'):
    """mn:88-114 — uniform init, drawn on the CPU RNG like the reference."""
    Bcp_init = [(torch.rand((B_dims[ii], rank)) * scale - (1 - non_negative[ii]) * (scale / 2)).to(device)
                for ii in range(len(B_dims))]
    return Bcp_init


def non_neg_fn(B_cp, non_negative, softplus_kwargs=None):
    """mn:116-146."""
    if softplus_kwargs is None:
        softplus_kwargs = _DEFAULT_SOFTPLUS
    for ii in range(len(B_cp)):
        if non_negative[ii]:
            yield torch.nn.functional.softplus(B_cp[ii], **softplus_kwargs)
        else:
            yield B_cp[ii]


class _ModelFn(torch.autograd.Function):
    """model with the reference's differentiability (gradients wrt the factors, not X):
    forward = tr_forward_mn, backward = tr_backward_mn + tr_finish_grad (mn:180-187, 361)."""

    @staticmethod
    def forward(ctx, X, weights, nn_mask, beta, thr, eng, *Bcp):
        theta = _flatten(Bcp, None, eng.dtype, eng.device)
        ctx.save_for_backward(X, theta, weights)
        ctx.meta = (nn_mask, beta, thr, eng, [tuple(b.shape) for b in Bcp])
        P, _ = eng.forward_mn(X, theta, weights, nn_mask, beta, thr, want_pred=False)
        return P

    @staticmethod
    def backward(ctx, dP):
        X, theta, weights = ctx.saved_tensors
        nn_mask, beta, thr, eng, shapes = ctx.meta
        gs = eng.backward_mn(X, dP.contiguous().to(eng.dtype), theta, weights, nn_mask, beta, thr)
        grad, _ = eng.finish(gs, 1.0, 0.0, theta, 0.0, nn_mask, beta, thr)
        outs, off = [], 0
        for shp in shapes:
            n = shp[0] * shp[1]
            outs.append(grad[off:off + n].reshape(shp))
            off += n
        return (None, None, None, None, None, None, *outs)


def model(X, Bcp, weights, non_negative, softplus_kwargs=None):
    """mn:148-187 — softmax(inner(X, outer(softplus(Bcp)), n_modes=len(Bcp)-1), dim=1): (N, C)
    probabilities; the LAST factor is the (C, rank) class factor.  Differentiable with respect to the
    factors like the reference's expression (through the CUDA backward kernels)."""
    if softplus_kwargs is None:
        softplus_kwargs = _DEFAULT_SOFTPLUS
    if not isinstance(X, torch.Tensor):
        raise TypeError('X must be a torch.Tensor')
    rank, C = Bcp[0].shape[1], Bcp[-1].shape[0]
    eng = _engine_for(X.shape[1:], rank, C, X.dtype, X.device)
    w = torch.as_tensor(weights).detach().to(device=X.device, dtype=X.dtype).contiguous()
    return _ModelFn.apply(X, w, nn_mask_of(non_negative, len(Bcp)), float(softplus_kwargs['beta']),
                          float(softplus_kwargs['threshold']), eng, *Bcp)


def L2_penalty(B_cp):
    """mn:189-204."""
    ii = 0
    for comp in B_cp:
        ii += torch.sqrt(torch.sum(comp ** 2))
    return ii


####################################
########### Main class #############
####################################

class CP_logistic_regression():
    def __init__(self, X, y, rank=5, non_negative=False, weights=None, Bcp_init=None, Bcp_init_scale=1,
                 device='cuda', softplus_kwargs=None, *, shard_group=None, n_classes=None, out_of_core=False,
                 chunk_samples=None):
        """mn:212-286.  X is stored as float32, y as int64 (mn:255-256).  With ``shard_group`` X / y
        are this rank's slice of the sample axis; ``n_classes`` may then be given explicitly
        (otherwise the number of distinct labels is all-reduced as a max of label+1).
        ``out_of_core=True`` (keyword-only extension): X stays in host memory (numpy / memmap / CPU
        tensor / anything with ``.shape`` and ``[lo:hi]``) and is streamed to the device in chunks of
        ``chunk_samples`` on every pass; gradients are exact full-batch sums."""
        self.device = device
        dev = self._torch_device()
        self._out_of_core = bool(out_of_core)
        self._chunk_samples = chunk_samples
        if self._out_of_core:
            self.X = X
        elif isinstance(X, torch.Tensor) and X.is_cuda:
            self.X = X.to(device=dev, dtype=torch.float32).contiguous()
        elif hasattr(X, 'shape') and len(X.shape) >= 1 and int(X.shape[0]) > 0:
            self.X = _engine.upload_resident(X, torch.float32, dev)   # host data: pinned, double-buffered upload
        else:
            self.X = torch.as_tensor(X, dtype=torch.float32).to(dev)
        self.y = torch.as_tensor(y, dtype=torch.long).to(dev)
        self._shard_group = shard_group
        self._eng = None

        if weights is None:
            self.weights = torch.ones((rank), device=dev)
        else:
            self.weights = torch.tensor(weights, dtype=torch.float32, device=dev)

        if softplus_kwargs is None:
            self.softplus_kwargs = {'beta': 50, 'threshold': 1}
        else:
            self.softplus_kwargs = softplus_kwargs

        self.rank = rank

        ndim = len(self.X.shape)
        if non_negative == True:  # noqa: E712  (mn:271-276)
            self.non_negative = [True] * ndim
        elif non_negative == False:  # noqa: E712
            self.non_negative = [False] * ndim
        else:
            self.non_negative = non_negative

        if n_classes is not None:
            self.n_classes = int(n_classes)
        elif shard_group is not None:
            t = torch.tensor([float(self.y.max().item()) + 1.0 if self.y.numel() else 0.0], dtype=torch.float64,
                             device=dev)
            sh = self._sharder()
            if sh.enabled:
                sh.dist.all_reduce(t, op=sh.dist.ReduceOp.MAX, group=sh.group)
            self.n_classes = int(t.item())
        else:
            self.n_classes = len(torch.unique(self.y))
        # labels index the class factor and the class weights: the reference fails with "Target ... is out of
        # bounds" from CrossEntropyLoss (mn:366) for labels outside [0, n_classes); here the check is up front
        if self.y.numel():
            lo, hi = int(self.y.min().item()), int(self.y.max().item())
            if lo < 0 or hi >= self.n_classes:
                raise ValueError(f'class labels must lie in [0, n_classes={self.n_classes}): found min {lo}, max {hi} '
                                 f'(use squeeze_integers to make labels consecutive from 0)')
        self._dims = [int(d) for d in self.X.shape[1:]]
        B_dims = np.concatenate((np.array(self.X.shape[1:]), [self.n_classes]))
        if Bcp_init is None:
            Bcp0 = make_BcpInit(B_dims, self.rank, self.non_negative, scale=Bcp_init_scale, device='cpu')
        else:
            Bcp0 = Bcp_init
        self._set_theta(Bcp0)
        self.loss_running = []

    # ---- parameter storage ---------------------------------------------------------------
    def _torch_device(self):
        dev = torch.device(self.device)
        if dev.type != 'cuda':
            raise _engine.TRError(f"device='{self.device}': tensor_regression_b200 has no CPU path; pass a CUDA device")
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        return dev

    def _set_theta(self, Bcp):
        dev = self._torch_device()
        want = self._dims + [self.n_classes]
        if len(Bcp) != len(want):
            raise ValueError(f'Bcp has {len(Bcp)} factors, expected {len(want)} (feature modes + class factor)')
        for b, d in zip(Bcp, want):
            if tuple(b.shape) != (d, self.rank):
                raise ValueError(f'factor shape {tuple(b.shape)} != {(d, self.rank)}')
        self.theta = _flatten(Bcp, None, torch.float32, dev)
        sizes, offs = _engine.factor_offsets(self._dims, self.rank, self.n_classes)
        self.Bcp = [self.theta[offs[m]:offs[m + 1]].view(want[m], self.rank) for m in range(len(want))]

    def _engine(self):
        if self._eng is None:
            self._eng = _engine.Engine(self._dims, self.rank, self.n_classes, torch.float32, self._torch_device())
        return self._eng

    def close(self):
        """Release the library handle and its device workspace (also done when the object is deleted)."""
        if self._eng is not None:
            self._eng.close()
            self._eng = None

    def _mask(self):
        return nn_mask_of(self.non_negative, len(self._dims) + 1)

    def _sp(self):
        return float(self.softplus_kwargs['beta']), float(self.softplus_kwargs['threshold'])

    def _sharder(self):
        g = self._shard_group
        if g is None:
            return _engine.ShardedSum(enabled=False)
        return _engine.ShardedSum(group=None if g == 'world' else g, engine=self._engine() if self._dims_known() else None)

    def _dims_known(self):
        return hasattr(self, '_dims') and hasattr(self, 'n_classes')

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_eng'] = None
        st['_shard_group'] = None
        for k in ('theta', 'weights', 'y'):
            st[k] = st[k].detach().cpu()
        if isinstance(st['X'], torch.Tensor):
            st['X'] = st['X'].detach().cpu()
        st.pop('Bcp')
        return st

    def __setstate__(self, st):
        theta = st.pop('theta')
        self.__dict__.update(st)
        dev = self._torch_device()
        self.weights, self.y = self.weights.to(dev), self.y.to(dev)
        if not self.__dict__.get('_out_of_core', False):
            self.X = self.X.to(dev)
        want = self._dims + [self.n_classes]
        sizes, offs = _engine.factor_offsets(self._dims, self.rank, self.n_classes)
        self._set_theta([theta[offs[m]:offs[m + 1]].view(want[m], self.rank) for m in range(len(want))])

    def return_self(self):
        return self.Bcp

    def _fwd_grad(self, eng, cw, beta, thr, gs, gs_chunk=None, streamer=None):
        """Unnormalised local sums of one closure evaluation over all of this rank's samples."""
        if streamer is None:
            return eng.fwd_grad_mn(self.X, self.y, cw, self.theta, self.weights, self._mask(), beta, thr, gradsum=gs)
        gs.zero_()
        for lo, hi, xd in streamer.chunks():
            eng.fwd_grad_mn(xd, self.y[lo:hi], cw, self.theta, self.weights, self._mask(), beta, thr, gradsum=gs_chunk)
            gs.add_(gs_chunk)
        return gs

    def _streamer(self):
        if not self._out_of_core:
            return None
        return _engine.HostStreamer(self.X, torch.float32, self._torch_device(), chunk_samples=self._chunk_samples)

    def _class_weights(self, weights):
        # same call as the reference (mn:365,449): None raises
        return torch.as_tensor(weights, dtype=torch.float32).to(self._torch_device()).contiguous()

    def fit(self,
            lambda_L2=0.01,
            max_iter=1000,
            tol=1e-5,
            patience=10,
            weights=None,
            verbose=False,
            running_loss_logging_interval=10,
            LBFGS_kwargs=None):
        """mn:291-387 — L-BFGS on the flat parameter vector; closure = two-pass CUDA path."""
        if LBFGS_kwargs is None:
            raise TypeError('LBFGS_kwargs must be a dict of torch.optim.LBFGS keyword arguments (got None)')
        cw = self._class_weights(weights)
        eng = self._engine()
        beta, thr = self._sp()
        sharder = self._sharder()
        W = sharder.total(cw[self.y].sum().item(), self.theta.device)

        optimizer = _lbfgs.LBFGS(eng, self.theta, **LBFGS_kwargs)    # torch.optim.LBFGS's algorithm, device-resident
        gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=self.theta.device)
        gs_chunk = torch.empty_like(gs)
        streamer = self._streamer()

        def closure(grad_out, loss_out):
            self._fwd_grad(eng, cw, beta, thr, gs, gs_chunk, streamer)
            sharder.sum_(gs)
            eng.finish(gs, 1.0 / W, 1.0 / W, self.theta, lambda_L2, self._mask(), beta, thr, grad=grad_out,
                       loss=loss_out)

        def logged_loss():
            # extra forward, CE without the penalty (mn:371-372): one pass over X, then the
            # reference's own loss expression on the (N, C) probabilities
            if streamer is not None:
                s = self._fwd_grad(eng, cw, beta, thr, gs, gs_chunk, streamer)[-1].reshape(1).clone()
            else:
                P, _ = eng.forward_mn(self.X, self.theta, self.weights, self._mask(), beta, thr, want_pred=False)
                s = torch.nn.functional.cross_entropy(P.double(), self.y, weight=cw.double(), reduction='sum').reshape(1)
            return (sharder.sum_(s) / W).item()

        convergence_reached = False
        for ii in range(max_iter):
            if ii % running_loss_logging_interval == 0:
                self.loss_running.append(logged_loss())
                if verbose == 2:
                    print(f'Iteration: {ii}, Loss: {self.loss_running[-1]}')

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

    def fit_Adam(self,
                 lambda_L2=0.01,
                 max_iter=1000,
                 tol=1e-5,
                 patience=10,
                 weights=None,
                 verbose=False,
                 Adam_kwargs=None):
        """mn:389-471."""
        if Adam_kwargs is None:
            raise TypeError('Adam_kwargs must be a dict of torch.optim.Adam keyword arguments (got None)')
        hyper = _adam_hyper(Adam_kwargs)
        cw = self._class_weights(weights)
        eng = self._engine()
        beta, thr = self._sp()
        sharder = self._sharder()
        W = sharder.total(cw[self.y].sum().item(), self.theta.device)
        m = torch.zeros_like(self.theta)
        v = torch.zeros_like(self.theta)
        vmax = torch.zeros_like(self.theta) if hyper['amsgrad'] else None
        gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=self.theta.device)
        grad = torch.empty_like(self.theta)
        loss = torch.empty(2, dtype=torch.float64, device=self.theta.device)
        gs_chunk = torch.empty_like(gs)
        streamer = self._streamer()

        convergence_reached = False
        for ii in range(max_iter):
            self._fwd_grad(eng, cw, beta, thr, gs, gs_chunk, streamer)
            sharder.sum_(gs)
            eng.finish(gs, 1.0 / W, 1.0 / W, self.theta, lambda_L2, self._mask(), beta, thr, grad=grad, loss=loss)
            eng.adam_step(self.theta, grad, m, v, vmax, ii + 1, lr=hyper['lr'], betas=hyper['betas'],
                          eps=hyper['eps'], weight_decay=hyper['weight_decay'], lr_groups=self._adam_lr_groups(hyper))
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

    def _adam_lr_groups(self, hyper):
        """None = one parameter group (mn:447); the hierarchical subclass returns one rate per factor."""
        return None

    def predict(self, X=None, y_true=None, Bcp=None, device=None):
        """mn:474-545 — returns (probabilities (N,C) numpy, argmax labels (N,) numpy)."""
        eng = self._engine()
        beta, thr = self._sp()
        theta = self.theta if Bcp is None else _flatten(Bcp, None, torch.float32, self.theta.device)

        def fwd(xb):
            return eng.forward_mn(xb, theta, self.weights, self._mask(), beta, thr, want_pred=False)[0]

        if X is None:
            X = self.X
        if isinstance(X, torch.Tensor) and X.is_cuda:
            logit = fwd(X.to(device=self.theta.device, dtype=torch.float32)).cpu().numpy()
        else:
            logit = _predict_streamed(X, torch.float32, self.theta.device, fwd)
        pred = np.argmax(logit, axis=1)
        return logit, pred

    def return_Bcp_final(self):
        """mn:548-560."""
        Bcp = list(non_neg_fn(self.Bcp, self.non_negative, softplus_kwargs=self.softplus_kwargs))
        return [Bcp[ii].detach().cpu().numpy() for ii in range(len(Bcp))]

    def make_confusion_matrix(self, prob_or_pred='pred', prob=None, pred=None, y_true=None):
        """mn:562-597."""
        if (prob is None) and (pred is None):
            prob, pred = self.predict()
            cm, acc = self.make_confusion_matrix(prob_or_pred='pred', pred=pred, y_true=y_true)

        if y_true is None:
            y_true = self.y.detach().cpu().numpy()

        if prob_or_pred == 'pred':
            cm = confusion_matrix(pred, y_true)
        elif prob_or_pred == 'prob':
            cm = confusion_matrix(prob, y_true)

        acc = np.sum(np.diag(cm)) / np.sum(cm)
        return cm, acc

    def detach_Bcp(self):
        """mn:599-608."""
        return [Bcp.detach().cpu().numpy() for Bcp in self.Bcp]

    def get_params(self):
        """mn:610-624 (the reference reads a ``self.bias`` that the multinomial model never
        creates, mn:619; the key is kept with value None so the dict round-trips)."""
        return {'X': self.X.detach().cpu().numpy() if isinstance(self.X, torch.Tensor) else np.asarray(self.X),
                'y': self.y.detach().cpu().numpy(),
                'weights': self.weights.detach().cpu().numpy(),
                'Bcp': self.detach_Bcp(),
                'bias': None,
                'non_negative': self.non_negative,
                'softplus_kwargs': self.softplus_kwargs,
                'rank': self.rank,
                'device': self.device,
                'loss_running': self.loss_running}

    def set_params(self, params):
        """mn:626-643."""
        self.device = params['device']
        dev = self._torch_device()
        self.X = torch.as_tensor(params['X'], dtype=torch.float32).to(dev)
        self.y = torch.as_tensor(params['y'], dtype=torch.long).to(dev)
        self.weights = torch.as_tensor(params['weights'], dtype=torch.float32).to(dev)
        self.non_negative = params['non_negative']
        self.softplus_kwargs = params['softplus_kwargs']
        self.rank = params['rank']
        self.loss_running = params['loss_running']
        self._dims = [int(d) for d in self.X.shape[1:]]
        self._eng = None
        self._set_theta([torch.as_tensor(b) for b in params['Bcp']])

    def display_params(self):
        """mn:645-659."""
        print('X:', self.X.shape)
        print('y:', self.y.shape)
        print('weights:', self.weights)
        print('Bcp:', self.Bcp)
        print('non_negative:', self.non_negative)
        print('softplus_kwargs:', self.softplus_kwargs)
        print('rank:', self.rank)
        print('device:', self.device)
        print('loss_running:', self.loss_running)

    def plot_outputs(self):
        """mn:661-696."""
        import matplotlib.pyplot as plt
        plt.figure()
        plt.plot(self.loss_running)
        plt.xlabel('logged iteration')
        plt.ylabel('loss')
        plt.title('loss')

        logit, pred = self.predict()
        fig, axs = plt.subplots(2)
        axs[0].imshow(idx_to_oneHot(pred, self.n_classes), aspect='auto', interpolation='none')
        axs[1].imshow(idx_to_oneHot(self.y.detach().cpu().numpy(), self.n_classes), aspect='auto',
                      interpolation='none')
        axs[1].set_xlabel('class')
        fig.suptitle('predictions')

        cm, acc = self.make_confusion_matrix(prob_or_pred='pred')
        fig = plt.figure()
        plt.imshow(cm)
        plt.ylabel('true class')
        plt.xlabel('predicted class')
        plt.title('confusion matrix (predictions)')

        Bcp_final = self.return_Bcp_final()
        fig, axs = plt.subplots(len(Bcp_final))
        for ii, val in enumerate(Bcp_final):
            axs[ii].set_title(f'factor {ii}')
            axs[ii].plot(val)
        fig.suptitle('components')
