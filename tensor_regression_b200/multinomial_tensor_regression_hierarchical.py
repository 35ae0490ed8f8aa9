"""Drop-in for the reference's ``multinomial_tensor_regression_hierarchical.py``:
the multinomial model with an UNWEIGHTED cross-entropy (``torch.nn.CrossEntropyLoss()``,
hier:360,446 — ``fit`` / ``fit_Adam`` take no class ``weights``) and Adam run over three explicit
parameter groups ``Bcp[0], Bcp[1], Bcp[2]`` that all use ``Adam_kwargs['lr']`` (hier:436-440), which
ties the class to 3-D X (two feature modes + the class factor).  Reference lines: ``hier:<lines>``.

Same math as ``multinomial_tensor_regression`` with class weights of one, so the fit runs on the
same CUDA path (forward / gradient / finish / Adam kernels of libtrb200.so); the module-level
functions are the multinomial module's.
"""
import numpy as np

from . import multinomial_tensor_regression as _mn
from .multinomial_tensor_regression import (squeeze_integers, confusion_matrix, idx_to_oneHot, make_BcpInit,  # noqa: F401
                                            non_neg_fn, model, L2_penalty)


class CP_logistic_regression(_mn.CP_logistic_regression):
    """hier:211-699.  Constructor, ``return_Bcp_final``, ``make_confusion_matrix``, ``detach_Bcp``,
    ``get/set/display_params`` and ``plot_outputs`` are the multinomial class's."""

    def _ones(self):
        return np.ones(self.n_classes, dtype=np.float32)

    def fit(self,
            lambda_L2=0.01,
            max_iter=1000,
            tol=1e-5,
            patience=10,
            verbose=False,
            running_loss_logging_interval=10,
            LBFGS_kwargs=None):
        """hier:291-383 — L-BFGS, unweighted CE on the probabilities (softmax twice) + penalty."""
        return super().fit(lambda_L2=lambda_L2, max_iter=max_iter, tol=tol, patience=patience, weights=self._ones(),
                           verbose=verbose, running_loss_logging_interval=running_loss_logging_interval,
                           LBFGS_kwargs=LBFGS_kwargs)

    def fit_Adam(self,
                 lambda_L2=0.01,
                 max_iter=1000,
                 tol=1e-5,
                 patience=10,
                 verbose=False,
                 Adam_kwargs=None):
        """hier:385-470.  The reference builds three parameter groups from ``Bcp[0..2]`` (hier:436-440): a
        model with any other number of factors fails there with an IndexError (fewer) or silently
        leaves factors untrained (more); here it is an explicit error either way."""
        if Adam_kwargs is None:
            raise TypeError('Adam_kwargs must be a dict of torch.optim.Adam keyword arguments (got None)')
        if 'lr' not in Adam_kwargs:
            raise KeyError('lr')                       # hier:437 indexes Adam_kwargs['lr']
        if len(self.Bcp) != 3:
            raise IndexError('the hierarchical class optimises exactly Bcp[0], Bcp[1], Bcp[2] (3-D X): '
                             f'got {len(self.Bcp)} factors')
        return super().fit_Adam(lambda_L2=lambda_L2, max_iter=max_iter, tol=tol, patience=patience,
                                weights=self._ones(), verbose=verbose, Adam_kwargs=Adam_kwargs)

    def _adam_lr_groups(self, hyper):
        """hier:436-440 — three explicit parameter groups {'params': Bcp[i], 'lr': Adam_kwargs['lr']}: one step
        size per factor through tr_adam_step_groups (a ``lr_groups`` list of three rates in Adam_kwargs-style
        callers can be set on the instance as ``self.lr_groups`` to give the factors different rates)."""
        g = getattr(self, 'lr_groups', None)
        return [float(x) for x in g] if g is not None else [float(hyper['lr'])] * 3

    def predict(self, X=None, y_true=None, Bcp=None, device=None, plot_pref=False):
        """hier:473-544 — (probabilities, argmax labels); ``plot_pref`` is accepted and unused, as in the
        reference (its plotting block is commented out)."""
        return super().predict(X=X, y_true=y_true, Bcp=Bcp, device=device)
