"""B200-native CP (Kruskal) tensor regression — drop-in for the fit / predict path of
kimerein/tensor_regression's ``standard_tensor_regression``, ``multinomial_tensor_regression`` and
``multinomial_tensor_regression_hierarchical``.

    from tensor_regression_b200 import standard_tensor_regression as STR
    from tensor_regression_b200 import multinomial_tensor_regression as MTR

The compute path is hand-written sm_100a CUDA behind a C ABI (include/tr_b200.h, built
in-tree as ``libtrb200.so``); there is no CPU fallback — without the library or without a
CUDA device every entry point raises.
"""
from . import _lib  # noqa: F401  (fails loudly if the extension is missing)
from . import engine  # noqa: F401
from . import standard_tensor_regression  # noqa: F401
from . import multinomial_tensor_regression  # noqa: F401
from . import multinomial_tensor_regression_hierarchical  # noqa: F401

__all__ = ['standard_tensor_regression', 'multinomial_tensor_regression',
           'multinomial_tensor_regression_hierarchical', 'engine']
