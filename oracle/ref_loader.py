"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used by oracle/make_golden.py and by tests that pin the restatement against the
reference itself.  /root/reference does not exist on the GPU box, so nothing that
runs there may depend on this succeeding.  Test infrastructure only.
"""
import importlib.util
import os

from . import tensorly_standin

REFERENCE_DIR = os.environ.get('TR_REFERENCE_DIR', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'standard_tensor_regression.py'))


def _load(name):
    tensorly_standin.install()
    path = os.path.join(REFERENCE_DIR, name + '.py')
    spec = importlib.util.spec_from_file_location('_tr_reference_' + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def standard():
    if 'std' not in _cache:
        _cache['std'] = _load('standard_tensor_regression')
    return _cache['std']


def multinomial():
    if 'mn' not in _cache:
        _cache['mn'] = _load('multinomial_tensor_regression')
    return _cache['mn']


def hierarchical():
    if 'hier' not in _cache:
        _cache['hier'] = _load('multinomial_tensor_regression_hierarchical')
    return _cache['hier']


def spectral():
    if 'spec' not in _cache:
        _cache['spec'] = _load('spectral_tensor_regression')
    return _cache['spec']
