"""tf.variable_scope / tf.get_variable equivalents over graph.VariableStore."""
import contextlib

import torch

_store = None
_gen = None


def set_store(store, seed=0):
    global _store, _gen
    _store = store
    _gen = torch.Generator().manual_seed(seed)


def store():
    assert _store is not None, 'no VariableStore active'
    return _store


def init_generator():
    return _gen


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    s = store()
    s.scope.append(name)
    try:
        yield
    finally:
        s.scope.pop()


def get_variable(name, shape, initializer, trainable=True):
    return store().get(name, tuple(int(x) for x in shape), initializer, trainable)
