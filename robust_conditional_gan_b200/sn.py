"""Drop-in for mnist/sn.py (== cifar10/common/ops/sn.py): spectral_normed_weight."""
import torch

from . import scope as S
from .nnops import SpectralNormOp

NO_OPS = 'NO_OPS'


def _trunc_normal(shape, std=1.0):
    t = torch.empty(tuple(shape), dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=S.init_generator())
    return t


def spectral_normed_weight(W, u=None, num_iters=1, update_collection=None, with_sigma=False, reuse=False):
    """mnist/sn.py:17-75.  Returns W_bar (a graph tensor whose gradient flows through the power
    iteration).  update_collection None: u <- u' after every step this program runs;
    NO_OPS: u is read but not updated.  num_iters must be 1 (the only value the reference uses)."""
    assert num_iters == 1, 'only one power iteration is implemented (the reference never uses more)'
    assert not with_sigma, 'with_sigma is not used on the training path'
    with S.variable_scope('spectral_norm'):
        if u is None:
            u = S.get_variable('u', [1, W.shape[-1]], _trunc_normal, trainable=False)
    # one power iteration per weight per program run: every call site of a run computes the same (W, u) -> W_bar
    # (the reference's repeated assigns are idempotent), so the op is shared.
    from .graph import cur
    prog = cur()
    cache = prog.__dict__.setdefault('sn_cache', {})
    op = cache.get(id(W))
    if op is None:
        op = cache[id(W)] = SpectralNormOp(W, u, update=False)
    if update_collection != NO_OPS and not op.updates_u:
        op.updates_u = True
        prog.add_update(u, op.u_new)
    return op.wbar
