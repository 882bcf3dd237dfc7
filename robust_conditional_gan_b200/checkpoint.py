"""Checkpoint save / restore under the reference's TF variable names (mnist/model.py:842-867,
cifar10/gan_resnet.py:906-914, 1007-1014).

The reference writes TF-1 `Saver` checkpoints: `<dir>/<prefix>-<step>.{index,meta,data-*}` plus a `checkpoint` text file
naming the latest one.  TensorFlow's table format cannot be produced without TensorFlow, so the container here is one `.npz`
per checkpoint with EXACTLY the Saver's contents and key names: every global variable under its TF name (trainable weights,
spectral-norm `u`, batch-norm moving statistics, `confusion_logits`) and the optimizer slots TF-Adam keeps per variable
(`<var>/Adam`, `<var>/Adam_1`) with `beta1_power` / `beta2_power` expressed as the step count.  A converter from a TF
checkpoint reader's {name: array} dict is therefore the identity (load_state(model, dict)).  The `checkpoint` index file has
the Saver's syntax (`model_checkpoint_path: "..."`) and `max_to_keep` pruning (gan_resnet.py:907 keeps 5)."""
import os
import re

import numpy as np
import torch


def state_of(model):
    """{TF name: numpy array} of everything a tf.train.Saver() would write for this model."""
    out = {n: v.data.detach().reshape(v.shape).cpu().numpy() for n, v in model.store.vars.items()}
    for k, g in model.store.groups.items():
        m, v_ = g.m.detach().cpu().numpy(), g.v.detach().cpu().numpy()
        for var in g.vars:
            sl = slice(var.offset, var.offset + var.numel())
            out[var.name + '/Adam'] = m[sl].reshape(var.shape)
            out[var.name + '/Adam_1'] = v_[sl].reshape(var.shape)
        out['__adam_steps__/' + k] = np.asarray(g.t, dtype=np.int64)       # beta1_power = beta1 ** t, beta2_power = beta2 ** t
    return out


def load_state(model, state, strict=True):
    """Inverse of state_of.  strict: every variable of the model must be present (optimizer slots are optional, as when a
    checkpoint converted from inference-only weights is restored)."""
    missing = [n for n in model.store.vars if n not in state]
    if strict and missing:
        raise KeyError('checkpoint lacks variables: %s' % missing[:8])
    model.store.load_state_dict({n: torch.as_tensor(np.asarray(state[n])) for n in model.store.vars if n in state})
    for k, g in model.store.groups.items():
        for var in g.vars:
            sl = slice(var.offset, var.offset + var.numel())
            if var.name + '/Adam' in state:
                g.m[sl].copy_(torch.as_tensor(np.asarray(state[var.name + '/Adam'])).reshape(-1))
                g.v[sl].copy_(torch.as_tensor(np.asarray(state[var.name + '/Adam_1'])).reshape(-1))
        if '__adam_steps__/' + k in state:
            g.t = int(state['__adam_steps__/' + k])
    return missing


def _index_path(ckpt_dir):
    return os.path.join(ckpt_dir, 'checkpoint')


def all_checkpoints(ckpt_dir):
    p = _index_path(ckpt_dir)
    if not os.path.exists(p):
        return []
    return re.findall(r'^all_model_checkpoint_paths: "(.*)"$', open(p).read(), flags=re.M)


def latest_checkpoint(ckpt_dir):
    """tf.train.latest_checkpoint / get_checkpoint_state(...).model_checkpoint_path: path prefix or None."""
    p = _index_path(ckpt_dir)
    if not os.path.exists(p):
        return None
    m = re.search(r'^model_checkpoint_path: "(.*)"$', open(p).read(), flags=re.M)
    if not m:
        return None
    path = m.group(1)
    return path if os.path.isabs(path) else os.path.join(ckpt_dir, path)


def save(model, ckpt_dir, prefix, global_step, max_to_keep=5, extra=None):
    """saver.save(sess, os.path.join(ckpt_dir, prefix), global_step=step) -> '<ckpt_dir>/<prefix>-<step>' (+ '.npz')."""
    os.makedirs(ckpt_dir, exist_ok=True)
    name = '%s-%d' % (prefix, int(global_step))
    state = state_of(model)
    for k, v in (extra or {}).items():
        state['__extra__/' + k] = np.asarray(v)
    tmp = os.path.join(ckpt_dir, name + '.tmp.npz')
    np.savez(tmp, **state)
    os.replace(tmp, os.path.join(ckpt_dir, name + '.npz'))
    kept = [c for c in all_checkpoints(ckpt_dir) if c != name] + [name]
    while max_to_keep and len(kept) > max_to_keep:
        old = kept.pop(0)
        try:
            os.remove(os.path.join(ckpt_dir, old + '.npz'))
        except OSError:
            pass
    with open(_index_path(ckpt_dir), 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % name)
        for c in kept:
            f.write('all_model_checkpoint_paths: "%s"\n' % c)
    return os.path.join(ckpt_dir, name)


def restore(model, path_prefix, strict=True):
    """saver.restore(sess, path_prefix).  Returns the dict of extras stored with the checkpoint."""
    with np.load(path_prefix + '.npz') as z:
        state = {k: z[k] for k in z.files}
    load_state(model, state, strict=strict)
    return {k[len('__extra__/'):]: v for k, v in state.items() if k.startswith('__extra__/')}


def step_of(path_prefix):
    """the reference's `int(next(re.finditer("(\\d+)(?!.*\\d)", ckpt_name)).group(0))` (mnist/model.py:863)"""
    return int(next(re.finditer(r"(\d+)(?!.*\d)", os.path.basename(path_prefix))).group(0))
