"""CPU oracle for the RCGAN training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`robust_conditional_gan_b200/`) imports, links or executes anything in this
directory; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may.

What it is: a restatement, in PyTorch-CPU (fp64 / fp32, autograd) and numpy / C,
of the arithmetic the reference (tkkiran/Robust-Conditional-GAN, TensorFlow 1.5)
asks TensorFlow and numpy for on the hot path.  Every
function cites the reference file:line it follows.

Pinning status
--------------
* label-noise sampler (`oracle/sampler.py`, `oracle/sampler_c.c`): PINNED -- it is
  checked bit-exactly against numpy's frozen legacy `np.random.*` stream, which is
  the very dependency the reference calls (mnist/model.py:795-834,
  cifar10/common/data/cifar10.py:29-38).  Golden vectors: tests/golden/sampler_*.npz.
* every floating-point op (conv, deconv, linear, BN, spectral norm, losses, Adam):
  **parity unpinned** -- the reference ships no tests, no golden vectors, and its
  arithmetic lives in TensorFlow 1.5, which is not installable here (no network,
  Python 3.12).  The restatement is anchored on the reference's call sites and on
  algebraic identities checked in tests/test_oracle_*.py (TF SAME padding,
  conv2d_transpose == dgrad of conv2d, closed-form SN gradient == autograd through
  the power iteration, ...).
"""
