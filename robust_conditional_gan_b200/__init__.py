"""rcgan-b200: B200-native RCGAN training hot path behind the reference's op / model / flag surface.

Layout
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/rcgan_b200.h) -> librcgan_b200.so
  _C.py        ctypes binding (no fallback: a missing library raises)
  graph.py     static-program runtime (the sess.run equivalent), nnops.py op classes
  ops.py, sn.py, model.py         drop-ins for mnist/{ops,sn,model}.py
  cifar/       drop-ins for cifar10/gan_resnet.py and cifar10/common/ops/*
  sampler.py   label-noise sampler host side
"""
