"""ctypes binding of librcgan_b200.so (the C ABI declared in include/rcgan_b200.h).

There is NO fallback: if the shared object is missing or a call fails, this raises.
PyTorch is used by the callers only for device memory and streams; every pointer handed
to the library is a raw `tensor.data_ptr()` and every launch goes to the caller's stream.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_long, c_size_t, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'librcgan_b200.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
HINGE_D_REAL, HINGE_D_FAKE, HINGE_G, CE_D_REAL, CE_D_FAKE, CE_G = 0, 1, 2, 3, 4, 5
MT_STATE_WORDS = 625
ABI_VERSION = 11        # == RCGAN_ABI_VERSION of include/rcgan_b200.h; bump both on every signature change


class RcganError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """rcgan_conv_desc"""
    _fields_ = [(n, c_int) for n in ('n', 'h', 'w', 'cin', 'ho', 'wo', 'cout', 'kh', 'kw', 'stride', 'pad_t', 'pad_l',
                                     'ldx', 'ldy', 'dtype')]


P = c_void_p
DP = POINTER(ConvDesc)


class ConvEpilogue(ctypes.Structure):
    """mirror of rcgan_conv_epilogue (include/rcgan_b200.h)"""
    _fields_ = [('mask', c_void_p), ('mask_act', c_int), ('mask_leak', c_float), ('res', c_void_p), ('res_up', c_int),
                ('ld_res', c_int), ('out2', c_void_p), ('out2_act', c_int), ('colstats', c_void_p)]


EP = POINTER(ConvEpilogue)
_SIGS = {
    'rcgan_last_error': (c_char_p, []),
    'rcgan_abi_version': (c_int, []),
    'rcgan_last_conv_variant': (c_char_p, []),
    'rcgan_conv_variant_log': (c_char_p, [c_int]),
    'rcgan_launch_count': (c_long, []),
    'rcgan_device_ok': (c_int, []),
    'rcgan_conv_wpack_bytes': (c_size_t, [DP]),
    'rcgan_conv_uses_tensor_cores': (c_int, [DP, c_int]),
    'rcgan_conv_wpack': (c_int, [DP, P, P, P, P]),
    'rcgan_conv_wpack_batched': (c_int, [c_int, P, P, P, P]),
    'rcgan_conv2d_fprop_ex': (c_int, [DP, P, P, P, P, c_int, c_int, c_float, EP, P]),
    'rcgan_conv2d_dgrad_ex': (c_int, [DP, P, P, P, P, c_int, c_int, c_float, c_int, EP, P]),
    'rcgan_conv2d_fprop': (c_int, [DP, P, P, P, P, P, c_int, c_int, c_float, P]),
    'rcgan_upconv2d_pack_bytes': (c_size_t, [DP]),
    'rcgan_upconv2d_fold': (c_int, [DP, P, P, P, P]),
    'rcgan_upconv2d_fprop': (c_int, [DP, P, P, P, P, c_int, c_int, c_float, P]),
    'rcgan_conv2d_fprop_res': (c_int, [DP, P, P, P, P, P, P, c_int, c_int, c_float, P]),
    'rcgan_conv2d_dgrad': (c_int, [DP, P, P, P, P, P, c_int, c_int, c_float, c_int, P]),
    'rcgan_conv2d_wgrad_workspace': (c_size_t, [DP]),
    'rcgan_conv2d_wgrad': (c_int, [DP, P, P, P, c_int, P, c_size_t, P]),
    'rcgan_im2col': (c_int, [DP, P, P, c_int, P]),
    'rcgan_wflip': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_act_bwd_colsum': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, c_int, P]),
    'rcgan_copy_batched': (c_int, [c_int, P, P, P, P]),
    'rcgan_wfold4': (c_int, [P, P, c_int, c_int, c_int, P]),
    'rcgan_wfold4_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    'rcgan_col2im': (c_int, [DP, P, c_int, P, P, c_int, c_int, c_float, c_int, P]),
    'rcgan_colsum': (c_int, [P, c_int, c_int, c_int, c_int, P, c_int, P]),
    'rcgan_bias_act_fwd': (c_int, [P, P, P, c_long, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    'rcgan_act_bwd': (c_int, [P, P, P, c_long, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    'rcgan_concat_label_fwd': (c_int, [P, c_int, P, P, c_int, c_long, c_int, c_int, c_int, c_int, P]),
    'rcgan_slice_bwd': (c_int, [P, c_int, P, c_int, c_long, c_int, c_int, c_int, P]),
    'rcgan_meanhw_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_meanhw_bwd': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_avgpool2_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_avgpool2_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_upsample2_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_upsample2_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'rcgan_add': (c_int, [P, P, P, c_long, c_int, P]),
    'rcgan_copy_acc': (c_int, [P, P, c_long, c_int, c_int, P]),
    'rcgan_cast': (c_int, [P, c_int, P, c_int, c_long, P]),
    'rcgan_bn_workspace': (c_size_t, [c_int, c_int, c_int]),
    'rcgan_bn_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, c_float, c_int, c_float, c_int, c_float, P, P,
                             P, P, c_size_t, P]),
    'rcgan_bn_bwd': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, c_int, c_float, P, P, c_int,
                             c_int, P, c_size_t, P, P]),
    'rcgan_bn_fwd_prestats': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, c_float, c_int, c_float, c_float, P, P, P, P, P]),
    'rcgan_colstats_floats': (c_size_t, [c_int]),
    'rcgan_bn_infer_bwd': (c_int, [P, P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, c_int, c_float, c_int, P, c_size_t, P]),
    'rcgan_bn_fwd_cat': (c_int, [P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_float, c_int, c_float, c_int, c_float,
                                 P, P, P, P, c_size_t, P]),
    'rcgan_bn_bwd_cat': (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, c_int, c_float, P, P, c_int,
                                 c_int, P, c_size_t, P, P]),
    'rcgan_recover_mse': (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P, P]),
    'rcgan_sgd': (c_int, [P, P, c_long, c_float, c_float, P]),
    'rcgan_sn_save_floats': (c_size_t, [c_int, c_int]),
    'rcgan_sn_workspace': (c_size_t, [c_int, c_int]),
    'rcgan_sn_fwd': (c_int, [P, P, c_int, c_int, P, P, P, P, c_size_t, P]),
    'rcgan_sn_bwd': (c_int, [P, P, P, c_int, c_int, P, P, c_int, P, c_size_t, P]),
    'rcgan_sn_workspace_batched': (c_size_t, [c_int, P, P]),
    'rcgan_sn_fwd_batched': (c_int, [c_int, P, P, P, P, P, P, P, P, c_size_t, P]),
    'rcgan_sn_bwd_batched': (c_int, [c_int, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    'rcgan_channel_loss': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P, c_int, P, P, P, P]),
    'rcgan_sigmoid_ce': (c_int, [P, P, c_long, c_float, P, P, P]),
    'rcgan_logit_loss': (c_int, [P, c_long, c_int, c_float, P, P, P]),
    'rcgan_softmax_rows_fwd': (c_int, [P, P, c_int, c_int, P]),
    'rcgan_softmax_rows_bwd': (c_int, [P, P, P, c_int, c_int, c_int, P]),
    'rcgan_gather_rows_fwd': (c_int, [P, P, P, c_int, c_int, P]),
    'rcgan_gather_rows_bwd': (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    'rcgan_adam_tf': (c_int, [P, P, P, P, c_long, c_float, P, c_float, c_float, c_float, c_float, POINTER(c_long),
                              POINTER(c_long), c_int, P]),
    'rcgan_zero': (c_int, [P, c_size_t, P]),
    'rcgan_sampler_table_host': (None, [POINTER(c_double), c_int, POINTER(c_double)]),
    'rcgan_mt_seed': (c_int, [P, c_uint32, P]),
    'rcgan_mt_shuffle_perm': (c_int, [P, P, c_int, P]),
    'rcgan_sample_labels_mnist': (c_int, [P, P, c_int, P, c_int, c_int, P, P, P, P]),
    'rcgan_sample_renoise_mnist': (c_int, [P, P, c_int, P, P, c_int, P, P, P]),
    'rcgan_sample_labels_cifar': (c_int, [P, P, c_int, P, c_int, P, P, P]),
    'rcgan_mt_uniform': (c_int, [P, P, c_long, c_double, c_double, P]),
    'rcgan_preprocess_cifar': (c_int, [P, P, P, c_int, c_int, P]),
    'rcgan_preprocess_cifar_u8': (c_int, [P, P, P, c_int, c_int, P]),
    'rcgan_random_fill': (c_int, [P, c_long, c_int, c_float, c_float, ctypes.c_ulonglong, P, c_uint32, P]),
}
EXPORTS = sorted(_SIGS)

_lib = None


def load():
    """Load the shared object (raises RcganError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RcganError('%s is missing: build it with `python -m robust_conditional_gan_b200.build` or '
                             '__graft_entry__.build(); there is no CPU or PyTorch fallback' % LIB_PATH)
        from . import build
        if build.sources_present() and build.needs_build() and os.path.exists(build.NVCC):
            build.build_library()             # sources changed since the library was built: rebuild, never call a stale ABI
        if build.sources_present() and build.needs_build():
            raise RcganError('%s is older than its sources under csrc/ or include/rcgan_b200.h: rebuild it with '
                             '`python -m robust_conditional_gan_b200.build` (a stale library would be called with the '
                             'wrong argument layout)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.rcgan_abi_version.restype = c_int
        got = lib.rcgan_abi_version()
        if got != ABI_VERSION:
            raise RcganError('%s has ABI version %d, this binding needs %d: rebuild the library' % (LIB_PATH, got, ABI_VERSION))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)       # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().rcgan_last_error().decode()


def last_conv_variant():
    """kernel variant the last conv entry point launched on this thread (test hook)"""
    return load().rcgan_last_conv_variant().decode()


def conv_variant_log(reset=True):
    """set of conv kernel variants launched since the last reset (test hook)"""
    return set(v for v in load().rcgan_conv_variant_log(1 if reset else 0).decode().split(';') if v)


def call(name, *args):
    """Invoke an int-returning entry point; raise with the library's message on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RcganError('%s failed (%d): %s' % (name, rc, last_error()))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()
