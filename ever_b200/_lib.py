"""ctypes binding of libevb200.so.  There is NO fallback: if the library is missing or a call fails the
product path raises (the oracle under oracle/ is never used here)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EVB_LIB') or os.path.join(_HERE, 'lib', 'libevb200.so')   # EVB_LIB: A/B builds (tools/)
_lib = None


class EvbError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EvbError('libevb200.so not built (%s): run `python -m ever_b200.build` / __graft_entry__.build(); '
                           'there is no CPU or PyTorch fallback' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.evb_last_cuda_error.restype = ctypes.c_char_p
        for fn in ('evb_conv2d_wgrad_workspace', 'evb_bn_workspace', 'evb_loss_workspace', 'evb_sgd_workspace',
                   'evb_relation_bwd_workspace', 'evb_bilinear_up_bwd_workspace'):
            getattr(_lib, fn).restype = ctypes.c_longlong
    return _lib


# kernels launched per C-ABI call (lower bounds; evb_conv2d_dgrad stride 2 launches up to 4)
_KERNELS = {'evb_zero_bytes': 0, 'evb_conv2d_wgrad': 2, 'evb_conv2d_wgrad(stem)': 2, 'evb_bn_stats': 2, 'evb_bn_bwd': 3, 'evb_bias_grad': 2,
            'evb_loss_stats': 2, 'evb_linear_bwd': 2, 'evb_grad_norm': 2, 'evb_relation_bwd': 2, 'evb_bilinear_up_bwd_sep': 2}
launches = [0]


def check(rc, what):
    launches[0] += _KERNELS.get(what, 1)
    if rc != 0:
        raise EvbError('%s failed: rc=%d cuda=%s' % (what, rc, lib().evb_last_cuda_error().decode()))


def ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
