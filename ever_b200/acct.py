"""Algorithmic work of a step, counted from the C-ABI calls the engine actually makes (bench.py's ``hbm_bytes_per_step`` and
``flops_per_step``): a proxy around the ctypes library that, per call, adds the bytes the op must move when every operand
is read once and every result written once (bf16 activations, fp32 parameters / gradients), and the convolution FLOPs
(2 * N * Ho * Wo * Cout * Cin * k^2 per direction).  Only installed while accounting is on
(``FarSegEngine.accounting()``): the product path calls the library directly."""


def _v(a):
    return getattr(a, 'value', a) or 0


def _conv_fwd(a):
    n, h, w, cin, cop, k, s, cout = _v(a[1]), _v(a[2]), _v(a[3]), _v(a[4]), _v(a[6]), _v(a[7]), _v(a[8]), _v(a[10])
    ho, wo = h // s, w // s
    b = 2 * (n * h * w * cin + n * ho * wo * cout + k * k * cop * cin)
    return b, 2 * n * ho * wo * cout * cin * k * k


def _conv_fwd_add(a):   # evb_conv2d_fwd: optional fused add operand (mode 2 = nearest-x2 of the coarser map)
    b, f = _conv_fwd(a)
    if _v(a[12]):
        n, h, w, s, cout = _v(a[1]), _v(a[2]), _v(a[3]), _v(a[8]), _v(a[10])
        b += 2 * n * (h // s) * (w // s) * cout // (4 if _v(a[13]) == 2 else 1)
    return b, f


def _conv_dgrad(a):
    n, ho, wo, cout, cip, k, h, w, cin, acc = (_v(a[1]), _v(a[2]), _v(a[3]), _v(a[4]), _v(a[6]), _v(a[7]), _v(a[10]),
                                               _v(a[11]), _v(a[12]), _v(a[13]))
    b = 2 * (n * ho * wo * cout + (2 if acc else 1) * n * h * w * cin + k * k * cip * cout)
    return b, 2 * n * ho * wo * cout * cin * k * k


def _conv_wgrad(a):
    n, h, w, cin, cout, k, s = _v(a[1]), _v(a[2]), _v(a[3]), _v(a[4]), _v(a[6]), _v(a[7]), _v(a[8])
    ho, wo = h // s, w // s
    return 2 * (n * h * w * cin + n * ho * wo * cout) + 4 * cout * cin * k * k, 2 * n * ho * wo * cout * cin * k * k


def _bn_apply(a):
    m, c = _v(a[5]), _v(a[6])
    return 2 * m * c * (3 if _v(a[3]) else 2), 0


def _bn_bwd(a):
    m, c = _v(a[15]), _v(a[16])
    ymask = (0.5 if _v(a[7]) == 3 else 2) if _v(a[2]) else 0     # mask_mode 3: uint32 per 8 channels = a quarter pass, twice
    passes = 4 + 1 + ymask + ((2 if _v(a[11]) else 1) if _v(a[10]) else 0)
    return int(2 * m * c * passes), 0


def _rows_c(mi, ci, passes):
    return lambda a: (2 * _v(a[mi]) * _v(a[ci]) * passes, 0)


def _maxpool_fwd(a):
    n, h, w, c = _v(a[3]), _v(a[4]), _v(a[5]), _v(a[6])
    return n * h * w * c * 2 + n * (h // 2) * (w // 2) * c * 3, 0


def _maxpool_bwd(a):
    n, h, w, c = _v(a[3]), _v(a[4]), _v(a[5]), _v(a[6])
    return n * h * w * c * 2 + n * (h // 2) * (w // 2) * c * 3, 0


def _bilinear(a):   # (src, .., .., dst, n, h, w, c, ldin, ldout, f)
    n, h, w, c, f = _v(a[4]), _v(a[5]), _v(a[6]), _v(a[7]), _v(a[10])
    return 2 * n * h * w * c * (1 + f * f), 0


def _bilinear_bwd(a):   # (dy, dlow, n, h, w, c, ldin, ldout, f, ws, cap)
    n, h, w, c, f = _v(a[2]), _v(a[3]), _v(a[4]), _v(a[5]), _v(a[8])
    return 2 * n * h * w * c * (1 + f * f), 0


def _im2col(kp_i, elem):
    def f(a):
        o = 2 if elem == 1 else 0     # evb_im2col_u8 has two extra pointer arguments (mean, std)
        n, cin, h, w, kp = _v(a[2 + o]), _v(a[3 + o]), _v(a[4 + o]), _v(a[5 + o]), _v(a[6 + o])
        s = _v(a[8 + o]) or 2
        return n * cin * h * w * elem + 2 * n * (h // s) * (w // s) * kp, 0
    return f


TABLE = {
    'evb_conv2d_fwd': _conv_fwd_add, 'evb_conv2d_fwd_stats': _conv_fwd, 'evb_conv2d_fwd_bias_stats': _conv_fwd,
    'evb_conv2d_dgrad': _conv_dgrad, 'evb_conv2d_wgrad': _conv_wgrad,
    'evb_bn_apply': _bn_apply,
    'evb_bn_apply_mask': lambda a: (int(2 * _v(a[6]) * _v(a[7]) * 3.25), 0), 'evb_bn_bwd': _bn_bwd, 'evb_bn_stats': _rows_c(1, 2, 1),
    'evb_maxpool3x3s2_fwd': _maxpool_fwd, 'evb_maxpool3x3s2_bwd': _maxpool_bwd,
    'evb_bilinear_up': _bilinear, 'evb_bilinear_up_bwd_sep': _bilinear_bwd,
    'evb_relation_fwd': _rows_c(9, 11, 3), 'evb_relation_bwd': _rows_c(12, 14, 5),
    'evb_relation_fwd_v2': _rows_c(10, 12, 3), 'evb_relation_bwd_v2': _rows_c(13, 15, 5),
    'evb_copy2d_bf16': lambda a: (4 * _v(a[4]) * _v(a[5]), 0),
    'evb_channel_scale': lambda a: (4 * _v(a[3]) * _v(a[4]) * _v(a[5]), 0),
    'evb_merge4': lambda a: (2 * 5 * _v(a[5]), 0),
    'evb_scale_add': lambda a: (2 * _v(a[4]) * (3 if _v(a[2]) else 2), 0),
    'evb_sumpool2': lambda a: (2 * _v(a[2]) * _v(a[3]) * _v(a[4]) * _v(a[5]) * (6 if _v(a[6]) else 5), 0),
    'evb_im2col_nchw': _im2col(6, 4), 'evb_im2col_u8': _im2col(8, 1),
    'evb_loss_stats': lambda a: (_v(a[2]) * (2 * _v(a[4]) + 8), 0),
    'evb_loss_grad': lambda a: (_v(a[2]) * (4 * _v(a[4]) + 8), 0),
    'evb_bias_grad': _rows_c(1, 2, 1),
    'evb_gap_fwd': lambda a: (2 * _v(a[2]) * _v(a[3]) * _v(a[4]), 0),
    'evb_gap_bwd': lambda a: (2 * _v(a[2]) * _v(a[3]) * _v(a[4]), 0),
    'evb_grad_norm': lambda a: (4 * _v(a[1]), 0),
    'evb_sgd_step': lambda a: (24 * _v(a[3]), 0), 'evb_sgd_step_masked': lambda a: (25 * _v(a[3]), 0),
}


class AcctLib:
    def __init__(self, lib):
        self._lib = lib
        self.bytes, self.flops, self.calls = {}, {}, {}

    def add(self, name, nbytes, nflops=0):
        self.bytes[name] = self.bytes.get(name, 0) + int(nbytes)
        self.flops[name] = self.flops.get(name, 0) + int(nflops)
        self.calls[name] = self.calls.get(name, 0) + 1

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        rule = TABLE.get(name)
        if rule is None:
            return fn

        def call(*a):
            b, f = rule(a)
            self.add(name, b, f)
            return fn(*a)
        return call

    def totals(self):
        return dict(bytes=sum(self.bytes.values()), flops=sum(self.flops.values()),
                    by_call={k: dict(bytes=self.bytes[k], flops=self.flops[k], calls=self.calls[k]) for k in self.bytes})
