"""FreeNet (patch-free hyperspectral classification, Z-Zheng/FreeNet; SURVEY.md row a14, BASELINE configs[4]) on the B200
engine: ``FreeNetB200`` registered in ``ever.registry.MODEL`` as 'FreeNetB200' (and 'FreeNet').

FreeNet is not in the reference tree (README.md:55 links it); what IS there are its building blocks -- ``SEBlock``
(ever/module/se_block.py:9-24), ``nn.GroupNorm``, the nearest top-down add (ever/module/fpn.py:96-105) and the ERModule
plugin surface -- so the parameter containers below use exactly those layouts (``SEBlock.seq.0/2``, GroupNorm affine) and the
network is restated from the published description (oracle/freenet_oracle.py, parity unpinned as a whole).

Arithmetic: the 3x3 / 1x1 convolutions (+bias, + nearest-x2 add in the epilogue) run on the tcgen05 implicit-GEMM kernel with
channel counts zero-padded to its 64-channel granularity (200 -> 256, 96 -> 128); GroupNorm = per-channel sums
(evb_bn_stats) -> per-group fold (evb_gn_fold) -> the BatchNorm apply kernel; its backward = evb_norm_bwd_reduce ->
evb_gn_bwd_consts -> evb_norm_bwd_apply; squeeze-excitation = GAP + two tiny linears + sigmoid + a channel scale.
"""
import ctypes

import torch
import torch.nn as nn

from ._ever_api import MODEL, ERModule
from .module import NativeStepMixin
from ._lib import check, ptr, stream
from .engine import BF16, Act, ConvP, FarSegEngine, _ceil, c_float, c_int, c_ll


# ------------------------------------------------------------------------------------------------ parameter containers
class _SEBlock(nn.Module):
    """attribute layout of ever.module.se_block.SEBlock (se_block.py:9-24)"""

    def __init__(self, in_channels, reduction):
        super().__init__()
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.seq = nn.Sequential(nn.Linear(in_channels, in_channels // reduction), nn.ReLU(inplace=True),
                                 nn.Linear(in_channels // reduction, in_channels), nn.Sigmoid())


def _conv3x3_gn_relu(cin, cout, groups):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1), nn.GroupNorm(groups, cout), nn.ReLU(inplace=True))


def _downsample2x(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1), nn.ReLU(inplace=True))


def _repeat_block(c, r, n):
    return nn.Sequential(*[nn.Sequential(_SEBlock(c, r), _conv3x3_gn_relu(c, c, r)) for _ in range(n)])


class _GNP:
    """GroupNorm parameters, zero-padded to the activation's channel count (padded channels: gamma 1, beta 0)"""

    def __init__(self, gn, c, dev):
        self.gn, self.c_real, self.c, self.groups = gn, gn.num_channels, c, gn.num_groups
        self.padded = c != gn.num_channels
        self.gamma_p = torch.ones(c, device=dev)
        self.beta_p = torch.zeros(c, device=dev)
        self.dgamma_s = torch.zeros(c, device=dev)   # this sample's per-channel sums (also the source of the constants)
        self.dbeta_s = torch.zeros(c, device=dev)
        self._gw = False
        self.name = None


class FreeNetEngine(FarSegEngine):
    def _read_config(self, module):
        self.ignore_index = 255
        self.ce_w, self.dice_w, self.smooth, self.sync_dice = 1.0, 0.0, 1.0, False
        self.K = int(module.config.num_classes)
        if self.K > 64:
            raise NotImplementedError('more than 64 classes')
        self.ld = 16 if self.K <= 16 else 32 if self.K <= 32 else 64

    def _collect(self):
        m = self.m
        self.kind, self.fs_v2, self.scene_shared = 'freenet', False, False
        self.convs, self.bns, self.bns_padded, self.gns = [], [], [], []
        self.project, self.drop_p = None, 0.0

        def C(conv, need_dgrad=True):
            co, ci = conv.weight.shape[:2]
            cp = ConvP(conv.weight, conv.bias, conv.stride[0], cout_pad=_ceil(co, 64), cin_pad=_ceil(ci, 64))
            cp.need_dgrad = need_dgrad
            self.convs.append(cp)
            return cp

        def G(gn, c):
            gp = _GNP(gn, c, self.dev)
            self.gns.append(gp)
            return gp
        ops, first = [], True
        for op in m.feature_ops:
            if isinstance(op, nn.Identity):
                ops.append(('feat',))
            elif isinstance(op[0], nn.Conv2d) and len(op) == 3:          # conv3x3_gn_relu
                cp = C(op[0], need_dgrad=not first)
                ops.append(('cgr', cp, G(op[1], cp.cop), op[2]))
            elif isinstance(op[0], nn.Conv2d):                           # downsample2x
                ops.append(('down', C(op[0]), op[1]))
            else:                                                        # repeat_block
                for blk in op:
                    ops.append(('se', blk[0]))
                    cp = C(blk[1][0])
                    ops.append(('cgr', cp, G(blk[1][1], cp.cop), blk[1][2]))
            first = False
        self.ops = ops
        self.reduce = [C(c) for c in m.reduce_1x1convs]
        self.fuse = [C(c) for c in m.fuse_3x3convs]
        cls = m.cls_pred_conv
        self.cls = ConvP(cls.weight, cls.bias, 1, cout_pad=64, cin_pad=_ceil(cls.weight.shape[1], 64))
        self.convs.append(self.cls)
        self.cls_scale = 1
        mods = {id(mod): path for path, mod in m.named_modules()}
        wname = {id(mod.weight): path for path, mod in m.named_modules() if isinstance(getattr(mod, 'weight', None), nn.Parameter)}
        for cp in self.convs:
            cp.name = wname.get(id(cp.weight))
        for gp in self.gns:
            gp.name = wname.get(id(gp.gn.weight))
        self._modname = mods
        self._alloc_packs()

    # ------------------------------------------------------------------ ops
    def _refresh_gn(self):
        st = stream()
        for gp in self.gns:
            gp._gw = False
            for src, dst in ((gp.gn.weight, gp.gamma_p), (gp.gn.bias, gp.beta_p)):
                check(self.L.evb_copy2d_f32(ptr(src), c_int(gp.c_real), ptr(dst), c_int(gp.c), c_int(1), c_int(gp.c_real),
                                            c_int(0), st), 'evb_copy2d_f32')

    def gn_act(self, x, gp, name=None, train=True):
        """y = relu(GroupNorm(x)) per sample (nn.GroupNorm normalises over (C / G, H, W) of each sample)"""
        L = self.L
        n, h, w, c = x.data.shape
        m_rows = h * w
        y = Act(self._new(n, h, w, c))
        folds = []
        for s in range(n):
            st8 = self._new(8, c, dtype=torch.float32)   # mean_c, rstd_c, bn scale, bn shift | scale, shift, gmean, grstd
            ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(c)))
            check(L.evb_bn_stats(ptr(x.data[s]), c_ll(m_rows), c_int(c), ptr(gp.gamma_p), ptr(gp.beta_p), None, None,
                                 c_float(0.0), c_float(1e-5), ptr(st8[0]), ptr(st8[1]), ptr(st8[2]), ptr(st8[3]), ptr(ws),
                                 stream()), 'evb_bn_stats')
            check(L.evb_gn_fold(ptr(st8[0]), ptr(st8[1]), c_float(1e-5), ptr(gp.gamma_p), ptr(gp.beta_p), c_int(c),
                                c_int(gp.c_real), c_int(gp.groups), c_float(gp.gn.eps), ptr(st8[4]), ptr(st8[5]), ptr(st8[6]),
                                ptr(st8[7]), stream()), 'evb_gn_fold')
            check(L.evb_bn_apply(ptr(x.data[s]), ptr(st8[4]), ptr(st8[5]), None, ptr(y.data[s]), c_ll(m_rows), c_int(c),
                                 c_int(1), stream()), 'evb_bn_apply')
            folds.append(st8)
        self._tf_fwd(y, name)
        if train:
            def bwd():
                if y.grad is None:
                    return
                self._tf_bwd(y)
                gx, acc_x = self._grad_into(x)
                assert not acc_x
                inv_m = 1.0 / (m_rows * (gp.c_real // gp.groups))
                for s in range(n):
                    st8 = folds[s]
                    ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(c)))
                    check(L.evb_norm_bwd_reduce(ptr(y.grad[s]), ptr(x.data[s]), None, ptr(st8[6]), ptr(st8[7]), ptr(st8[4]),
                                                ptr(st8[5]), c_int(2), ptr(gp.dgamma_s), ptr(gp.dbeta_s), c_int(0),
                                                c_ll(m_rows), c_int(c), ptr(ws), stream()), 'evb_norm_bwd_reduce')
                    k = self._new(2, c, dtype=torch.float32)
                    check(L.evb_gn_bwd_consts(ptr(gp.dgamma_s), ptr(gp.dbeta_s), ptr(gp.gamma_p), ptr(st8[6]), ptr(st8[7]),
                                              c_int(c), c_int(gp.c_real), c_int(gp.groups), c_float(inv_m), ptr(k[0]),
                                              ptr(k[1]), stream()), 'evb_gn_bwd_consts')
                    check(L.evb_norm_bwd_apply(ptr(y.grad[s]), ptr(x.data[s]), None, ptr(st8[4]), ptr(st8[5]), ptr(k[0]),
                                               ptr(k[1]), c_int(2), ptr(gx[s]), None, c_int(0), c_ll(m_rows), c_int(c),
                                               stream()), 'evb_norm_bwd_apply')
                    acc = 1 if (self.accumulate or gp._gw) else 0
                    for src, prm in ((gp.dgamma_s, gp.gn.weight), (gp.dbeta_s, gp.gn.bias)):
                        if self._g(prm) is not None:
                            check(L.evb_copy2d_f32(ptr(src), c_int(c), ptr(self._g(prm)), c_int(gp.c_real), c_int(1),
                                                   c_int(gp.c_real), c_int(acc), stream()), 'evb_copy2d_f32')
                    gp._gw = True
            self.tape.append(bwd)
        return y

    def relu(self, x, name=None, train=True):
        L = self.L
        n, h, w, c = x.data.shape
        y = Act(self._new(n, h, w, c))
        one, zero = self._const(c)
        check(L.evb_bn_apply(ptr(x.data), ptr(one), ptr(zero), None, ptr(y.data), c_ll(n * h * w), c_int(c), c_int(1),
                             stream()), 'evb_bn_apply')
        self._tf_fwd(y, name)
        if train:
            def bwd():
                if y.grad is None:
                    return
                self._tf_bwd(y)
                gx, acc = self._grad_into(x)
                assert not acc
                check(L.evb_norm_bwd_apply(ptr(y.grad), ptr(y.data), ptr(y.data), ptr(one), ptr(zero), ptr(zero), ptr(zero),
                                           c_int(1), ptr(gx), None, c_int(0), c_ll(n * h * w), c_int(c), stream()),
                      'evb_norm_bwd_apply')
            self.tape.append(bwd)
        return y

    def _const(self, c):
        if getattr(self, '_consts', None) is None or self._consts[0].numel() < c:
            self._consts = (torch.ones(max(c, 2048), device=self.dev), torch.zeros(max(c, 2048), device=self.dev))
        return self._consts

    def se_block(self, x, se, train=True):
        """SEBlock.forward (se_block.py:19-24): y = x * sigmoid(W2 relu(W1 gap(x) + b1) + b2)"""
        L = self.L
        n, h, w, c = x.data.shape
        l1, l2 = se.seq[0], se.seq[2]
        cr, hid, hw = l1.in_features, l1.out_features, h * w
        f32 = torch.float32
        v = self._new(n, c, dtype=f32)
        one, zero = self._const(c)
        scr3 = self._new(3, c, dtype=f32)
        for s in range(n):   # global average pool = the per-channel mean of the BatchNorm statistics kernel (full-grid column
            ws = self._ws(L.evb_bn_workspace(c_ll(hw), c_int(c)))   # sums; evb_gap_fwd is written for the 16 x 16 scene map)
            check(L.evb_bn_stats(ptr(x.data[s]), c_ll(hw), c_int(c), ptr(one), ptr(zero), None, None, c_float(0.0),
                                 c_float(1e-5), ptr(v[s]), ptr(scr3[0]), ptr(scr3[1]), ptr(scr3[2]), ptr(ws), stream()),
                  'evb_bn_stats')
        vc = self._new(n, cr, dtype=f32)      # the real channels, compact rows for the linears
        check(L.evb_copy2d_f32(ptr(v), c_int(c), ptr(vc), c_int(cr), c_int(n), c_int(cr), c_int(0), stream()), 'evb_copy2d_f32')
        h1, s2 = self._new(n, hid, dtype=f32), self._new(n, cr, dtype=f32)
        check(L.evb_linear_fwd(ptr(vc), ptr(l1.weight), ptr(l1.bias), ptr(h1), c_int(n), c_int(cr), c_int(hid), c_int(1),
                               stream()), 'evb_linear_fwd')
        check(L.evb_linear_fwd(ptr(h1), ptr(l2.weight), ptr(l2.bias), ptr(s2), c_int(n), c_int(hid), c_int(cr), c_int(0),
                               stream()), 'evb_linear_fwd')
        sigc = self._new(n, cr, dtype=f32)
        check(L.evb_sigmoid_fwd(ptr(s2), ptr(sigc), c_int(n * cr), stream()), 'evb_sigmoid_fwd')
        sig = torch.zeros(n, c, dtype=f32, device=self.dev) if c != cr else sigc
        if c != cr:
            check(L.evb_copy2d_f32(ptr(sigc), c_int(cr), ptr(sig), c_int(c), c_int(n), c_int(cr), c_int(0), stream()),
                  'evb_copy2d_f32')
        y = Act(self._new(n, h, w, c))
        check(L.evb_channel_scale(ptr(x.data), ptr(sig), ptr(y.data), c_int(n), c_ll(hw), c_int(c), stream()),
              'evb_channel_scale')
        self._tf_fwd(y, self._modname.get(id(se)))
        if train:
            def bwd():
                if y.grad is None:
                    return
                self._tf_bwd(y)
                gx, acc = self._grad_into(x)
                assert not acc
                check(L.evb_channel_scale(ptr(y.grad), ptr(sig), ptr(gx), c_int(n), c_ll(hw), c_int(c), stream()),
                      'evb_channel_scale')
                one, zero = self._const(c)
                dsig = self._new(n, c, dtype=f32)
                scr = self._new(c, dtype=f32)
                for s in range(n):    # d(sig)[n, c] = sum over pixels of dy * x
                    ws = self._ws(L.evb_bn_workspace(c_ll(hw), c_int(c)))
                    check(L.evb_norm_bwd_reduce(ptr(y.grad[s]), ptr(x.data[s]), None, ptr(zero), ptr(one), ptr(one), ptr(zero),
                                                c_int(0), ptr(dsig[s]), ptr(scr), c_int(0), c_ll(hw), c_int(c), ptr(ws),
                                                stream()), 'evb_norm_bwd_reduce')
                dsc = self._new(n, cr, dtype=f32)
                check(L.evb_copy2d_f32(ptr(dsig), c_int(c), ptr(dsc), c_int(cr), c_int(n), c_int(cr), c_int(0), stream()),
                      'evb_copy2d_f32')
                ds2 = self._new(n, cr, dtype=f32)
                check(L.evb_sigmoid_bwd(ptr(dsc), ptr(sigc), ptr(ds2), c_int(n * cr), stream()), 'evb_sigmoid_bwd')
                a_ = c_int(1 if self.accumulate else 0)
                dh1, dvc = self._new(n, hid, dtype=f32), self._new(n, cr, dtype=f32)
                check(L.evb_linear_bwd(ptr(ds2), ptr(s2), ptr(h1), ptr(l2.weight), ptr(self._g(l2.weight)), ptr(self._g(l2.bias)),
                                       ptr(dh1), c_int(n), c_int(hid), c_int(cr), c_int(0), a_, c_int(0), stream()),
                      'evb_linear_bwd')
                check(L.evb_linear_bwd(ptr(dh1), ptr(h1), ptr(vc), ptr(l1.weight), ptr(self._g(l1.weight)), ptr(self._g(l1.bias)),
                                       ptr(dvc), c_int(n), c_int(cr), c_int(hid), c_int(1), a_, c_int(0), stream()),
                      'evb_linear_bwd')
                dv = torch.zeros(n, c, dtype=f32, device=self.dev) if c != cr else dvc
                if c != cr:
                    check(L.evb_copy2d_f32(ptr(dvc), c_int(cr), ptr(dv), c_int(c), c_int(n), c_int(cr), c_int(0), stream()),
                          'evb_copy2d_f32')
                check(L.evb_gap_bwd(ptr(dv), ptr(gx), c_int(n), c_int(hw), c_int(c), stream()), 'evb_gap_bwd')
            self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ network
    def _features(self, x, train):
        L = self.L
        n, cin, h, w = x.shape
        if h % 8 or w % 8:
            raise ValueError('FreeNetB200 needs H, W divisible by 8 (three stride-2 stages and their nearest-x2 top-down adds)')
        c0 = self.convs[0]
        a = self._new(n, h, w, c0.cip)
        # NCHW fp32 -> NHWC bf16 with the channels zero-padded to the conv tile: an im2col with a 1x1 window
        check(L.evb_im2col_nchw(ptr(x), ptr(a), c_int(n), c_int(cin), c_int(h), c_int(w), c_int(c0.cip), c_int(1), c_int(1),
                                c_int(0), stream()), 'evb_im2col_nchw')
        self._refresh_gn()
        y = Act(a, needs_grad=False)
        feats = []
        for op in self.ops:
            if op[0] == 'feat':
                feats.append(y)
            elif op[0] == 'cgr':
                o = self.conv(y, op[1], bias=True, train=train, dgrad=op[1].need_dgrad)
                y = self.gn_act(o, op[2], name=self._modname.get(id(op[3])), train=train)
            elif op[0] == 'down':
                o = self.conv(y, op[1], bias=True, train=train)
                y = self.relu(o, name=self._modname.get(id(op[2])), train=train)
            else:
                y = self.se_block(y, op[1], train=train)
        out = self.conv(self.conv(feats[3], self.reduce[3], bias=True, train=train), self.fuse[0], bias=True, train=train)
        for i in range(3):
            lvl = 2 - i
            inner = self.conv(feats[lvl], self.reduce[lvl], bias=True, add=out, add_mode=2, train=train,
                              name=self.fuse[i + 1].name + ':in' if self.fuse[i + 1].name else None)
            out = self.conv(inner, self.fuse[i + 1], bias=True, train=train)
        return out

    def _network_losses(self, x, labels):
        y, wmask = labels['cls'], labels['w']
        lab = torch.where((wmask > 0) & (y > 0), y.long() - 1, torch.full_like(y, 255, dtype=torch.long))
        out = self._features(x, True)
        cls, logits = self._classify(out, self.cls, 1, True)
        self._loss_stats(cls, logits, lab, self.K, 1, ('cls_loss', 'dice_loss'))

    def _forward_part2(self):
        out = super()._forward_part2()
        out.pop('dice_loss', None)   # weight 0: FreeNet's loss is the masked cross-entropy alone
        return out

    @torch.no_grad()
    def forward_eval(self, x, return_mask=False):
        L = self.L
        self.tape = []
        x = x.contiguous().float()
        self.pack_weights()
        out = self._features(x, False)
        cls, logits = self._classify(out, self.cls, 1, False)
        n, hh, ww, ld = logits.shape
        prob = self._new(n, self.K, hh, ww, dtype=torch.float32)
        mask = self._new(n, hh, ww, dtype=torch.uint8)
        check(L.evb_softmax_nchw(ptr(logits), ptr(prob), ptr(mask), c_ll(n * hh * ww), c_int(hh * ww), c_int(self.K), c_int(ld),
                                 stream()), 'evb_softmax_nchw')
        self.last_logits = logits
        return (prob, mask) if return_mask else prob


# ------------------------------------------------------------------------------------------------ the plugin model
@MODEL.register('FreeNetB200')
class FreeNetB200(NativeStepMixin, ERModule):
    """forward(x[N, Cin, H, W], y, w): training -> {'cls_loss'}; eval -> softmax probabilities [N, K, H, W].
    y: labels 1..K (0 = unlabelled), w: {0, 1} mask of the training pixels (or y = dict(cls=..., w=...))."""

    def __init__(self, config=None):
        super().__init__(config)
        cfg = self.config
        r = int(16 * float(cfg.reduction_ratio))
        ch = [int(c * float(cfg.reduction_ratio) / r) * r for c in cfg.block_channels]
        nb = tuple(cfg.num_blocks)
        ops = [_conv3x3_gn_relu(int(cfg.in_channels), ch[0], r), _repeat_block(ch[0], r, nb[0]), nn.Identity()]
        for i in range(1, 4):
            ops += [_downsample2x(ch[i - 1], ch[i]), _repeat_block(ch[i], r, nb[i]), nn.Identity()]
        self.feature_ops = nn.ModuleList(ops)
        inner = int(int(cfg.inner_dim) * float(cfg.reduction_ratio))
        self.reduce_1x1convs = nn.ModuleList([nn.Conv2d(c, inner, 1) for c in ch])
        self.fuse_3x3convs = nn.ModuleList([nn.Conv2d(inner, inner, 3, 1, 1) for _ in ch])
        self.cls_pred_conv = nn.Conv2d(inner, int(cfg.num_classes), 1)
        self.engine = None

    def set_default_config(self):
        self.config.update(dict(in_channels=204, num_classes=16, block_channels=(96, 128, 192, 256), num_blocks=(1, 1, 1, 1),
                                inner_dim=128, reduction_ratio=1.0, cuda_graph=False))

    def _engine(self):
        if self.engine is None:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('FreeNetB200 computes only on a CUDA (sm_100a) device; there is no CPU path')
            object.__setattr__(self, 'engine', FreeNetEngine(self))
        return self.engine

    def forward(self, x, y=None, w=None):
        eng = self._engine()
        if self.training:
            if isinstance(y, dict):
                y, w = y['cls'], y['w']
            if y is None or w is None:
                raise ValueError('training forward needs labels y (1..K, 0 = unlabelled) and the training mask w')
            labels = dict(cls=y.contiguous(), w=w.contiguous())
            return self._train_forward(eng, x.contiguous().float(), labels)
        return eng.forward_eval(x)


MODEL.register('FreeNet', FreeNetB200, override=True) if hasattr(MODEL, 'register') else None
