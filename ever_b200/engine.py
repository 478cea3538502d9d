"""FarSeg step engine: schedules the libevb200.so kernels for forward, loss and backward.

Host side of the hot path (the reference's equivalent is autograd over nn.Modules:
ResNetEncoder.forward ever/module/resnet.py:183-211, FarSegHead.forward ever/module/fs_relation.py:174-181,
FPN.forward fpn.py:80-115, FSRelation.forward fs_relation.py:57-73, AssymetricDecoder.forward fpn.py:183-193,
losses loss.py:54-75 + F.cross_entropy).  Here every op is one or two C-ABI calls on raw device pointers;
torch tensors are only the memory they point at.  A small tape of backward closures replaces autograd.

Data layout: activations NHWC bf16; master weights / BN parameters / gradients fp32 in two flat arenas
(parameters are views into them, so one NCCL all-reduce and one fused SGD kernel cover the whole model);
per-step bf16 weight packs [tap][Cout][Cin] (forward) and [tap][Cin][Cout] (dgrad).
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream

c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
BF16 = torch.bfloat16


def _ceil(a, b):
    return (a + b - 1) // b * b


class Act:
    """An activation tensor (NHWC bf16) plus its gradient slot."""
    __slots__ = ('data', 'grad', 'has_grad', 'needs_grad')

    def __init__(self, data, needs_grad=True):
        self.data, self.grad, self.has_grad, self.needs_grad = data, None, False, needs_grad

    @property
    def shape(self):
        return self.data.shape


class ConvP:
    """One convolution's parameters + bf16 packs."""

    def __init__(self, conv, cout_pad=None, cin_pad=None, as_matrix=False):
        self.weight, self.bias = conv.weight, conv.bias
        co, ci, kh, kw = conv.weight.shape
        self.stride = conv.stride[0]
        if as_matrix:  # 7x7 stem lowered to a GEMM over im2col rows: [Co][Ci*49]
            ci, kh, kw = ci * kh * kw, 1, 1
        self.co, self.ci, self.k = co, ci, kh
        self.cop = cout_pad or _ceil(co, 64)   # rows of the forward pack / channels of dy
        self.cip = cin_pad or _ceil(ci, 64)
        self.wf = self.wb = None
        self.need_dgrad = True


class BNP:
    def __init__(self, bn):
        self.bn = bn
        self.c = bn.num_features


class FarSegEngine:
    def __init__(self, module):
        self.m = module
        self.L = lib()
        p0 = next(module.parameters())
        self.dev = p0.device
        self.world = 1
        self.rank = 0
        self._flatten_params()
        self._collect()
        self.tape = []
        self.ws = None
        self.ws_bytes = 0
        self.accumulate = False      # gradient accumulation into existing .grad (forward_times > 1)
        self._saved_for_backward = None
        self.debug = None            # dict -> named activations are recorded (tests / diagnostics)
        cfg = module.config
        self.ignore_index = int(cfg.loss.ignore_index)
        self.ce_w = float(cfg.loss.ce.weight)
        self.dice_w = float(cfg.loss.dice.weight)
        self.smooth = float(cfg.loss.dice.smooth)
        self.sync_dice = bool(cfg.loss.dice.sync_statistics)
        self.K = module.head.fpn_decoder.num_classes
        # K == 1: binary head -> masked BCE-with-logits + sigmoid Dice (ever/module/loss.py:66-68,229-235)

    # ------------------------------------------------------------------ parameters
    def _flatten_params(self):
        params = list(self.m.parameters())
        n = sum(p.numel() for p in params)
        # 16-byte aligned offsets so float4 epilogue loads of biases work
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += _ceil(p.numel(), 4)
        self.flat_w = torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.params = params
        self.grad_views = []
        for p, o in zip(params, offs):
            self.flat_w[o:o + p.numel()].view_as(p).copy_(p.data)
            p.data = self.flat_w[o:o + p.numel()].view_as(p)
            self.grad_views.append(self.flat_g[o:o + p.numel()].view_as(p))
        self.n_params = n

    def set_distributed(self, rank, world):
        """Data parallel over `world` ranks: Dice statistics and the flat gradient arena are all-reduced."""
        self.rank, self.world = int(rank), int(world)

    def attach_grads(self):
        for p, g in zip(self.params, self.grad_views):
            if p.requires_grad:
                p.grad = g

    def _collect(self):
        m = self.m
        r = m.en.resnet
        self.kind = r.kind
        self.convs = []

        def C(conv, **kw):
            cp = ConvP(conv, **kw)
            self.convs.append(cp)
            return cp
        self.stem_kp = _ceil(r.conv1.in_channels * 49, 64)
        self.stem = C(r.conv1, as_matrix=True)
        self.stem.need_dgrad = False
        self.stem_bn = BNP(r.bn1)
        self.stages = []
        for li in range(1, 5):
            blocks = []
            for b in getattr(r, 'layer%d' % li):
                d = dict(stride=b.stride, c1=C(b.conv1), b1=BNP(b.bn1), c2=C(b.conv2), b2=BNP(b.bn2))
                if self.kind == 'bottleneck':
                    d.update(c3=C(b.conv3), b3=BNP(b.bn3))
                if b.downsample is not None:
                    d.update(cd=C(b.downsample[0]), bd=BNP(b.downsample[1]))
                blocks.append(d)
            self.stages.append(blocks)
        h = m.head
        self.fpn_inner = [C(getattr(h.fpn, 'fpn_inner%d' % i)[0]) for i in range(1, 5)]
        self.fpn_layer = [C(getattr(h.fpn, 'fpn_layer%d' % i)[0]) for i in range(1, 5)]
        fs = h.fs_relation
        self.scene = [(s[0], s[2]) for s in fs.scene_encoder]  # (conv 2048->256, conv 256->256) as linears
        self.content = [(C(s[0]), BNP(s[1])) for s in fs.content_encoders]
        self.reenc = [(C(s[0]), BNP(s[1])) for s in fs.feature_reencoders]
        dec = h.fpn_decoder
        self.dec_blocks = [[(C(layer[0]), BNP(layer[1])) for layer in blk] for blk in dec.blocks]
        self.dec_nup = list(dec.num_upsample)
        self.cls = C(dec.classifier[0], cout_pad=64)
        self.cls_scale = dec.scale_factor
        # bf16 pack arena
        tot = 0
        for cp in self.convs:
            kk = cp.k * cp.k
            cp._off_f = tot
            tot += kk * cp.cop * cp.cip
            cp._off_b = tot
            if cp.need_dgrad:
                tot += kk * cp.cip * cp.cop
            tot = _ceil(tot, 64)
        self.pack = torch.zeros(tot, dtype=BF16, device=self.dev)
        for cp in self.convs:
            kk = cp.k * cp.k
            cp.wf = self.pack[cp._off_f:cp._off_f + kk * cp.cop * cp.cip].view(kk, cp.cop, cp.cip)
            if cp.need_dgrad:
                cp.wb = self.pack[cp._off_b:cp._off_b + kk * cp.cip * cp.cop].view(kk, cp.cip, cp.cop)
        # padded classifier bias (64 floats) and scratch for padded wgrad outputs
        self.cls_bias_pad = torch.zeros(64, dtype=torch.float32, device=self.dev)
        self.scr_cls_dw = torch.zeros(64 * self.cls.ci, dtype=torch.float32, device=self.dev)
        self.scr_cls_db = torch.zeros(64, dtype=torch.float32, device=self.dev)
        self.scr_stem_dw = torch.zeros(64 * self.stem_kp, dtype=torch.float32, device=self.dev)

    def _build_pack_table(self):
        rows, bmap, nblk = [], [], 0
        for i, cp in enumerate(self.convs):
            kk = cp.k * cp.k
            nb = ((cp.co + 31) // 32) * ((cp.ci + 31) // 32)   # one block per 32x32 (co, ci) tile, all taps
            rows.append([cp.weight.data_ptr(), cp.wf.data_ptr(), cp.wb.data_ptr() if cp.need_dgrad else 0, cp.co, cp.ci, kk,
                         cp.cop, cp.cip, cp.cip, cp.cop, nblk, 0])
            bmap += [i] * nb
            nblk += nb
        self._pack_desc = torch.tensor(rows, dtype=torch.int64, device=self.dev)
        self._pack_map = torch.tensor(bmap, dtype=torch.int32, device=self.dev)
        self._pack_nblk = nblk
        self._pack_ptrs = [cp.weight.data_ptr() for cp in self.convs]

    def pack_weights(self):
        """fp32 master -> bf16 packs for every convolution, one launch (once per step, after the optimizer)."""
        st = stream()
        if getattr(self, '_pack_desc', None) is None or self._pack_ptrs != [cp.weight.data_ptr() for cp in self.convs]:
            self._build_pack_table()
        check(self.L.evb_pack_weights_batched(ptr(self._pack_desc), ptr(self._pack_map), c_int(self._pack_nblk), st),
              'evb_pack_weights_batched')
        check(self.L.evb_copy2d_f32(ptr(self.cls.bias), c_int(self.K), ptr(self.cls_bias_pad), c_int(64), c_int(1),
                                    c_int(self.K), c_int(0), st), 'evb_copy2d_f32')

    # ------------------------------------------------------------------ workspace
    def _ws(self, nbytes):
        if nbytes > self.ws_bytes:
            self.ws_bytes = max(nbytes, 64 << 20)
            self.ws = torch.empty(self.ws_bytes // 4, dtype=torch.float32, device=self.dev)
        return self.ws

    def _dbg(self, name, t):
        if self.debug is not None:
            self.debug[name] = t.data if isinstance(t, Act) else t

    def _new(self, *shape, dtype=BF16):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    # ------------------------------------------------------------------ ops (forward + tape)
    def _grad_into(self, act, shape=None):
        """Return (buffer, accumulate_flag) for writing a gradient contribution of `act`."""
        if act.grad is None:
            act.grad = self._new(*act.data.shape)
        acc = act.has_grad
        act.has_grad = True
        return act.grad, acc

    def conv(self, x, cp, stride=None, bias=None, add=None, add_mode=0, out_channels=None, train=True):
        L = self.L
        stride = cp.stride if stride is None else stride
        n, h, w, cin = x.data.shape
        ho, wo = h // stride, w // stride
        cout = out_channels or cp.co
        y = Act(self._new(n, ho, wo, cout))
        check(L.evb_conv2d_fwd(ptr(x.data), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(cp.wf), c_int(cp.cop),
                               c_int(cp.k), c_int(stride), ptr(y.data), c_int(cout), ptr(bias),
                               ptr(add.data if add is not None else None), c_int(add_mode), c_int(0), stream()),
              'evb_conv2d_fwd')
        if train:
            def bwd():
                if y.grad is None:
                    return
                dy = y.grad
                st = stream()
                wgrad_target, copy_rows = cp.weight.grad, None
                if cp is self.cls:
                    wgrad_target, copy_rows = self.scr_cls_dw, self.K
                nbytes = L.evb_conv2d_wgrad_workspace(c_int(n), c_int(ho), c_int(wo), c_int(cin), c_int(cout),
                                                      c_int(cp.k), c_int(0), c_int(0))
                ws = self._ws(nbytes)
                acc = self.accumulate and copy_rows is None
                if cp.weight.grad is not None:
                  check(L.evb_conv2d_wgrad(ptr(x.data), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(dy), c_int(cout),
                                         c_int(cp.k), c_int(stride), ptr(wgrad_target), c_int(1 if acc else 0), ptr(ws),
                                         c_ll(self.ws_bytes), c_int(0), c_int(0), st), 'evb_conv2d_wgrad')
                if copy_rows is not None and cp.weight.grad is not None:
                    check(L.evb_copy2d_f32(ptr(wgrad_target), c_int(cp.ci), ptr(cp.weight.grad), c_int(cp.ci),
                                           c_int(copy_rows), c_int(cp.ci), c_int(1 if self.accumulate else 0), st),
                          'evb_copy2d_f32')
                if cp.bias is not None and cp.bias.grad is not None:
                    m_rows = n * ho * wo
                    ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(cout)))
                    if cp is self.cls:
                        check(L.evb_bias_grad(ptr(dy), c_ll(m_rows), c_int(cout), ptr(self.scr_cls_db), None, c_int(0),
                                              ptr(ws), st), 'evb_bias_grad')
                        check(L.evb_copy2d_f32(ptr(self.scr_cls_db), c_int(64), ptr(cp.bias.grad), c_int(self.K),
                                               c_int(1), c_int(self.K), c_int(1 if self.accumulate else 0), st),
                              'evb_copy2d_f32')
                    else:
                        check(L.evb_bias_grad(ptr(dy), c_ll(m_rows), c_int(cout), ptr(cp.bias.grad), None,
                                              c_int(1 if self.accumulate else 0), ptr(ws), st), 'evb_bias_grad')
                if add is not None and add.needs_grad:
                    g, acc2 = self._grad_into(add)
                    if add_mode == 2:
                        check(L.evb_sumpool2(ptr(dy), ptr(g), c_int(n), c_int(ho // 2), c_int(wo // 2), c_int(cout),
                                             c_int(1 if acc2 else 0), st), 'evb_sumpool2')
                    else:
                        check(L.evb_scale_add(ptr(dy), c_float(1.0), ptr(g if acc2 else None), ptr(g),
                                              c_ll(dy.numel()), st), 'evb_scale_add')
                if x.needs_grad and cp.need_dgrad:
                    g, acc2 = self._grad_into(x)
                    check(L.evb_conv2d_dgrad(ptr(dy), c_int(n), c_int(ho), c_int(wo), c_int(cout), ptr(cp.wb),
                                             c_int(cp.cip), c_int(cp.k), c_int(stride), ptr(g), c_int(h), c_int(w),
                                             c_int(cin), c_int(1 if acc2 else 0), c_int(0), st), 'evb_conv2d_dgrad')
            self.tape.append(bwd)
        return y

    def _bn_fold(self, x, bp, train):
        """Batch statistics (training) or running statistics (eval) -> mean, rstd, scale, shift."""
        L = self.L
        bn = bp.bn
        c = bp.c
        stats = self._new(4, c, dtype=torch.float32)
        mean, rstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        if train and bn.training:
            m_rows = x.data.numel() // c
            ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(c)))
            mom = 0.1 if bn.momentum is None else bn.momentum
            check(L.evb_bn_stats(ptr(x.data), c_ll(m_rows), c_int(c), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
                                 ptr(bn.running_var), c_float(mom), c_float(bn.eps), ptr(mean), ptr(rstd), ptr(scale),
                                 ptr(shift), ptr(ws), stream()), 'evb_bn_stats')
            self._bn_tracked.append(bn)
        else:
            check(L.evb_bn_fold(ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), c_float(bn.eps),
                                c_int(c), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), stream()), 'evb_bn_fold')
        return mean, rstd, scale, shift

    def _bn_backward(self, dy, x, bp, fold, mask_mode, ymask, dres_act):
        L = self.L
        mean, rstd, scale, shift = fold
        c = bp.c
        m_rows = x.data.numel() // c
        ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(c)))
        gx, _ = self._grad_into(x)
        dres, dres_acc = (None, False)
        if dres_act is not None and dres_act.needs_grad:
            dres, dres_acc = self._grad_into(dres_act)
        check(L.evb_bn_bwd(ptr(dy), ptr(x.data), ptr(ymask), ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
                           c_int(mask_mode), c_int(0 if bp.bn.training else 1), ptr(gx), ptr(dres), c_int(1 if dres_acc else 0),
                           ptr(bp.bn.weight.grad), ptr(bp.bn.bias.grad), c_int(1 if self.accumulate else 0),
                           c_ll(m_rows), c_int(c), ptr(ws), stream()), 'evb_bn_bwd')

    def bn_act(self, x, bp, relu=True, res=None, train=True):
        L = self.L
        fold = self._bn_fold(x, bp, train)
        y = Act(self._new(*x.data.shape))
        c = bp.c
        check(L.evb_bn_apply(ptr(x.data), ptr(fold[2]), ptr(fold[3]), ptr(res.data if res is not None else None),
                             ptr(y.data), c_ll(x.data.numel() // c), c_int(c), c_int(1 if relu else 0), stream()),
              'evb_bn_apply')
        if train:
            def bwd():
                if y.grad is None:
                    return
                if relu and res is None:   # mask recomputed from x: one tensor read less per pass
                    self._bn_backward(y.grad, x, bp, fold, 2, None, None)
                else:
                    self._bn_backward(y.grad, x, bp, fold, 1 if relu else 0, y.data if relu else None, res)
            self.tape.append(bwd)
        return y

    def bn_relu_up(self, x, bp, f=2, train=True):
        """decoder stage: bilinear x f of relu(bn(x)) (fpn.py:163-169).  BN+ReLU is applied once at the low resolution
        (the tensor is f*f times smaller than the output), then a pure bilinear kernel writes the up-sampled map."""
        L = self.L
        fold = self._bn_fold(x, bp, train)
        n, h, w, c = x.data.shape
        low_y = self._new(n, h, w, c)
        check(L.evb_bn_apply(ptr(x.data), ptr(fold[2]), ptr(fold[3]), None, ptr(low_y), c_ll(n * h * w), c_int(c), c_int(1),
                             stream()), 'evb_bn_apply')
        y = Act(self._new(n, h * f, w * f, c))
        check(L.evb_bilinear_up(ptr(low_y), None, None, ptr(y.data), c_int(n), c_int(h), c_int(w), c_int(c), c_int(c),
                                c_int(c), c_int(f), stream()), 'evb_bilinear_up')
        if train:
            def bwd():
                if y.grad is None:
                    return
                low = self._new(n, h, w, c)
                check(L.evb_bilinear_up_bwd(ptr(y.grad), ptr(low), c_int(n), c_int(h), c_int(w), c_int(c), c_int(c),
                                            c_int(c), c_int(f), stream()), 'evb_bilinear_up_bwd')
                self._bn_backward(low, x, bp, fold, 2, None, None)
            self.tape.append(bwd)
        return y

    def maxpool(self, x, train=True):
        L = self.L
        n, h, w, c = x.data.shape
        y = Act(self._new(n, h // 2, w // 2, c))
        idx = self._new(n, h // 2, w // 2, c, dtype=torch.uint8)
        check(L.evb_maxpool3x3s2_fwd(ptr(x.data), ptr(y.data), ptr(idx), c_int(n), c_int(h), c_int(w), c_int(c), stream()),
              'evb_maxpool3x3s2_fwd')
        if train:
            def bwd():
                if y.grad is None or not x.needs_grad:
                    return
                g, acc = self._grad_into(x)
                assert not acc
                check(L.evb_maxpool3x3s2_bwd(ptr(y.grad), ptr(idx), ptr(g), c_int(n), c_int(h), c_int(w), c_int(c),
                                             stream()), 'evb_maxpool3x3s2_bwd')
            self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ network
    def _block(self, x, d, train):
        if self.kind == 'bottleneck':
            a1 = self.bn_act(self.conv(x, d['c1'], train=train), d['b1'], True, train=train)
            a2 = self.bn_act(self.conv(a1, d['c2'], train=train), d['b2'], True, train=train)
            o3 = self.conv(a2, d['c3'], train=train)
            last_bn = d['b3']
        else:
            a1 = self.bn_act(self.conv(x, d['c1'], train=train), d['b1'], True, train=train)
            o3 = self.conv(a1, d['c2'], train=train)
            last_bn = d['b2']
        idt = x
        if 'cd' in d:
            idt = self.bn_act(self.conv(x, d['cd'], train=train), d['bd'], False, train=train)
        return self.bn_act(o3, last_bn, True, res=idt, train=train)

    def _encoder(self, x_nchw, train):
        L = self.L
        n, cin, h, w = x_nchw.shape
        if h % 32 or w % 32:
            raise ValueError('FarSegB200 needs H, W divisible by 32 (FPN nearest-x2 adds, SURVEY.md section 5)')
        a = self._new(n, h // 2, w // 2, self.stem_kp)
        check(L.evb_stem_im2col(ptr(x_nchw), ptr(a), c_int(n), c_int(cin), c_int(h), c_int(w), c_int(self.stem_kp),
                                stream()), 'evb_stem_im2col')
        xa = Act(a, needs_grad=False)
        y0 = self.conv(xa, self.stem, stride=1, train=False)   # distinct name: the closure below must keep THIS Act
        if train:
            stem, ho, wo = self.stem, h // 2, w // 2

            def stem_bwd():
                if y0.grad is None:
                    return
                nbytes = L.evb_conv2d_wgrad_workspace(c_int(n), c_int(ho), c_int(wo), c_int(self.stem_kp), c_int(64),
                                                      c_int(1), c_int(0), c_int(0))
                ws = self._ws(nbytes)
                check(L.evb_conv2d_wgrad(ptr(a), c_int(n), c_int(ho), c_int(wo), c_int(self.stem_kp), ptr(y0.grad),
                                         c_int(64), c_int(1), c_int(1), ptr(self.scr_stem_dw), c_int(0), ptr(ws),
                                         c_ll(self.ws_bytes), c_int(0), c_int(0), stream()), 'evb_conv2d_wgrad(stem)')
                check(L.evb_copy2d_f32(ptr(self.scr_stem_dw), c_int(self.stem_kp), ptr(stem.weight.grad), c_int(stem.ci),
                                       c_int(64), c_int(stem.ci), c_int(1 if self.accumulate else 0), stream()),
                      'evb_copy2d_f32')
            self.tape.append(stem_bwd)
        y = y0
        self._dbg('stem_conv', y)
        y = self.bn_act(y, self.stem_bn, True, train=train)
        self._dbg('stem_act', y)
        y = self.maxpool(y, train=train)
        self._dbg('pool', y)
        feats = []
        freeze_at = int(self.m.config.encoder.freeze_at)
        for si, blocks in enumerate(self.stages):
            if train and freeze_at >= si + 1:
                y.needs_grad = False   # everything that produced y is frozen: no gradient flows further down
            for d in blocks:
                y = self._block(y, d, train)
            if train and freeze_at >= si + 2:
                y.needs_grad = False   # this stage and everything below it are frozen
            feats.append(y)
            self._dbg('c%d' % (si + 2), y)
        return feats

    def _scene_mlp(self, scene, n, train):
        """4 x (1x1 conv -> ReLU -> 1x1 conv) on the N x C5 x 1 x 1 scene embedding (fs_relation.py:22-28)."""
        L = self.L
        outs = []
        c5 = scene.shape[1]
        dscene = self._new(n, c5, dtype=torch.float32) if train else None
        self._dscene_used = False
        for (l1, l2) in self.scene:
            co = l1.out_channels
            hid = self._new(n, co, dtype=torch.float32)
            sf = self._new(n, co, dtype=torch.float32)
            check(L.evb_linear_fwd(ptr(scene), ptr(l1.weight), ptr(l1.bias), ptr(hid), c_int(n), c_int(c5), c_int(co),
                                   c_int(1), stream()), 'evb_linear_fwd')
            check(L.evb_linear_fwd(ptr(hid), ptr(l2.weight), ptr(l2.bias), ptr(sf), c_int(n), c_int(co), c_int(co),
                                   c_int(0), stream()), 'evb_linear_fwd')
            dsf = self._new(n, co, dtype=torch.float32) if train else None
            outs.append((sf, dsf))
            if train:
                def bwd(l1=l1, l2=l2, hid=hid, sf=sf, dsf=dsf, co=co):
                    acc = 1 if self.accumulate else 0
                    dhid = self._new(n, co, dtype=torch.float32)
                    check(L.evb_linear_bwd(ptr(dsf), ptr(sf), ptr(hid), ptr(l2.weight), ptr(l2.weight.grad),
                                           ptr(l2.bias.grad), ptr(dhid), c_int(n), c_int(co), c_int(co), c_int(0),
                                           c_int(acc), c_int(0), stream()), 'evb_linear_bwd')
                    check(L.evb_linear_bwd(ptr(dhid), ptr(hid), ptr(scene), ptr(l1.weight), ptr(l1.weight.grad),
                                           ptr(l1.bias.grad), ptr(dscene), c_int(n), c_int(c5), c_int(co), c_int(1),
                                           c_int(acc), c_int(1 if self._dscene_used else 0), stream()), 'evb_linear_bwd')
                    self._dscene_used = True
                self.tape.append(bwd)
        return outs, dscene

    def _head(self, feats, train):
        L = self.L
        # ---- FPN (top-down, nearest x2 fused into the lateral 1x1 epilogue)
        inner = [None] * 4
        inner[3] = self.conv(feats[3], self.fpn_inner[3], train=train)
        ps = [None] * 4
        ps[3] = self.conv(inner[3], self.fpn_layer[3], train=train)
        for i in (2, 1, 0):
            inner[i] = self.conv(feats[i], self.fpn_inner[i], add=inner[i + 1], add_mode=2, train=train)
            ps[i] = self.conv(inner[i], self.fpn_layer[i], train=train)
        for i in range(4):
            self._dbg('p%d' % (i + 2), ps[i])
        # ---- scene embedding
        c5 = feats[3]
        n, h5, w5, cc5 = c5.data.shape
        scene = self._new(n, cc5, dtype=torch.float32)
        check(L.evb_gap_fwd(ptr(c5.data), ptr(scene), c_int(n), c_int(h5 * w5), c_int(cc5), stream()), 'evb_gap_fwd')
        if train:
            # runs last among head closures (registered first): needs dscene complete
            holder = {}

            def gap_bwd():
                g, acc = self._grad_into(c5)
                if not acc:
                    g.zero_()
                check(L.evb_gap_bwd(ptr(holder['dscene']), ptr(g), c_int(n), c_int(h5 * w5), c_int(cc5), stream()),
                      'evb_gap_bwd')
            self.tape.append(gap_bwd)
        sfs, dscene = self._scene_mlp(scene, n, train)
        if train:
            holder['dscene'] = dscene
        # ---- FS-Relation per level
        zs = []
        for i in range(4):
            p = ps[i]
            (cc, cb), (rc, rb) = self.content[i], self.reenc[i]
            u1 = self.conv(p, cc, bias=cc.bias, train=train)
            u2 = self.conv(p, rc, bias=rc.bias, train=train)
            f1 = self._bn_fold(u1, cb, train)
            f2 = self._bn_fold(u2, rb, train)
            nn_, hh, ww, c = u1.data.shape
            m_rows = nn_ * hh * ww
            z = Act(self._new(nn_, hh, ww, c))
            rel = self._new(m_rows, dtype=torch.float32)
            sf, dsf = sfs[i]
            check(L.evb_relation_fwd(ptr(u1.data), ptr(u2.data), ptr(f1[2]), ptr(f1[3]), ptr(f2[2]), ptr(f2[3]), ptr(sf),
                                     ptr(z.data), ptr(rel), c_ll(m_rows), c_int(hh * ww), c_int(c), stream()),
                  'evb_relation_fwd')
            if train:
                def bwd(z=z, u1=u1, u2=u2, f1=f1, f2=f2, sf=sf, dsf=dsf, rel=rel, cb=cb, rb=rb, m_rows=m_rows, hh=hh,
                        ww=ww, c=c):
                    if z.grad is None:
                        return
                    g1 = self._new(*u1.data.shape)
                    g2 = self._new(*u2.data.shape)
                    ws = self._ws(L.evb_relation_bwd_workspace(c_ll(m_rows), c_int(hh * ww), c_int(c)))
                    check(L.evb_relation_bwd(ptr(z.grad), ptr(u1.data), ptr(u2.data), ptr(f1[2]), ptr(f1[3]), ptr(f2[2]),
                                             ptr(f2[3]), ptr(sf), ptr(rel), ptr(g1), ptr(g2), ptr(dsf), c_ll(m_rows),
                                             c_int(hh * ww), c_int(c), ptr(ws), stream()), 'evb_relation_bwd')
                    self._bn_backward(g1, u1, cb, f1, 0, None, None)
                    self._bn_backward(g2, u2, rb, f2, 0, None, None)
                # must run BEFORE the conv backward closures of u1/u2 (registered earlier => run later): ok
                self.tape.append(bwd)
            zs.append(z)
            self._dbg('z%d' % i, z)
            self._dbg('rel%d' % i, rel)
            self._dbg('sf%d' % i, sf)
        # ---- asymmetric decoder
        outs = []
        for i in range(4):
            y = zs[i]
            for (cp, bp) in self.dec_blocks[i]:
                o = self.conv(y, cp, train=train)
                y = self.bn_relu_up(o, bp, 2, train=train) if self.dec_nup[i] else self.bn_act(o, bp, True, train=train)
            outs.append(y)
            self._dbg('dec%d' % i, y)
        merged = Act(self._new(*outs[0].data.shape))
        check(L.evb_merge4(ptr(outs[0].data), ptr(outs[1].data), ptr(outs[2].data), ptr(outs[3].data), ptr(merged.data),
                           c_ll(merged.data.numel()), stream()), 'evb_merge4')
        if train:
            def bwd():
                if merged.grad is None:
                    return
                dq = self._new(*merged.data.shape)
                check(L.evb_scale_add(ptr(merged.grad), c_float(0.25), None, ptr(dq), c_ll(dq.numel()), stream()),
                      'evb_scale_add')
                for o in outs:
                    o.grad, o.has_grad = dq, True
            self.tape.append(bwd)
        # ---- classifier (Cout padded to 64) + bilinear x4 on the K class channels (channel stride 16)
        cls = self.conv(merged, self.cls, bias=self.cls_bias_pad, out_channels=64, train=train)
        n, h4, w4, _ = cls.data.shape
        f = self.cls_scale
        logits = self._new(n, h4 * f, w4 * f, 16)
        check(L.evb_bilinear_up(ptr(cls.data), None, None, ptr(logits), c_int(n), c_int(h4), c_int(w4), c_int(16),
                                c_int(64), c_int(16), c_int(f), stream()), 'evb_bilinear_up(logits)')
        self._dbg('merged', merged)
        self._dbg('cls', cls)
        self._dbg('logits', logits)
        return cls, logits

    # ------------------------------------------------------------------ public steps
    def _forward_part1(self, x, labels):
        """pack weights, encoder, head, loss statistics (everything before the Dice all-reduce)."""
        L = self.L
        self.tape = []
        self._bn_tracked = []
        x = x.contiguous().float()
        labels = labels.contiguous()
        if labels.dtype != torch.int64:
            labels = labels.long()
        self.pack_weights()
        self.attach_grads()
        feats = self._encoder(x, True)
        cls, logits = self._head(feats, True)
        n, hh, ww, _ = logits.shape
        npx = n * hh * ww
        k = self.K
        stats = self._new(2 + 3 * k, dtype=torch.float32)
        ws = self._ws(L.evb_loss_workspace(c_ll(npx), c_int(k)))
        check(L.evb_loss_stats(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(16), c_int(self.ignore_index),
                               ptr(stats), ptr(ws), stream()), 'evb_loss_stats')
        if self._bn_tracked:
            torch._foreach_add_([bn.num_batches_tracked for bn in self._bn_tracked], 1)
        self._part1 = (cls, logits, labels, stats, npx)
        if self.world > 1 and self.sync_dice:
            if getattr(self, '_dice_global', None) is None:
                self._dice_global = torch.zeros(3 * k, dtype=torch.float32, device=self.dev)

    def _dice_allreduce(self):
        """all_reduce_sum of the Dice statistics (ever/module/loss.py:20-23,46-48); eager NCCL, never graph-captured."""
        if self.world > 1 and self.sync_dice:
            import torch.distributed as dist
            self._dice_global.copy_(self._part1[3][2:])
            dist.all_reduce(self._dice_global)

    def _forward_part2(self):
        L = self.L
        cls, logits, labels, stats, npx = self._part1
        k = self.K
        glob = self.world > 1 and self.sync_dice
        dice_stats = self._dice_global if glob else stats[2:]
        scale = float(self.world) if glob else 1.0
        losses = self._new(2, dtype=torch.float32)
        coef = self._new(1 + 2 * k, dtype=torch.float32)
        check(L.evb_loss_finalize(ptr(stats), ptr(dice_stats), c_int(k), c_float(self.smooth), c_float(self.ce_w),
                                  c_float(self.dice_w), c_float(scale), ptr(losses), ptr(coef), stream()),
              'evb_loss_finalize')
        self._saved_for_backward = (cls, logits, labels, coef, npx)
        first = 'bce_loss' if self.K == 1 else 'ce_loss'
        return {first: losses[0] * self.ce_w if self.ce_w != 1.0 else losses[0],
                'dice_loss': losses[1] * self.dice_w if self.dice_w != 1.0 else losses[1]}

    def forward_train(self, x, labels):
        self._forward_part1(x, labels)
        self._dice_allreduce()
        return self._forward_part2()

    def backward(self, allreduce=True):
        L = self.L
        if self._saved_for_backward is None:
            raise RuntimeError('backward() without a preceding training forward')
        cls, logits, labels, coef, npx = self._saved_for_backward
        self._saved_for_backward = None
        k = self.K
        n, hh, ww, _ = logits.shape
        dlogits = self._new(n, hh, ww, 16)
        check(L.evb_loss_grad(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(16), c_int(self.ignore_index),
                              ptr(coef), ptr(dlogits), stream()), 'evb_loss_grad')
        f = self.cls_scale
        cls.grad = torch.zeros_like(cls.data)   # padding channels 16..63 stay zero
        cls.has_grad = True
        check(L.evb_bilinear_up_bwd(ptr(dlogits), ptr(cls.grad), c_int(n), c_int(hh // f), c_int(ww // f), c_int(16),
                                    c_int(16), c_int(64), c_int(f), stream()), 'evb_bilinear_up_bwd(logits)')
        for fn in reversed(self.tape):
            fn()
        self.tape = []
        if allreduce:
            self.allreduce_grads()

    def allreduce_grads(self):
        """The one gradient exchange of the step: NCCL all-reduce (mean) of the flat fp32 gradient arena
        (reference: DDP bucketed all-reduce, ever/trainer/th_ddp_trainer.py:25-30)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.AVG)

    # ------------------------------------------------------------------ fused optimizer
    def sgd_step(self, lr, momentum=0.9, weight_decay=1e-4, max_norm=35.0):
        """clip_grad_norm_(max_norm, 2) + torch.optim.SGD step + zero_grad over the flat arenas
        (ever/interface/module.py:83-108, ever/opt/optimizer.py:7-9) as two kernels; lr lives on the device."""
        L = self.L
        n = self.flat_w.numel()
        if not hasattr(self, '_mom'):
            self._mom = torch.zeros_like(self.flat_w)
            self._lr = torch.zeros(1, dtype=torch.float32, device=self.dev)
            self._norm = torch.zeros(2, dtype=torch.float32, device=self.dev)
            self._sgd_ws = torch.empty(L.evb_sgd_workspace(c_ll(n)) // 4 + 4, dtype=torch.float32, device=self.dev)
            self._sgd_first = True
        self._lr.fill_(float(lr))
        check(L.evb_grad_norm(ptr(self.flat_g), c_ll(n), c_float(max_norm if max_norm else 0.0), ptr(self._norm),
                              ptr(self._sgd_ws), stream()), 'evb_grad_norm')
        check(L.evb_sgd_step(ptr(self.flat_w), ptr(self.flat_g), ptr(self._mom), c_ll(n), ptr(self._lr),
                             c_float(momentum), c_float(weight_decay), ptr(self._norm), c_int(1 if self._sgd_first else 0),
                             c_int(1), stream()), 'evb_sgd_step')
        self._sgd_first = False
        return self._norm

    # ------------------------------------------------------------------ CUDA-graph step
    def capture_step(self, x, labels):
        """Capture forward + loss + backward for fixed-shape device inputs into CUDA graph(s).
        Returns (replay_fn, losses dict).  NCCL is never captured: with world > 1 the step is two graphs with the
        (eager) Dice-statistics all-reduce between them; the gradient all-reduce and the optimizer run after."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: sets kernel attributes, sizes the workspace
                self.forward_train(x, labels)
                self.backward(allreduce=False)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        split = self.world > 1 and self.sync_dice
        pool = torch.cuda.graph_pool_handle()
        g1 = torch.cuda.CUDAGraph()
        if not split:
            with torch.cuda.graph(g1, pool=pool):
                out = self.forward_train(x, labels)
                self.backward(allreduce=False)
            return g1.replay, out
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, pool=pool):
            self._forward_part1(x, labels)
        self._dice_allreduce()
        torch.cuda.synchronize()
        with torch.cuda.graph(g2, pool=pool):
            out = self._forward_part2()
            self.backward(allreduce=False)

        def replay():
            g1.replay()
            self._dice_allreduce()
            g2.replay()
        return replay, out

    @torch.no_grad()
    def forward_eval(self, x, return_mask=False):
        L = self.L
        self.tape = []
        x = x.contiguous().float()
        self.pack_weights()
        feats = self._encoder(x, False)
        cls, logits = self._head(feats, False)
        n, hh, ww, _ = logits.shape
        prob = self._new(n, self.K, hh, ww, dtype=torch.float32)
        mask = self._new(n, hh, ww, dtype=torch.uint8)
        check(L.evb_softmax_nchw(ptr(logits), ptr(prob), ptr(mask), c_ll(n * hh * ww), c_int(hh * ww), c_int(self.K),
                                 c_int(16), stream()), 'evb_softmax_nchw')
        self.last_logits = logits
        if return_mask:
            return prob, mask
        return prob
