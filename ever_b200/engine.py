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
import os

import torch

from ._lib import check, lib, ptr, stream

c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
BF16 = torch.bfloat16


def _ceil(a, b):
    return (a + b - 1) // b * b


class Act:
    """An activation tensor (NHWC bf16) plus its gradient slot."""
    __slots__ = ('data', 'grad', 'has_grad', 'needs_grad', 'stats', 'name')

    def __init__(self, data, needs_grad=True, name=None):
        self.data, self.grad, self.has_grad, self.needs_grad = data, None, False, needs_grad
        self.stats = None   # (partial sums, nblk) of BN batch statistics emitted by the producing conv's epilogue
        self.name = name    # reference module path whose output this tensor is (teacher-forced parity tests)

    @property
    def shape(self):
        return self.data.shape


class ConvP:
    """One convolution's parameters + bf16 packs.

    The conv may use a Cin sub-range of its weight tensor (w_off / w_ld, e.g. the two halves of ChangeMixin's first
    conv) and may be zero-padded to the 64-channel granularity of the tensor-core kernel (cop / cip > co / ci): then
    the weight gradient goes through a dense scratch [cop][cip*k*k] and only the valid block is copied out, and the
    bias lives in a padded fp32 buffer."""

    def __init__(self, weight, bias=None, stride=1, co=None, ci=None, k=None, cout_pad=None, cin_pad=None, w_off=0, w_ld=0,
                 groups=1):
        self.weight, self.bias = weight, bias
        wco, wci, kh, kw = weight.shape
        self.stride = stride
        # grouped conv (ResNeXt conv2, _resnets.py:84): runs as a dense conv over a block-diagonal fp32 image of the weight
        # (dense_w, refreshed before every pack); the dense weight gradient lands in gscr and its diagonal blocks are read
        # back (csrc/grouped.cu)
        self.groups = groups
        self.dense_w = None
        self.co = wco if co is None else co
        self.ci = wci * groups if ci is None else ci
        self.k = kh if k is None else k
        self.kk = self.k * self.k
        self.cop = cout_pad or _ceil(self.co, 64)   # rows of the forward pack / channels of dy
        self.cip = cin_pad or _ceil(self.ci, 64)
        self.w_off, self.w_ld = w_off, w_ld
        self.direct_grad = (self.cop == self.co and self.cip == self.ci and w_off == 0 and w_ld == 0 and groups == 1)
        self.wf = self.wb = None
        self.need_dgrad = True
        self.gscr = self.bias_pad = self.dbias_scr = None
        self._gw = False   # weight/bias gradient already written in this step (shared convs accumulate)
        self.name = None   # module path of the nn.Conv2d in the plugin model == the reference's (set by _collect)

    @staticmethod
    def from_conv(conv, cout_pad=None, as_matrix=False):
        if as_matrix:  # 7x7 stem lowered to a GEMM over im2col rows: [Co][Ci*49]
            co, ci, kh, kw = conv.weight.shape
            return ConvP(conv.weight, conv.bias, 1, co=co, ci=ci * kh * kw, k=1, cout_pad=cout_pad)
        return ConvP(conv.weight, conv.bias, conv.stride[0], cout_pad=cout_pad, groups=conv.groups)


class BNP:
    """BatchNorm parameters; cpad > C gives zero-padded shadow buffers (padded channels: gamma 1, beta 0 -> stay 0)."""

    def __init__(self, bn, cpad=None):
        # train.sync_bn (ever/trainer/th_ddp_trainer.py:21-22) converts every BatchNorm2d into nn.SyncBatchNorm: batch
        # statistics over the batch of ALL ranks -- one small all-reduce per layer forward and one backward (engine._bn_fold /
        # _bn_backward), captured as NCCL nodes in the step graph
        self.sync = isinstance(bn, torch.nn.SyncBatchNorm)
        self.bn = bn
        self.c_real = bn.num_features
        self.c = cpad or bn.num_features
        self.padded = self.c != self.c_real
        self._gw = False
        self.name = None
        if self.padded:
            dev = bn.weight.device
            self.gamma_p = torch.ones(self.c, device=dev)
            self.beta_p = torch.zeros(self.c, device=dev)
            self.rm_p = torch.zeros(self.c, device=dev)
            self.rv_p = torch.ones(self.c, device=dev)
            self.dgamma_p = torch.zeros(self.c, device=dev)
            self.dbeta_p = torch.zeros(self.c, device=dev)

    gamma = property(lambda self: self.gamma_p if self.padded else self.bn.weight)
    beta = property(lambda self: self.beta_p if self.padded else self.bn.bias)
    rm = property(lambda self: self.rm_p if self.padded else self.bn.running_mean)
    rv = property(lambda self: self.rv_p if self.padded else self.bn.running_var)


class FarSegEngine:
    def __init__(self, module):
        self.m = module
        self.L = lib()
        p0 = next(module.parameters())
        self.dev = p0.device
        self.world = 1
        self.rank = 0
        try:   # the reference's Dice all_reduce_sum runs whenever torch.distributed is initialised (ever/module/loss.py:20-23)
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        self._flatten_params()
        self._collect()
        if os.environ.get('EVB_BN_REDUCE_BLOCKS'):   # grid cap of the BatchNorm backward reduction, blocks per SM (1..4)
            self.L.evb_set_bn_reduce_blocks(c_int(int(os.environ['EVB_BN_REDUCE_BLOCKS'])))
        self.tape = []
        self.ws = None
        self.ws_bytes = 0
        self._wsmap, self._side_used = {}, False
        self._tape_tags, self._fork_idx, self._join_idx = [], None, None
        # pyramid levels 1..3 of the head run on their own streams (parallel graph branches; level 0 stays on main)
        # the critical-path streams (graph-capture stream, level branches) get a higher CUDA stream priority than the
        # weight-gradient side stream, so its CTAs fill the SMs the dgrad / BatchNorm chain leaves free instead of competing
        # for them (measured -0.7 % step time; EVB_PRIORITY=0 turns it off)
        self._prio = -1 if os.environ.get('EVB_PRIORITY', '1') == '1' else 0
        self.level_streams = ([torch.cuda.Stream(device=self.dev, priority=self._prio) for _ in range(3)]
                              if os.environ.get('EVB_NO_LEVEL_STREAMS', '0') != '1' else None)
        if self.fs_v2 and self.scene_shared:
            # a shared FSRelationV2 `project` BatchNorm updates its running statistics once per level: keep the level order
            self.level_streams = None
        # weight gradients run on a second stream (parallel graph branch): they overlap the dgrad / BN chain
        self.side = torch.cuda.Stream(device=self.dev) if os.environ.get('EVB_NO_SIDE_STREAM', '0') != '1' else None
        self.accumulate = False      # gradient accumulation into existing .grad (forward_times > 1)
        # Gradient buckets for the overlapped all-reduce (world > 1): backward is cut where these encoder stages start
        # (stage index 3 = layer4, 2 = layer3, 1 = layer2).  Bucket 0 = layer4 + head (75 % of the R50 parameters) is
        # all-reduced while layer3's backward runs, bucket 1 = layer3 while layer2 runs, ...; only the last ~1 MB (layer1 +
        # stem) is exposed.  Measured at 2 GPUs: 9.53 ms/step monolithic -> 9.37 (profiles/r02_allreduce_overlap.md).
        self.ar_split_stages = tuple(int(v) for v in os.environ.get('EVB_AR_SPLITS', '3,2,1').split(',') if v != '')
        self._tape_splits = []
        self._ar_pending = []
        self._reduced_in_graph = False
        self.fuse_bn_stats = True    # BN batch statistics in the producing conv's epilogue (evb_conv2d_fwd_stats)
        self._saved_for_backward = None
        self._graphs = {}
        self.debug = None            # dict -> named activations are recorded (tests / diagnostics)
        self.tf = None               # callable(kind, name, tensor) -> teacher forcing hook (tests; see _tf_fwd)
        self._read_config(module)

    def _read_config(self, module):
        cfg = module.config
        self.ignore_index = int(cfg.loss.ignore_index)
        self.ce_w = float(cfg.loss.ce.weight)
        self.dice_w = float(cfg.loss.dice.weight)
        self.smooth = float(cfg.loss.dice.smooth)
        self.sync_dice = bool(cfg.loss.dice.sync_statistics)
        self.K = module.head.fpn_decoder.num_classes
        if self.K > 64:
            raise NotImplementedError('more than 64 classes (the classifier runs as one 64-channel tensor-core tile)')
        # elements per pixel of the logit / logit-gradient rows [N, H, W, ld] (bf16): 16 covers K <= 16 with 32-byte rows
        self.ld = 16 if self.K <= 16 else 32 if self.K <= 32 else 64
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
        self._gv = {}
        self._slots = []   # (offset, numel) of every parameter in the arenas
        for p, o in zip(params, offs):
            self.flat_w[o:o + p.numel()].view_as(p).copy_(p.data)
            p.data = self.flat_w[o:o + p.numel()].view_as(p)
            gv = self.flat_g[o:o + p.numel()].view_as(p)
            self.grad_views.append(gv)
            self._gv[id(p)] = gv
            self._slots.append((o, p.numel()))
        self.n_params = n
        self._train_mask = None

    def _g(self, p):
        """the gradient slot of parameter p in the flat arena (None for a frozen parameter).  The engine always writes
        here; whether ``p.grad`` aliases the slot (native path, attach_grads) or autograd accumulates a copy of it into
        ``p.grad`` (autograd / DDP path, ever_b200.module._StepFn) is the caller's business."""
        return self._gv[id(p)] if p.requires_grad else None

    def trainable_mask(self):
        """uint8 mask over the arena slots: 1 where the slot belongs to a parameter with requires_grad (torch.optim.SGD
        skips parameters without a gradient: frozen weights must see neither weight decay nor momentum)."""
        key = tuple(p.requires_grad for p in self.params)
        if self._train_mask is None or self._train_mask[0] != key:
            if all(key):
                self._train_mask = (key, None)
            else:
                m = torch.zeros(self.flat_w.numel(), dtype=torch.uint8)
                for (o, nel), rg in zip(self._slots, key):
                    if rg:
                        m[o:o + nel] = 1
                self._train_mask = (key, m.to(self.dev))
        return self._train_mask[1]

    def accounting(self, on=True):
        """count the algorithmic bytes / FLOPs of the C-ABI calls made from now on (ever_b200.acct); returns the counter,
        accounting(False) restores the direct library binding"""
        from .acct import AcctLib
        if on:
            if not isinstance(self.L, AcctLib):
                self.L = AcctLib(self.L)
            return self.L
        if isinstance(self.L, AcctLib):
            acct, self.L = self.L, self.L._lib
            return acct
        return None

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

        self.bns, self.bns_padded = [], []

        def BNP_(bn, **kw):
            bp = BNP(bn, **kw)
            self.bns.append(bp)
            if bp.padded:
                self.bns_padded.append(bp)
            return bp
        self._mk_bn = BNP_

        def C(conv, **kw):
            cp = ConvP.from_conv(conv, **kw)
            self.convs.append(cp)
            return cp
        self.deep_stem = bool(getattr(r, 'deep_stem', False))
        if self.deep_stem:
            # v1c stem (_resnets.py:137-147): 3x3 s2 (Cin->32) as a GEMM over im2col rows, then 3x3 32->32 and 3x3 32->64
            # on the tensor-core kernel with the 32-channel tensors zero-padded to its 64-channel granularity
            st = r.stem
            self.stem_window = (3, 2, 1)
            self.stem_kp = _ceil(st[0].in_channels * 9, 64)
            self.stem = C(st[0], as_matrix=True, cout_pad=64)
            self.stem_bn = BNP_(st[1], cpad=64)
            c2 = ConvP(st[3].weight, None, 1, cout_pad=64, cin_pad=64)
            c3 = ConvP(st[6].weight, None, 1, cin_pad=64)
            self.convs += [c2, c3]
            self.stem_tail = [(c2, BNP_(st[4], cpad=64)), (c3, BNP_(st[7]))]
        else:
            self.stem_window = (7, 2, 3)
            self.stem_kp = _ceil(r.conv1.in_channels * 49, 64)
            self.stem = C(r.conv1, as_matrix=True)   # [64][Cin*49] GEMM; Cin*49 padded to stem_kp
            self.stem_bn = BNP_(r.bn1)
            self.stem_tail = []
        self.stem.need_dgrad = False
        self.stages = []
        for li in range(1, 5):
            blocks = []
            for b in getattr(r, 'layer%d' % li):
                d = dict(stride=b.stride, c1=C(b.conv1), b1=BNP_(b.bn1), c2=C(b.conv2), b2=BNP_(b.bn2))
                if self.kind == 'bottleneck':
                    d.update(c3=C(b.conv3), b3=BNP_(b.bn3))
                if b.downsample is not None:
                    d.update(cd=C(b.downsample[0]), bd=BNP_(b.downsample[1]))
                blocks.append(d)
            self.stages.append(blocks)
        h = m.head
        self.fpn_inner = [C(getattr(h.fpn, 'fpn_inner%d' % i)[0]) for i in range(1, 5)]
        self.fpn_layer = [C(getattr(h.fpn, 'fpn_layer%d' % i)[0]) for i in range(1, 5)]
        fs = h.fs_relation
        # scene MLPs (conv 2048->256, conv 256->256) as linears: one per level, or one shared (scale_aware_proj=False)
        self.scene_shared = not getattr(fs, 'scale_aware_proj', True)
        self.scene = ([(fs.scene_encoder[0], fs.scene_encoder[2])] if self.scene_shared
                      else [(s[0], s[2]) for s in fs.scene_encoder])
        self.content = [(C(s[0]), BNP_(s[1])) for s in fs.content_encoders]
        self.reenc = [(C(s[0]), BNP_(s[1])) for s in fs.feature_reencoders]
        # FSRelationV2 (fs_relation.py:76-163): GroupNorm scene encoder, `project` conv on cat([r * p, p]), Dropout2d
        self.fs_v2 = getattr(fs, 'version', 1) == 2
        self.project, self.drop_p = None, 0.0
        if self.fs_v2:
            encs = [fs.scene_encoder] if self.scene_shared else list(fs.scene_encoder)
            self.scene = [(s[0], s[1], s[3], s[4], s[5]) for s in encs]     # conv, GN, conv, GN, final ReLU (its name)
            projs = [fs.project] if self.scene_shared else list(fs.project)
            pj = [(C(s[0]), BNP_(s[1]), s[3]) for s in projs]
            self.project = pj * 4 if self.scene_shared else pj
            self.drop_p = float(fs.dropout_p)
        dec = h.fpn_decoder
        self.dec_blocks = [[(C(layer[0]), BNP_(layer[1])) for layer in blk] for blk in dec.blocks]
        self.dec_nup = list(dec.num_upsample)
        self.cls = C(dec.classifier[0], cout_pad=64)
        self.cls_scale = dec.scale_factor
        self._collect_extra()
        # reference module paths (== state_dict prefixes) of every conv / BN the engine schedules
        wname = {id(mod.weight): path for path, mod in m.named_modules()
                 if isinstance(getattr(mod, 'weight', None), torch.nn.Parameter)}
        for cp in self.convs:
            if not cp.w_ld:
                cp.name = wname.get(id(cp.weight))
        for bp in self.bns:
            bp.name = wname.get(id(bp.bn.weight))
        self._alloc_packs()

    def _collect_extra(self):
        """hook for derived engines (extra heads): append their ConvP objects to self.convs here"""

    def _alloc_packs(self):
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
        # padded convs: dense scratch for the weight gradient, padded bias + bias-gradient buffers
        for cp in self.convs:
            if not cp.direct_grad:
                cp.gscr = torch.zeros(cp.cop * cp.cip * cp.kk, dtype=torch.float32, device=self.dev)
            if cp.groups > 1:
                if cp.cop != cp.co or cp.cip != cp.ci or cp.w_ld:
                    raise NotImplementedError('grouped convolution with padded channels')
                cp.dense_w = torch.zeros(cp.co * cp.ci * cp.kk, dtype=torch.float32, device=self.dev)
            if cp.bias is not None and cp.cop != cp.co:
                cp.bias_pad = torch.zeros(cp.cop, dtype=torch.float32, device=self.dev)
                cp.dbias_scr = torch.zeros(cp.cop, dtype=torch.float32, device=self.dev)

    def _build_pack_table(self):
        rows, bmap, nblk = [], [], 0
        for i, cp in enumerate(self.convs):
            kk = cp.k * cp.k
            nb = ((cp.co + 63) // 64) * ((cp.ci + 63) // 64)   # one block per 64x64 (co, ci) tile, all taps
            src = cp.dense_w.data_ptr() if cp.groups > 1 else cp.weight.data_ptr() + 4 * cp.w_off
            rows.append([src, cp.wf.data_ptr(), cp.wb.data_ptr() if cp.need_dgrad else 0,
                         cp.co, cp.ci, kk, cp.cop, cp.cip, cp.cip, cp.cop, nblk, cp.w_ld])
            bmap += [i] * nb
            nblk += nb
        self._pack_desc = torch.tensor(rows, dtype=torch.int64, device=self.dev)
        self._pack_map = torch.tensor(bmap, dtype=torch.int32, device=self.dev)
        self._pack_nblk = nblk
        # optional (EVB_PACK_OVERLAP=1; measured neutral on B200, off by default): blocks of the convolutions the forward
        # pass needs first (stem, layer1, layer2) are packed on the main stream, the rest (layer3, layer4, head: ~95 % of
        # the parameters) on the side stream while those layers run
        stages = getattr(self, 'stages', ())
        first_late = stages[2][0]['c1'] if len(stages) > 2 else None
        self._pack_split = nblk
        if first_late is not None and self.side is not None and os.environ.get('EVB_PACK_OVERLAP', '0') == '1':
            self._pack_split = rows[self.convs.index(first_late)][10]
        self._pack_ptrs = [cp.weight.data_ptr() for cp in self.convs]

    def pack_weights(self):
        """fp32 master -> bf16 packs for every convolution, one launch (once per step, after the optimizer)."""
        st = stream()
        if getattr(self, '_pack_desc', None) is None or self._pack_ptrs != [cp.weight.data_ptr() for cp in self.convs]:
            self._build_pack_table()
        self._pack_ev = None
        for cp in self.convs:
            if cp.groups > 1:
                check(self.L.evb_group_expand(ptr(cp.weight), ptr(cp.dense_w), c_int(cp.co), c_int(cp.ci), c_int(cp.kk),
                                              c_int(cp.groups), st), 'evb_group_expand')
        if self._pack_split < self._pack_nblk:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                check(self.L.evb_pack_weights_tiled(ptr(self._pack_desc), ptr(self._pack_map), c_int(self._pack_split),
                                                    c_int(self._pack_nblk - self._pack_split), stream()),
                      'evb_pack_weights_tiled')
                self._pack_ev = torch.cuda.Event()
                self._pack_ev.record(self.side)
        check(self.L.evb_pack_weights_tiled(ptr(self._pack_desc), ptr(self._pack_map), c_int(0), c_int(self._pack_split), st),
              'evb_pack_weights_tiled')
        if hasattr(self.L, 'add'):   # accounting: read the fp32 master once, write both bf16 layouts
            self.L.add('evb_pack_weights_tiled', sum(cp.co * cp.ci * cp.kk * (4 + 2 + (2 if cp.need_dgrad else 0))
                                                     for cp in self.convs))
        for cp in self.convs:
            cp._gw = False
            if cp.bias_pad is not None:
                check(self.L.evb_copy2d_f32(ptr(cp.bias), c_int(cp.co), ptr(cp.bias_pad), c_int(cp.cop), c_int(1),
                                            c_int(cp.co), c_int(0), st), 'evb_copy2d_f32')
        for bp in self.bns:
            bp._gw = False
        for bp in self.bns_padded:   # refresh the zero-padded shadow BN parameters
            bn = bp.bn
            for src, dst in ((bn.weight, bp.gamma_p), (bn.bias, bp.beta_p), (bn.running_mean, bp.rm_p),
                             (bn.running_var, bp.rv_p)):
                check(self.L.evb_copy2d_f32(ptr(src), c_int(bp.c_real), ptr(dst), c_int(bp.c), c_int(1), c_int(bp.c_real),
                                            c_int(0), st), 'evb_copy2d_f32')

    # ------------------------------------------------------------------ workspace
    def _ws(self, nbytes):
        """scratch workspace of the stream the caller is launching on (main, a pyramid-level branch stream or the
        weight-gradient side stream): kernels on different streams never share scratch"""
        key = torch.cuda.current_stream().cuda_stream
        ent = self._wsmap.get(key)
        if ent is None or ent[1] < nbytes:
            nb = max(int(nbytes), 32 << 20)
            ent = (torch.empty(nb // 4, dtype=torch.float32, device=self.dev), nb)
            self._wsmap[key] = ent
        return ent[0]

    def _ws_cap(self):
        ent = self._wsmap.get(torch.cuda.current_stream().cuda_stream)
        return ent[1] if ent is not None else 0

    def _param_grads_async(self, tensors, fn):
        """Run fn() (weight / bias gradient kernels: nothing downstream in backward reads their outputs) on the side
        stream, forked from the current point of the main stream; backward() joins before the optimizer.  `tensors`
        are kept alive for the side stream (caching-allocator stream bookkeeping)."""
        if self.side is None:
            fn()
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        # no record_stream needed: every tensor the side stream reads is referenced by the tape until backward() has
        # joined the side stream (see backward(): join first, then drop the tape)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            fn()
        self._side_used = True

    def _join_pack(self):
        if getattr(self, '_pack_ev', None) is not None:
            torch.cuda.current_stream().wait_event(self._pack_ev)
            self._pack_ev = None

    def _join_side(self):
        if self.side is not None and self._side_used:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_used = False

    def _dbg(self, name, t):
        if self.debug is not None:
            self.debug[name] = t.data if isinstance(t, Act) else t

    def _tf_fwd(self, act, name=None):
        """teacher forcing (parity tests): hand the freshly computed activation to ``self.tf('fwd', name, tensor)``, which
        compares it with the reference's tensor of that module path and overwrites it, so that every op is checked on the
        reference's own inputs (no error amplification through the ~50-layer ReLU/BN chain)."""
        if name is not None:
            act.name = name
        if self.tf is not None and act.name is not None:
            self.tf('fwd', act.name, act.data)
        return act

    def _tf_bwd(self, act):
        """same for the completed gradient of an activation, at the top of the backward closure of its producer"""
        if self.tf is not None and act.name is not None and act.grad is not None:
            self.tf('bwd', act.name, act.grad)

    def _new(self, *shape, dtype=BF16):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _zero(self, t):
        """t[:] = 0 on the current stream as a memset (a memset node of the captured step, not a fill kernel)"""
        assert t.is_contiguous()
        check(self.L.evb_zero_bytes(ptr(t), c_ll(t.numel() * t.element_size()), stream()), 'evb_zero_bytes')
        return t

    # ------------------------------------------------------------------ ops (forward + tape)
    def _grad_into(self, act, shape=None):
        """Return (buffer, accumulate_flag) for writing a gradient contribution of `act`."""
        if act.grad is None:
            act.grad = self._new(*act.data.shape)
        acc = act.has_grad
        act.has_grad = True
        return act.grad, acc

    def _wgrad(self, cp, x_data, dy, n, h, w, cin, cout, stride):
        """weight gradient of one conv use; shared convs (used twice in a step) accumulate."""
        L = self.L
        st = stream()
        if self._g(cp.weight) is None:
            return
        ho, wo = h // stride, w // stride
        nbytes = L.evb_conv2d_wgrad_workspace(c_int(n), c_int(ho), c_int(wo), c_int(cin), c_int(cout), c_int(cp.k), c_int(0),
                                              c_int(0))
        ws = self._ws(nbytes)
        acc = self.accumulate or cp._gw
        target = self._g(cp.weight) if cp.direct_grad else cp.gscr
        check(L.evb_conv2d_wgrad(ptr(x_data), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(dy), c_int(cout), c_int(cp.k),
                                 c_int(stride), ptr(target), c_int(1 if (acc and cp.direct_grad) else 0), ptr(ws),
                                 c_ll(self._ws_cap()), c_int(0), c_int(0), st), 'evb_conv2d_wgrad')
        if cp.groups > 1:        # the diagonal blocks of the dense gradient are the grouped conv's [co][ci/g][k][k]
            check(L.evb_group_extract(ptr(cp.gscr), ptr(self._g(cp.weight)), c_int(cp.co), c_int(cp.ci), c_int(cp.kk),
                                      c_int(cp.groups), c_int(1 if acc else 0), st), 'evb_group_extract')
        elif not cp.direct_grad:   # copy the valid [co][ci*kk] block of the dense scratch [cop][cip*kk] into the OIHW grad
            dst = ctypes.c_void_p(self._g(cp.weight).data_ptr() + 4 * cp.w_off)
            check(L.evb_copy2d_f32(ptr(cp.gscr), c_int(cp.cip * cp.kk), dst, c_int(cp.w_ld or cp.ci * cp.kk), c_int(cp.co),
                                   c_int(cp.ci * cp.kk), c_int(1 if acc else 0), st), 'evb_copy2d_f32')

    def _bias_grad(self, cp, dy, m_rows, cout):
        L = self.L
        st = stream()
        if cp.bias is None or self._g(cp.bias) is None:
            return
        acc = self.accumulate or cp._gw
        ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(cout)))
        if cp.bias_pad is not None:
            check(L.evb_bias_grad(ptr(dy), c_ll(m_rows), c_int(cout), ptr(cp.dbias_scr), None, c_int(0), ptr(ws), st),
                  'evb_bias_grad')
            check(L.evb_copy2d_f32(ptr(cp.dbias_scr), c_int(cp.cop), ptr(self._g(cp.bias)), c_int(cp.co), c_int(1), c_int(cp.co),
                                   c_int(1 if acc else 0), st), 'evb_copy2d_f32')
        else:
            check(L.evb_bias_grad(ptr(dy), c_ll(m_rows), c_int(cout), ptr(self._g(cp.bias)), None, c_int(1 if acc else 0), ptr(ws),
                                  st), 'evb_bias_grad')

    def conv(self, x, cp, stride=None, bias=False, add=None, add_mode=0, train=True, dgrad=True, stats=False,
             bias_grad_zero=False, name=None):
        """y = conv(x) (+bias) (+add).  bias=True uses cp.bias (padded copy when the conv is channel-padded).
        stats=True (training, no add): the epilogue also emits the BN batch-statistic partial sums of y.
        bias_grad_zero: the conv feeds a training-mode BatchNorm directly, so d(loss)/d(bias) = sum_rows dx_BN is
        identically zero (dx = a*g + k0 - c2*x sums to a*sum(g) + M*k0 - c2*M*mean = 0 by the definition of k0): the
        gradient is written as exact zeros instead of summing rounding noise over the rows (the reference's value there is
        ~1e-8 of the other gradients)."""
        L = self.L
        stride = cp.stride if stride is None else stride
        n, h, w, cin = x.data.shape
        ho, wo = h // stride, w // stride
        cout = cp.cop
        bias_t = (cp.bias_pad if cp.bias_pad is not None else cp.bias) if bias else None
        y = Act(self._new(n, ho, wo, cout))
        if stats and self.fuse_bn_stats and add is None:
            partial = self._new(2 * cout * 320, dtype=torch.float32)
            nblk = ctypes.c_int(0)
            if bias_t is None:
                check(L.evb_conv2d_fwd_stats(ptr(x.data), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(cp.wf),
                                             c_int(cp.cop), c_int(cp.k), c_int(stride), ptr(y.data), c_int(cout),
                                             ptr(partial), ctypes.byref(nblk), stream()), 'evb_conv2d_fwd_stats')
            else:
                check(L.evb_conv2d_fwd_bias_stats(ptr(x.data), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(cp.wf),
                                                  c_int(cp.cop), c_int(cp.k), c_int(stride), ptr(y.data), c_int(cout),
                                                  ptr(bias_t), ptr(partial), ctypes.byref(nblk), stream()),
                      'evb_conv2d_fwd_bias_stats')
            y.stats = (partial, nblk.value)
        else:
            check(L.evb_conv2d_fwd(ptr(x.data), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(cp.wf), c_int(cp.cop),
                                   c_int(cp.k), c_int(stride), ptr(y.data), c_int(cout), ptr(bias_t),
                                   ptr(add.data if add is not None else None), c_int(add_mode), c_int(0), stream()),
                  'evb_conv2d_fwd')
        self._tf_fwd(y, name or cp.name)
        if train:
            def bwd():
                if y.grad is None:
                    return
                self._tf_bwd(y)
                dy = y.grad
                st = stream()
                def param_grads():
                    self._wgrad(cp, x.data, dy, n, h, w, cin, cout, stride)
                    if bias and bias_grad_zero and self._g(cp.bias) is not None:
                        if not (self.accumulate or cp._gw):
                            self._zero(self._g(cp.bias))
                    elif bias:
                        self._bias_grad(cp, dy, n * ho * wo, cout)
                    cp._gw = True
                self._param_grads_async((x.data, dy), param_grads)
                if add is not None and add.needs_grad:
                    g, acc2 = self._grad_into(add)
                    if add_mode == 2:
                        check(L.evb_sumpool2(ptr(dy), ptr(g), c_int(n), c_int(ho // 2), c_int(wo // 2), c_int(cout),
                                             c_int(1 if acc2 else 0), st), 'evb_sumpool2')
                    else:
                        check(L.evb_scale_add(ptr(dy), c_float(1.0), ptr(g if acc2 else None), ptr(g),
                                              c_ll(dy.numel()), st), 'evb_scale_add')
                if x.needs_grad and cp.need_dgrad and dgrad:
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
        sync = bp.sync and self.world > 1 and train and bn.training
        if train and bn.training and x.data.numel() // c <= 1 and not sync:
            # F.batch_norm's own check (torch/nn/functional.py _verify_batch_size): the reference fails the same way on a
            # 1 x 32 x 32 tile, whose c5 is a single pixel
            n_, h_, w_, _ = x.data.shape
            raise ValueError('Expected more than 1 value per channel when training, got input size %s'
                             % str(torch.Size([n_, bp.c_real, h_, w_])))
        if sync and x.stats is None:
            raise NotImplementedError('SyncBatchNorm behind a convolution with a fused add (no epilogue statistics)')
        if sync:
            import torch.distributed as dist
            m_rows = x.data.numel() // c
            mom = 0.1 if bn.momentum is None else bn.momentum
            partial, nblk = x.stats
            sums = self._new(2, c, dtype=torch.float32)
            check(L.evb_bn_partial_sums(ptr(partial), c_int(nblk), c_int(c), ptr(sums), stream()), 'evb_bn_partial_sums')
            dist.all_reduce(sums)     # every rank contributes the same number of rows (equal per-GPU batches under DDP)
            check(L.evb_bn_finalize_sums(ptr(sums), ctypes.c_double(float(m_rows) * self.world), c_int(c), ptr(bp.gamma),
                                         ptr(bp.beta), ptr(bp.rm), ptr(bp.rv), c_float(mom), c_float(bn.eps), ptr(mean),
                                         ptr(rstd), ptr(scale), ptr(shift), stream()), 'evb_bn_finalize_sums')
            if bp.padded:
                for src, dst in ((bp.rm_p, bn.running_mean), (bp.rv_p, bn.running_var)):
                    check(L.evb_copy2d_f32(ptr(src), c_int(bp.c), ptr(dst), c_int(bp.c_real), c_int(1), c_int(bp.c_real),
                                           c_int(0), stream()), 'evb_copy2d_f32')
            self._bn_tracked.append(bn)
        elif train and bn.training and x.stats is not None:
            m_rows = x.data.numel() // c
            mom = 0.1 if bn.momentum is None else bn.momentum
            partial, nblk = x.stats
            check(L.evb_bn_finalize(ptr(partial), c_int(nblk), c_ll(m_rows), c_int(c), ptr(bp.gamma), ptr(bp.beta), ptr(bp.rm),
                                    ptr(bp.rv), c_float(mom), c_float(bn.eps), ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
                                    stream()), 'evb_bn_finalize')
            if bp.padded:
                for src, dst in ((bp.rm_p, bn.running_mean), (bp.rv_p, bn.running_var)):
                    check(L.evb_copy2d_f32(ptr(src), c_int(bp.c), ptr(dst), c_int(bp.c_real), c_int(1), c_int(bp.c_real),
                                           c_int(0), stream()), 'evb_copy2d_f32')
            self._bn_tracked.append(bn)
        elif train and bn.training:
            m_rows = x.data.numel() // c
            ws = self._ws(L.evb_bn_workspace(c_ll(m_rows), c_int(c)))
            mom = 0.1 if bn.momentum is None else bn.momentum
            check(L.evb_bn_stats(ptr(x.data), c_ll(m_rows), c_int(c), ptr(bp.gamma), ptr(bp.beta), ptr(bp.rm), ptr(bp.rv),
                                 c_float(mom), c_float(bn.eps), ptr(mean), ptr(rstd), ptr(scale), ptr(shift), ptr(ws),
                                 stream()), 'evb_bn_stats')
            if bp.padded:   # running statistics back into the module's (unpadded) buffers
                for src, dst in ((bp.rm_p, bn.running_mean), (bp.rv_p, bn.running_var)):
                    check(L.evb_copy2d_f32(ptr(src), c_int(bp.c), ptr(dst), c_int(bp.c_real), c_int(1), c_int(bp.c_real),
                                           c_int(0), stream()), 'evb_copy2d_f32')
            self._bn_tracked.append(bn)
        else:
            check(L.evb_bn_fold(ptr(bp.gamma), ptr(bp.beta), ptr(bp.rm), ptr(bp.rv), c_float(bn.eps), c_int(c), ptr(scale),
                                ptr(shift), ptr(mean), ptr(rstd), stream()), 'evb_bn_fold')
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
        acc = self.accumulate or bp._gw
        dgam = bp.dgamma_p if bp.padded else self._g(bp.bn.weight)
        dbet = bp.dbeta_p if bp.padded else self._g(bp.bn.bias)
        if bp.sync and self.world > 1 and bp.bn.training:
            # SyncBatchNorm backward (torch/nn/modules/_functions.py SyncBatchNorm.backward): the parameter gradients are this
            # rank's sums (DDP / the arena all-reduce average them like every other gradient); d(input) uses the sums of ALL
            # ranks and the global row count
            import torch.distributed as dist
            fresh = self._new(2, c, dtype=torch.float32)
            check(L.evb_norm_bwd_reduce(ptr(dy), ptr(x.data), ptr(ymask), ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
                                        c_int(mask_mode), ptr(fresh[0]), ptr(fresh[1]), c_int(0), c_ll(m_rows), c_int(c),
                                        ptr(ws), stream()), 'evb_norm_bwd_reduce')
            for src, dst in ((fresh[0], dgam), (fresh[1], dbet)):
                if dst is not None:
                    check(L.evb_copy2d_f32(ptr(src), c_int(c), ptr(dst), c_int(c), c_int(1), c_int(c),
                                           c_int(1 if (acc and not bp.padded) else 0), stream()), 'evb_copy2d_f32')
            glob = fresh.clone()
            dist.all_reduce(glob)
            k = self._new(2, c, dtype=torch.float32)
            check(L.evb_bn_bwd_consts(ptr(scale), ptr(rstd), ptr(mean), ptr(glob[0]), ptr(glob[1]),
                                      c_float(1.0 / (float(m_rows) * self.world)), c_int(c), ptr(k[0]), ptr(k[1]), stream()),
                  'evb_bn_bwd_consts')
            check(L.evb_norm_bwd_apply(ptr(dy), ptr(x.data), ptr(ymask), ptr(scale), ptr(shift), ptr(k[0]), ptr(k[1]),
                                       c_int(mask_mode), ptr(gx), ptr(dres), c_int(1 if dres_acc else 0), c_ll(m_rows), c_int(c),
                                       stream()), 'evb_norm_bwd_apply')
        else:
            self._bn_bwd_local(dy, x, ymask, fold, mask_mode, bp, gx, dres, dres_acc, dgam, dbet, acc, m_rows, c, ws)
        if bp.padded and self._g(bp.bn.weight) is not None:
            for src, dst in ((bp.dgamma_p, self._g(bp.bn.weight)), (bp.dbeta_p, self._g(bp.bn.bias))):
                check(L.evb_copy2d_f32(ptr(src), c_int(bp.c), ptr(dst), c_int(bp.c_real), c_int(1), c_int(bp.c_real),
                                       c_int(1 if acc else 0), stream()), 'evb_copy2d_f32')
        bp._gw = True

    def _bn_bwd_local(self, dy, x, ymask, fold, mask_mode, bp, gx, dres, dres_acc, dgam, dbet, acc, m_rows, c, ws):
        L = self.L
        mean, rstd, scale, shift = fold
        check(L.evb_bn_bwd(ptr(dy), ptr(x.data), ptr(ymask), ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
                           c_int(mask_mode), c_int(0 if bp.bn.training else 1), ptr(gx), ptr(dres), c_int(1 if dres_acc else 0),
                           ptr(dgam), ptr(dbet), c_int(1 if (acc and not bp.padded) else 0),
                           c_ll(m_rows), c_int(c), ptr(ws), stream()), 'evb_bn_bwd')

    def bn_act(self, x, bp, relu=True, res=None, train=True, name=None):
        L = self.L
        fold = self._bn_fold(x, bp, train)
        y = Act(self._new(*x.data.shape))
        c = bp.c
        # block tail relu(bn(x) + identity): the backward needs the ReLU survivors; one uint32 of bits per 8-channel vector
        # (a quarter of y's bytes) is written here instead of re-reading y twice there.  Not under teacher forcing (the
        # output is overwritten by the reference's tensor, whose mask must be used) nor for SyncBatchNorm's split backward.
        bits = None
        if train and relu and res is not None and self.tf is None and not (bp.sync and self.world > 1):
            bits = self._new(x.data.numel() // 8, dtype=torch.int32)
            check(L.evb_bn_apply_mask(ptr(x.data), ptr(fold[2]), ptr(fold[3]), ptr(res.data), ptr(y.data), ptr(bits),
                                      c_ll(x.data.numel() // c), c_int(c), stream()), 'evb_bn_apply_mask')
        else:
            check(L.evb_bn_apply(ptr(x.data), ptr(fold[2]), ptr(fold[3]), ptr(res.data if res is not None else None),
                                 ptr(y.data), c_ll(x.data.numel() // c), c_int(c), c_int(1 if relu else 0), stream()),
                  'evb_bn_apply')
        self._tf_fwd(y, name)
        if train:
            def bwd():
                if y.grad is None:
                    return
                self._tf_bwd(y)
                if relu and res is None:   # mask recomputed from x: one tensor read less per pass
                    self._bn_backward(y.grad, x, bp, fold, 2, None, None)
                elif bits is not None:
                    self._bn_backward(y.grad, x, bp, fold, 3, bits, res)
                else:
                    self._bn_backward(y.grad, x, bp, fold, 1 if relu else 0, y.data if relu else None, res)
            self.tape.append(bwd)
        return y

    def bn_relu_up(self, x, bp, f=2, train=True, name=None):
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
        self._tf_fwd(y, name)
        if train:
            def bwd():
                if y.grad is None:
                    return
                self._tf_bwd(y)
                low = self._new(n, h, w, c)
                ws = self._ws(L.evb_bilinear_up_bwd_workspace(c_int(n), c_int(h), c_int(w), c_int(c), c_int(f)))
                check(L.evb_bilinear_up_bwd_sep(ptr(y.grad), ptr(low), c_int(n), c_int(h), c_int(w), c_int(c), c_int(c),
                                                c_int(c), c_int(f), ptr(ws), c_ll(self._ws_cap()), stream()),
                      'evb_bilinear_up_bwd_sep')
                self._bn_backward(low, x, bp, fold, 2, None, None)
            self.tape.append(bwd)
        return y

    def maxpool(self, x, train=True, name=None):
        L = self.L
        n, h, w, c = x.data.shape
        y = Act(self._new(n, h // 2, w // 2, c))
        idx = self._new(n, h // 2, w // 2, c, dtype=torch.uint8)
        check(L.evb_maxpool3x3s2_fwd(ptr(x.data), ptr(y.data), ptr(idx), c_int(n), c_int(h), c_int(w), c_int(c), stream()),
              'evb_maxpool3x3s2_fwd')
        self._tf_fwd(y, name)
        if train:
            def bwd():
                if y.grad is None or not x.needs_grad:
                    return
                self._tf_bwd(y)
                g, acc = self._grad_into(x)
                assert not acc
                check(L.evb_maxpool3x3s2_bwd(ptr(y.grad), ptr(idx), ptr(g), c_int(n), c_int(h), c_int(w), c_int(c),
                                             stream()), 'evb_maxpool3x3s2_bwd')
            self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ network
    def _block(self, x, d, train):
        # activation names: the block's shared nn.ReLU is called 2 (BasicBlock) or 3 (Bottleneck) times: '<block>.relu#k'
        pfx = d['c1'].name.rsplit('.', 1)[0]
        if self.kind == 'bottleneck':
            a1 = self.bn_act(self.conv(x, d['c1'], train=train, stats=train), d['b1'], True, train=train, name=pfx + '.relu#0')
            a2 = self.bn_act(self.conv(a1, d['c2'], train=train, stats=train), d['b2'], True, train=train, name=pfx + '.relu#1')
            o3 = self.conv(a2, d['c3'], train=train, stats=train)
            last_bn, last = d['b3'], pfx + '.relu#2'
        else:
            a1 = self.bn_act(self.conv(x, d['c1'], train=train, stats=train), d['b1'], True, train=train, name=pfx + '.relu#0')
            o3 = self.conv(a1, d['c2'], train=train, stats=train)
            last_bn, last = d['b2'], pfx + '.relu#1'
        idt = x
        if 'cd' in d:
            idt = self.bn_act(self.conv(x, d['cd'], train=train, stats=train), d['bd'], False, train=train, name=d['bd'].name)
        return self.bn_act(o3, last_bn, True, res=idt, train=train, name=last)

    def _encoder(self, x_nchw, train):
        L = self.L
        u8 = x_nchw.dtype == torch.uint8   # raw HWC uint8 tiles [N,H,W,Cin]: normalisation fused into the im2col
        if u8:
            n, h, w, cin = x_nchw.shape
        else:
            n, cin, h, w = x_nchw.shape
        if h % 32 or w % 32:
            raise ValueError('FarSegB200 needs H, W divisible by 32 (FPN nearest-x2 adds, SURVEY.md section 5)')
        a = self._new(n, h // 2, w // 2, self.stem_kp)
        ks, sstride, spad = self.stem_window
        if u8:
            if getattr(self, '_in_mean', None) is None:
                icfg = self.m.config.input
                self._in_mean = torch.tensor(list(icfg.mean), dtype=torch.float32, device=self.dev)
                self._in_std = torch.tensor(list(icfg.std), dtype=torch.float32, device=self.dev)
            check(L.evb_im2col_u8(ptr(x_nchw), ptr(self._in_mean), ptr(self._in_std), ptr(a), c_int(n), c_int(cin),
                                  c_int(h), c_int(w), c_int(self.stem_kp), c_int(ks), c_int(sstride), c_int(spad), stream()),
                  'evb_im2col_u8')
        else:
            check(L.evb_im2col_nchw(ptr(x_nchw), ptr(a), c_int(n), c_int(cin), c_int(h), c_int(w), c_int(self.stem_kp),
                                    c_int(ks), c_int(sstride), c_int(spad), stream()), 'evb_im2col_nchw')
        xa = Act(a, needs_grad=False)
        y0 = self.conv(xa, self.stem, stride=1, train=False, stats=train)   # distinct name: the closure below keeps THIS Act
        if train:
            stem, ho, wo = self.stem, h // 2, w // 2

            def stem_bwd():
                if y0.grad is None:
                    return
                self._tf_bwd(y0)
                def param_grads():
                    self._wgrad(stem, a, y0.grad, n, ho, wo, self.stem_kp, 64, 1)
                    stem._gw = True
                self._param_grads_async((a, y0.grad), param_grads)
            self.tape.append(stem_bwd)
        y = y0
        self._dbg('stem_conv', y)
        # names: the trunk's nn.ReLU ('en.resnet.relu', called once) or the ReLU after each deep-stem BN ('...stem.2/5/8')
        rpfx = self.stem.name.rsplit('.', 1)[0] if self.stem.name else None   # 'en.resnet' or 'en.resnet.stem'

        def _relu_name(bp_):
            if bp_.name is None or rpfx is None:
                return None
            if not self.deep_stem:
                return rpfx + '.relu#0'
            head_, idx = bp_.name.rsplit('.', 1)
            return '%s.%d' % (head_, int(idx) + 1)
        y = self.bn_act(y, self.stem_bn, True, train=train, name=_relu_name(self.stem_bn))
        for cp_, bp_ in self.stem_tail:   # deep stem: two more 3x3 conv + BN + ReLU
            y = self.bn_act(self.conv(y, cp_, train=train, stats=train), bp_, True, train=train, name=_relu_name(bp_))
        self._dbg('stem_act', y)
        y = self.maxpool(y, train=train, name=(rpfx[:-5] if self.deep_stem else rpfx) + '.maxpool' if rpfx else None)
        self._dbg('pool', y)
        feats = []
        freeze_at = int(self.m.config.encoder.freeze_at)
        for si, blocks in enumerate(self.stages):
            if train and si in self.ar_split_stages:
                self._tape_splits.append(len(self.tape))   # backward segment boundary (gradient bucket, see backward())
            if si == 2:
                self._join_pack()      # layer3 onwards reads the packs written on the side stream
            if train and freeze_at >= si + 1:
                y.needs_grad = False   # everything that produced y is frozen: no gradient flows further down
            for d in blocks:
                y = self._block(y, d, train)
            if train and freeze_at >= si + 2:
                y.needs_grad = False   # this stage and everything below it are frozen
            feats.append(y)
            self._dbg('c%d' % (si + 2), y)
        self._join_pack()
        return feats

    def _scene_mlp_one(self, scene, n, l1, l2, train, dscene, extra=()):
        """one scene MLP (1x1 conv -> ReLU -> 1x1 conv on the N x C5 x 1 x 1 scene embedding, fs_relation.py:22-35) as two
        tiny linears on the caller's stream.  Backward: d(sf) (+ the `extra` d(sf) buffers of other levels, summed in
        order) -> both linears' parameter gradients and this MLP's own d(scene) buffer `dscene` (overwritten)."""
        L = self.L
        c5 = scene.shape[1]
        co = l1.out_channels
        hid = self._new(n, co, dtype=torch.float32)
        sf = self._new(n, co, dtype=torch.float32)
        check(L.evb_linear_fwd(ptr(scene), ptr(l1.weight), ptr(l1.bias), ptr(hid), c_int(n), c_int(c5), c_int(co),
                               c_int(1), stream()), 'evb_linear_fwd')
        check(L.evb_linear_fwd(ptr(hid), ptr(l2.weight), ptr(l2.bias), ptr(sf), c_int(n), c_int(co), c_int(co),
                               c_int(0), stream()), 'evb_linear_fwd')
        dsf = self._new(n, co, dtype=torch.float32) if train else None
        sf_name = None
        if self.tf is not None:
            sf_name = {id(mod): path for path, mod in self.m.named_modules()}.get(id(l2))
            self.tf('fwd', sf_name, sf)
        if train:
            def bwd():
                acc = 1 if self.accumulate else 0
                for d in extra:
                    check(L.evb_copy2d_f32(ptr(d), c_int(co), ptr(dsf), c_int(co), c_int(n), c_int(co), c_int(1),
                                           stream()), 'evb_copy2d_f32')
                if self.tf is not None:
                    self.tf('bwd', sf_name, dsf)
                dhid = self._new(n, co, dtype=torch.float32)
                check(L.evb_linear_bwd(ptr(dsf), ptr(sf), ptr(hid), ptr(l2.weight), ptr(self._g(l2.weight)),
                                       ptr(self._g(l2.bias)), ptr(dhid), c_int(n), c_int(co), c_int(co), c_int(0),
                                       c_int(acc), c_int(0), stream()), 'evb_linear_bwd')
                check(L.evb_linear_bwd(ptr(dhid), ptr(hid), ptr(scene), ptr(l1.weight), ptr(self._g(l1.weight)),
                                       ptr(self._g(l1.bias)), ptr(dscene), c_int(n), c_int(c5), c_int(co), c_int(1),
                                       c_int(acc), c_int(0), stream()), 'evb_linear_bwd')
            self.tape.append(bwd)
        return sf, dsf

    def _scene_mlp_v2(self, scene, n, mods, train, dscene, extra=()):
        """FSRelationV2 scene encoder (fs_relation.py:86-96): 1x1 conv -> GroupNorm(32) -> ReLU, twice, on the N x C5 scene
        vector; the second ReLU output stays fp32 (autocast runs group_norm in fp32), the first is consumed by a bf16 conv."""
        L = self.L
        l1, g1, l2, g2, last = mods
        c5, co, G = scene.shape[1], l1.out_channels, g1.num_groups
        h1, a1, h2, sf = (self._new(n, co, dtype=torch.float32) for _ in range(4))
        st1, st2 = self._new(n, G, 2, dtype=torch.float32), self._new(n, G, 2, dtype=torch.float32)
        st = stream()
        check(L.evb_linear_fwd(ptr(scene), ptr(l1.weight), ptr(l1.bias), ptr(h1), c_int(n), c_int(c5), c_int(co), c_int(0), st),
              'evb_linear_fwd')
        check(L.evb_groupnorm_relu_fwd(ptr(h1), ptr(g1.weight), ptr(g1.bias), ptr(a1), ptr(st1), c_int(n), c_int(co), c_int(G),
                                       c_float(g1.eps), c_int(1), st), 'evb_groupnorm_relu_fwd')
        check(L.evb_linear_fwd(ptr(a1), ptr(l2.weight), ptr(l2.bias), ptr(h2), c_int(n), c_int(co), c_int(co), c_int(0), st),
              'evb_linear_fwd')
        check(L.evb_groupnorm_relu_fwd(ptr(h2), ptr(g2.weight), ptr(g2.bias), ptr(sf), ptr(st2), c_int(n), c_int(co), c_int(G),
                                       c_float(g2.eps), c_int(0), st), 'evb_groupnorm_relu_fwd')
        dsf = self._new(n, co, dtype=torch.float32) if train else None
        sf_name = None
        if self.tf is not None:
            sf_name = {id(mod): path for path, mod in self.m.named_modules()}.get(id(last))
            self.tf('fwd', sf_name, sf)
        if train:
            def bwd():
                acc = c_int(1 if self.accumulate else 0)
                st_ = stream()
                for d in extra:
                    check(L.evb_copy2d_f32(ptr(d), c_int(co), ptr(dsf), c_int(co), c_int(n), c_int(co), c_int(1), st_),
                          'evb_copy2d_f32')
                if self.tf is not None:
                    self.tf('bwd', sf_name, dsf)
                dh2, da1, dh1 = (self._new(n, co, dtype=torch.float32) for _ in range(3))
                check(L.evb_groupnorm_relu_bwd(ptr(dsf), ptr(h2), ptr(sf), ptr(g2.weight), ptr(st2), ptr(dh2),
                                               ptr(self._g(g2.weight)), ptr(self._g(g2.bias)), c_int(n), c_int(co), c_int(G), acc,
                                               st_), 'evb_groupnorm_relu_bwd')
                check(L.evb_linear_bwd(ptr(dh2), ptr(h2), ptr(a1), ptr(l2.weight), ptr(self._g(l2.weight)), ptr(self._g(l2.bias)),
                                       ptr(da1), c_int(n), c_int(co), c_int(co), c_int(0), acc, c_int(0), st_), 'evb_linear_bwd')
                check(L.evb_groupnorm_relu_bwd(ptr(da1), ptr(h1), ptr(a1), ptr(g1.weight), ptr(st1), ptr(dh1),
                                               ptr(self._g(g1.weight)), ptr(self._g(g1.bias)), c_int(n), c_int(co), c_int(G), acc,
                                               st_), 'evb_groupnorm_relu_bwd')
                check(L.evb_linear_bwd(ptr(dh1), ptr(h1), ptr(scene), ptr(l1.weight), ptr(self._g(l1.weight)),
                                       ptr(self._g(l1.bias)), ptr(dscene), c_int(n), c_int(c5), c_int(co), c_int(0), acc, c_int(0),
                                       st_), 'evb_linear_bwd')
            self.tape.append(bwd)
        return sf, dsf

    def dropout2d(self, x, mask, name=None):
        """nn.Dropout2d in training mode: y = bf16(x * mask[n, c]); the backward is the same map on the gradient"""
        L = self.L
        n, h, w, c = x.data.shape
        y = Act(self._new(n, h, w, c))
        check(L.evb_channel_scale(ptr(x.data), ptr(mask), ptr(y.data), c_int(n), c_ll(h * w), c_int(c), stream()),
              'evb_channel_scale')
        self._tf_fwd(y, name)

        def bwd():
            if y.grad is None:
                return
            self._tf_bwd(y)
            g, acc = self._grad_into(x)
            assert not acc
            check(L.evb_channel_scale(ptr(y.grad), ptr(mask), ptr(g), c_int(n), c_ll(h * w), c_int(c), stream()),
                  'evb_channel_scale')
        self.tape.append(bwd)
        return y

    def _level(self, i, inner_i, sf_pair, train, scene=None, dscenes=None):
        """one pyramid level: [its scene MLP ->] p_i = fpn_layer(inner_i) -> FS-Relation -> decoder chain (runs on the
        caller's stream).  With scale_aware_proj the level owns its scene MLP (sf_pair None): the four tiny-linear chains run
        in the parallel level branches instead of back to back on the main stream, forward and backward."""
        L = self.L
        if sf_pair is None:
            n_ = scene.shape[0]
            dsc = self._new(n_, scene.shape[1], dtype=torch.float32) if train else None
            if train:
                dscenes.append(dsc)
            if self.fs_v2:
                sf_pair = self._scene_mlp_v2(scene, n_, self.scene[i], train, dsc)
            else:
                l1, l2 = self.scene[i]
                sf_pair = self._scene_mlp_one(scene, n_, l1, l2, train, dsc)
        p = self.conv(inner_i, self.fpn_layer[i], train=train)
        self._dbg('p%d' % (i + 2), p)
        (cc, cb), (rc, rb) = self.content[i], self.reenc[i]
        # conv + bias -> training-mode BN: statistics from the conv epilogue, bias gradient identically zero
        u1 = self.conv(p, cc, bias=True, train=train, stats=train, bias_grad_zero=train and cb.bn.training)
        u2 = self.conv(p, rc, bias=True, train=train, stats=train, bias_grad_zero=train and rb.bn.training)
        f1 = self._bn_fold(u1, cb, train)
        f2 = self._bn_fold(u2, rb, train)
        nn_, hh, ww, c = u1.data.shape
        m_rows = nn_ * hh * ww
        sf, dsf = sf_pair
        if self.fs_v2:
            return self._level_v2(i, p, u1, u2, cb, rb, f1, f2, sf, dsf, train)
        z = Act(self._new(nn_, hh, ww, c))
        rel = self._new(m_rows, dtype=torch.float32)
        check(L.evb_relation_fwd(ptr(u1.data), ptr(u2.data), ptr(f1[2]), ptr(f1[3]), ptr(f2[2]), ptr(f2[3]), ptr(sf),
                                 ptr(z.data), ptr(rel), c_ll(m_rows), c_int(hh * ww), c_int(c), stream()), 'evb_relation_fwd')
        d0 = self.dec_blocks[i][0][0]
        self._tf_fwd(z, d0.name + ':in' if d0.name else None)   # r * reenc(p) is the input of the level's first decoder conv
        if train:
            def bwd():
                if z.grad is None:
                    return
                self._tf_bwd(z)
                g1 = self._new(*u1.data.shape)
                g2 = self._new(*u2.data.shape)
                ws = self._ws(L.evb_relation_bwd_workspace(c_ll(m_rows), c_int(hh * ww), c_int(c)))
                check(L.evb_relation_bwd(ptr(z.grad), ptr(u1.data), ptr(u2.data), ptr(f1[2]), ptr(f1[3]), ptr(f2[2]),
                                         ptr(f2[3]), ptr(sf), ptr(rel), ptr(g1), ptr(g2), ptr(dsf), c_ll(m_rows),
                                         c_int(hh * ww), c_int(c), ptr(ws), stream()), 'evb_relation_bwd')
                self._bn_backward(g1, u1, cb, f1, 0, None, None)
                self._bn_backward(g2, u2, rb, f2, 0, None, None)
            # runs BEFORE the conv backward closures of u1/u2 (registered earlier => run later)
            self.tape.append(bwd)
        self._dbg('z%d' % i, z)
        self._dbg('rel%d' % i, rel)
        self._dbg('sf%d' % i, sf)
        return self._decoder_chain(i, z, train)

    def _level_v2(self, i, p, u1, u2, cb, rb, f1, f2, sf, dsf, train):
        """FSRelationV2 tail of a level (fs_relation.py:142-163): cat([r * reenc(p), p]) -> project (1x1 conv 2C -> C, BN,
        ReLU, Dropout2d) -> decoder chain.  The concatenation is one [N, H, W, 2C] buffer: the relation kernel writes its
        first half, a strided copy the second; backward splits the buffer's gradient the same way."""
        L = self.L
        nn_, hh, ww, c = u1.data.shape
        m_rows = nn_ * hh * ww
        pc, pb, drop_mod = self.project[i]
        named = pc.name is not None and not self.scene_shared    # a shared `project` is called once per level: no unique path
        cat = Act(self._new(nn_, hh, ww, 2 * c))
        rel = self._new(m_rows, dtype=torch.float32)
        second = ctypes.c_void_p(cat.data.data_ptr() + 2 * c)
        check(L.evb_relation_fwd_v2(ptr(u1.data), ptr(u2.data), ptr(f1[2]), ptr(f1[3]), ptr(f2[2]), ptr(f2[3]), ptr(sf),
                                    ptr(cat.data), c_int(2 * c), ptr(rel), c_ll(m_rows), c_int(hh * ww), c_int(c), stream()),
              'evb_relation_fwd_v2')
        check(L.evb_copy2d_bf16(ptr(p.data), c_int(c), second, c_int(2 * c), c_ll(m_rows), c_int(c), c_int(0), stream()),
              'evb_copy2d_bf16')
        self._tf_fwd(cat, pc.name + ':in' if named else None)
        if train:
            def bwd():
                if cat.grad is None:
                    return
                self._tf_bwd(cat)
                g1 = self._new(*u1.data.shape)
                g2 = self._new(*u2.data.shape)
                ws = self._ws(L.evb_relation_bwd_workspace(c_ll(m_rows), c_int(hh * ww), c_int(c)))
                check(L.evb_relation_bwd_v2(ptr(cat.grad), c_int(2 * c), ptr(u1.data), ptr(u2.data), ptr(f1[2]), ptr(f1[3]),
                                            ptr(f2[2]), ptr(f2[3]), ptr(sf), ptr(rel), ptr(g1), ptr(g2), ptr(dsf), c_ll(m_rows),
                                            c_int(hh * ww), c_int(c), ptr(ws), stream()), 'evb_relation_bwd_v2')
                gp, acc = self._grad_into(p)    # d(cat)[..., C:] flows straight into d(p)
                check(L.evb_copy2d_bf16(ctypes.c_void_p(cat.grad.data_ptr() + 2 * c), c_int(2 * c), ptr(gp), c_int(c),
                                        c_ll(m_rows), c_int(c), c_int(1 if acc else 0), stream()), 'evb_copy2d_bf16')
                self._bn_backward(g1, u1, cb, f1, 0, None, None)
                self._bn_backward(g2, u2, rb, f2, 0, None, None)
            self.tape.append(bwd)
        q = self.conv(cat, pc, train=train, stats=train)
        pfx = pc.name.rsplit('.', 1)[0] if named else None
        y = self.bn_act(q, pb, True, train=train, name=pfx + '.2' if pfx else None)
        if train and self.drop_p > 0 and drop_mod.training:
            y = self.dropout2d(y, self._drop_masks[i], name=pfx + '.3' if pfx else None)
        self._dbg('z%d' % i, y)
        return self._decoder_chain(i, y, train)

    def _decoder_chain(self, i, z, train):
        y = z
        for (cp, bp) in self.dec_blocks[i]:
            o = self.conv(y, cp, train=train, stats=train)
            # '<layer>.3' = the layer's last module (UpsamplingBilinear2d, or Identity after the ReLU when there is no upsample)
            nm = cp.name.rsplit('.', 1)[0] + '.3' if cp.name else None
            y = (self.bn_relu_up(o, bp, 2, train=train, name=nm) if self.dec_nup[i]
                 else self.bn_act(o, bp, True, train=train, name=nm))
        self._dbg('dec%d' % i, y)
        return y

    def _head(self, feats, train):
        L = self.L
        # ---- FPN (top-down, nearest x2 fused into the lateral 1x1 epilogue)
        inner = [None] * 4
        # the fused lateral + top-down sum is the reference's `last_inner`, i.e. the input of fpn_layer{i} ('<conv>:in')
        def _in(cp_):
            return cp_.name + ':in' if cp_.name else None
        inner[3] = self.conv(feats[3], self.fpn_inner[3], train=train, name=_in(self.fpn_layer[3]))
        for i in (2, 1, 0):
            inner[i] = self.conv(feats[i], self.fpn_inner[i], add=inner[i + 1], add_mode=2, train=train,
                                 name=_in(self.fpn_layer[i]))
        # ---- scene embedding
        c5 = feats[3]
        n, h5, w5, cc5 = c5.data.shape
        scene = self._new(n, cc5, dtype=torch.float32)
        check(L.evb_gap_fwd(ptr(c5.data), ptr(scene), c_int(n), c_int(h5 * w5), c_int(cc5), stream()), 'evb_gap_fwd')
        scene_name = None
        if self.tf is not None:
            l1_ = self.scene[0][0]
            scene_name = {id(mod): path for path, mod in self.m.named_modules()}.get(id(l1_)) + ':in'
            self.tf('fwd', scene_name, scene)
        if train:
            # runs last among head closures (registered first): needs dscene complete
            holder = {}

            def gap_bwd():
                g, acc = self._grad_into(c5)
                if not acc:
                    self._zero(g)
                ds = holder['dscenes']
                for d in ds[1:]:     # d(scene) of the other scene MLPs, summed in level order (deterministic)
                    check(L.evb_copy2d_f32(ptr(d), c_int(cc5), ptr(ds[0]), c_int(cc5), c_int(n), c_int(cc5), c_int(1),
                                           stream()), 'evb_copy2d_f32')
                if self.tf is not None:
                    self.tf('bwd', scene_name, ds[0])
                check(L.evb_gap_bwd(ptr(ds[0]), ptr(g), c_int(n), c_int(h5 * w5), c_int(cc5), stream()), 'evb_gap_bwd')
            self.tape.append(gap_bwd)
        dscenes = []
        if train:
            holder['dscenes'] = dscenes
        if self.scene_shared:
            # scale_aware_proj=False (fs_relation.py:29-35,63-66): one MLP on the main stream, every level reads the same
            # scene vector; each level's relation backward writes its own d(sf) buffer, summed before the MLP's backward
            l1 = self.scene[0][0]
            extra = [self._new(n, l1.out_channels, dtype=torch.float32) if train else None for _ in range(3)]
            dsc = self._new(n, cc5, dtype=torch.float32) if train else None
            if train:
                dscenes.append(dsc)
            if self.fs_v2:
                sf, dsf = self._scene_mlp_v2(scene, n, self.scene[0], train, dsc, extra=extra if train else ())
            else:
                sf, dsf = self._scene_mlp_one(scene, n, l1, self.scene[0][1], train, dsc, extra=extra if train else ())
            sfs = [(sf, dsf)] + [(sf, e) for e in extra]
        else:
            sfs = [None] * 4
        if self.fs_v2 and train and self.drop_p > 0:
            # Dropout2d channel masks of the four `project` blocks, drawn in level order from torch's generator exactly as
            # feature_dropout does (bf16 noise of shape [N, C, 1, 1]: bernoulli_(1 - p) then div_(1 - p)), so a run seeded
            # like the reference drops the same channels
            keep = 1.0 - self.drop_p
            cdrop = self.project[0][0].co
            self._drop_masks = [torch.empty(n, cdrop, 1, 1, dtype=BF16, device=self.dev).bernoulli_(keep).div_(keep)
                                .float().view(n, cdrop).contiguous() for _ in range(4)]
        # ---- per pyramid level: FPN output conv -> FS-Relation -> decoder chain.  The four chains are independent:
        #      levels 1..3 (small maps, latency-bound kernels) run on their own streams = parallel graph branches
        outs = [None] * 4
        main = torch.cuda.current_stream()
        if train:
            self._fork_idx = len(self.tape)
        ev_fork = None
        if self.level_streams is not None:
            ev_fork = torch.cuda.Event()
            ev_fork.record(main)
        for i in range(4):
            st_i = self.level_streams[i - 1] if (self.level_streams is not None and i > 0) else None
            t0 = len(self.tape)
            if st_i is not None:
                st_i.wait_event(ev_fork)
                with torch.cuda.stream(st_i):
                    outs[i] = self._level(i, inner[i], sfs[i], train, scene, dscenes)
                if train:
                    self._tape_tags.append((t0, len(self.tape), st_i))
            else:
                outs[i] = self._level(i, inner[i], sfs[i], train, scene, dscenes)
        if self.level_streams is not None:
            for st_i in self.level_streams:
                main.wait_stream(st_i)
        if train:
            self._join_idx = len(self.tape)
        merged = Act(self._new(*outs[0].data.shape))
        check(L.evb_merge4(ptr(outs[0].data), ptr(outs[1].data), ptr(outs[2].data), ptr(outs[3].data), ptr(merged.data),
                           c_ll(merged.data.numel()), stream()), 'evb_merge4')
        d00 = self.dec_blocks[0][0][0]
        dname = d00.name.split('.blocks.')[0] + '.dropout' if d00.name else None
        drop_p = float(getattr(self.m.head.fpn_decoder, 'dropout_rate', -1))
        drop_on = train and drop_p > 0 and self.m.head.fpn_decoder.dropout.training
        self._tf_fwd(merged, None if drop_on else dname)
        if train:
            def bwd():
                if merged.grad is None:
                    return
                self._tf_bwd(merged)
                dq = self._new(*merged.data.shape)
                check(L.evb_scale_add(ptr(merged.grad), c_float(0.25), None, ptr(dq), c_ll(dq.numel()), stream()),
                      'evb_scale_add')
                for o in outs:
                    o.grad, o.has_grad = dq, True
            self.tape.append(bwd)
        self._dbg('merged', merged)
        if drop_on:
            # classifier dropout (fpn.py:175-176,190): the survivors are drawn by torch's own dropout kernel on a tensor of ones
            # of the reference's shape / dtype (NCHW bf16), i.e. from the same Philox stream as the reference's call; the scale
            # 1 / (1 - p) is applied in fp32 inside the kernel, as torch's fused dropout does
            nb, hh_, ww_, cc_ = merged.data.shape
            keep = torch.nn.functional.dropout(torch.ones(nb, cc_, hh_, ww_, dtype=BF16, device=self.dev), drop_p, True)
            keep = keep.permute(0, 2, 3, 1).contiguous()
            dropped = Act(self._new(nb, hh_, ww_, cc_))
            sc = c_float(1.0 / (1.0 - drop_p))
            check(L.evb_dropout_apply(ptr(merged.data), ptr(keep), sc, ptr(dropped.data), c_ll(merged.data.numel()), stream()),
                  'evb_dropout_apply')
            self._tf_fwd(dropped, dname)

            def bwd_drop():
                if dropped.grad is None:
                    return
                self._tf_bwd(dropped)
                g, acc = self._grad_into(merged)
                assert not acc
                check(L.evb_dropout_apply(ptr(dropped.grad), ptr(keep), sc, ptr(g), c_ll(g.numel()), stream()),
                      'evb_dropout_apply')
            self.tape.append(bwd_drop)
            return dropped
        return merged

    def _classify(self, feat, cp, f, train, name='logits'):
        """1x1 / 3x3 classifier conv (Cout zero-padded to 64) + bilinear x f on the first 16 channels -> logits
        [N, f*h, f*w, 16] (AssymetricDecoder classifier, fpn.py:178-181)."""
        L = self.L
        cls = self.conv(feat, cp, bias=True, train=train)
        n, h4, w4, _ = cls.data.shape
        ld = self.ld if cp is self.cls else 16
        logits = self._new(n, h4 * f, w4 * f, ld)
        if f == 1:   # classifier at the output resolution (no up-sampling): the first ld channels of the padded conv output
            check(L.evb_copy2d_bf16(ptr(cls.data), c_int(64), ptr(logits), c_int(ld), c_ll(n * h4 * w4), c_int(ld), c_int(0),
                                    stream()), 'evb_copy2d_bf16')
        else:
            check(L.evb_bilinear_up(ptr(cls.data), None, None, ptr(logits), c_int(n), c_int(h4), c_int(w4), c_int(ld),
                                    c_int(64), c_int(ld), c_int(f), stream()), 'evb_bilinear_up(logits)')
        if self.tf is not None and cp.name and f != 1:
            self.tf('fwd', cp.name.rsplit('.', 1)[0] + '.1', logits)
        self._dbg('cls' if name == 'logits' else name + '_cls', cls)
        self._dbg(name, logits)
        return cls, logits

    # ------------------------------------------------------------------ loss groups
    def _loss_stats(self, cls, logits, labels, k, f, names, weight=1.0):
        """Pass A of one loss group (CE+Dice for k >= 2, BCE+Dice for k == 1) on logits [N,H,W,16]."""
        L = self.L
        labels = labels.contiguous()
        if labels.dtype != torch.int64:
            labels = labels.long()
        n, hh, ww, _ = logits.shape
        npx = n * hh * ww
        stats = self._new(2 + 3 * k, dtype=torch.float32)
        ws = self._ws(L.evb_loss_workspace(c_ll(npx), c_int(k)))
        check(L.evb_loss_stats(ptr(logits), ptr(labels), c_ll(npx), c_int(k), c_int(logits.shape[-1]), c_int(self.ignore_index),
                               ptr(stats), ptr(ws), stream()), 'evb_loss_stats')
        g = dict(cls=cls, logits=logits, labels=labels, stats=stats, npx=npx, k=k, f=f, names=names, weight=weight,
                 name=cls.name.rsplit('.', 1)[0] + '.1' if (cls.name and f != 1) else None)
        self._groups.append(g)
        return g

    def _network_losses(self, x, labels):
        """encoder + head + loss statistics of every loss group (overridden by derived engines)."""
        feats = self._encoder(x, True)
        merged = self._head(feats, True)
        cls, logits = self._classify(merged, self.cls, self.cls_scale, True)
        first = 'bce_loss' if self.K == 1 else 'ce_loss'
        self._loss_stats(cls, logits, labels, self.K, self.cls_scale, (first, 'dice_loss'))

    # ------------------------------------------------------------------ public steps
    def _forward_part1(self, x, labels):
        """pack weights, encoder, head, loss statistics (everything before the Dice all-reduce)."""
        self.tape = []
        self._tape_tags, self._fork_idx, self._join_idx = [], None, None
        self._tape_splits = []
        self._bn_tracked = []
        self._groups = []
        x = x.contiguous() if x.dtype == torch.uint8 else x.contiguous().float()
        self.pack_weights()
        self._network_losses(x, labels)
        if self._bn_tracked:
            torch._foreach_add_([bn.num_batches_tracked for bn in self._bn_tracked], 1)
        if self.world > 1 and self.sync_dice:
            tot = sum(3 * g['k'] for g in self._groups)
            if getattr(self, '_dice_global', None) is None or self._dice_global.numel() != tot:
                self._dice_global = torch.zeros(tot, dtype=torch.float32, device=self.dev)

    def _dice_allreduce(self):
        """all_reduce_sum of the Dice statistics of every loss group in one message (ever/module/loss.py:20-23,46-48);
        eager NCCL, never graph-captured."""
        if self.world > 1 and self.sync_dice:
            import torch.distributed as dist
            off = 0
            for g in self._groups:
                self._dice_global[off:off + 3 * g['k']].copy_(g['stats'][2:])
                off += 3 * g['k']
            dist.all_reduce(self._dice_global)

    def _forward_part2(self):
        L = self.L
        glob = self.world > 1 and self.sync_dice
        out, off = {}, 0
        for g in self._groups:
            k = g['k']
            dice_stats = self._dice_global[off:off + 3 * k] if glob else g['stats'][2:]
            off += 3 * k
            scale = float(self.world) if glob else 1.0
            losses = self._new(2, dtype=torch.float32)
            coef = self._new(1 + 2 * k, dtype=torch.float32)
            wt = g['weight']
            check(L.evb_loss_finalize(ptr(g['stats']), ptr(dice_stats), c_int(k), c_float(self.smooth),
                                      c_float(self.ce_w * wt), c_float(self.dice_w * wt), c_float(scale), ptr(losses),
                                      ptr(coef), stream()), 'evb_loss_finalize')
            g['coef'] = coef
            w0, w1 = self.ce_w * wt, self.dice_w * wt
            out[g['names'][0]] = losses[0] * w0 if w0 != 1.0 else losses[0]
            out[g['names'][1]] = losses[1] * w1 if w1 != 1.0 else losses[1]
        self._saved_for_backward = self._groups
        return out

    def forward_train(self, x, labels):
        self._forward_part1(x, labels)
        self._dice_allreduce()
        return self._forward_part2()

    def backward(self, allreduce=True, attach=True, upstream=None):
        """Backward of the step whose losses the last training forward returned; gradients land in the flat arena.
        attach: alias every trainable parameter's ``.grad`` to its arena slot (native path).  upstream: {loss name: 0-dim
        device tensor} = d(total)/d(loss) from autograd (Launcher divides by forward_times, GradScaler multiplies,
        ever/core/launcher.py:196, ever/interface/module.py:76-81); applied on the device to the loss-gradient coefficients,
        which the logit gradient is linear in -- no host synchronisation."""
        L = self.L
        if self._saved_for_backward is None:
            raise RuntimeError('backward() without a preceding training forward')
        groups = self._saved_for_backward
        self._saved_for_backward = None
        if isinstance(groups, str):   # graph_forward already replayed the backward
            if allreduce or self._ar_pending:   # buckets issued by the replay are in flight: the arena is read next
                self.allreduce_grads()
            if attach:
                self.attach_grads()
            return
        self._loss_backward(groups, upstream)
        for lo, hi in self._segments():
            self._run_tape(lo, hi)
            if allreduce and self.world > 1:
                self._allreduce_bucket(lo)
        self._join_side()
        self.tape = []
        self._tape_tags, self._fork_idx, self._join_idx = [], None, None
        if allreduce:
            self.allreduce_grads()
        if attach:
            self.attach_grads()

    def _loss_backward(self, groups, upstream=None):
        """d(loss)/d(logits) of every loss group and its bilinear backward into the classifier outputs"""
        L = self.L
        for g in groups:
            if upstream is not None:   # coef = {first-loss scale, Dice A_c[k], B_c[k]} (evb_loss_finalize)
                u0, u1 = upstream.get(g['names'][0]), upstream.get(g['names'][1])
                g['coef'][:1].mul_(u0.to(torch.float32) if u0 is not None else 0.0)
                g['coef'][1:].mul_(u1.to(torch.float32) if u1 is not None else 0.0)
            logits, cls, k, f = g['logits'], g['cls'], g['k'], g['f']
            n, hh, ww, ld = logits.shape
            dlogits = self._new(n, hh, ww, ld)
            check(L.evb_loss_grad(ptr(logits), ptr(g['labels']), c_ll(g['npx']), c_int(k), c_int(ld),
                                  c_int(self.ignore_index), ptr(g['coef']), ptr(dlogits), stream()), 'evb_loss_grad')
            if self.tf is not None and g.get('name'):
                self.tf('bwd', g['name'], dlogits)
            cls.grad = self._zero(torch.empty_like(cls.data))   # padding channels 16..63 stay zero
            cls.has_grad = True
            if f == 1:
                check(L.evb_copy2d_bf16(ptr(dlogits), c_int(ld), ptr(cls.grad), c_int(64), c_ll(n * hh * ww), c_int(ld),
                                        c_int(0), stream()), 'evb_copy2d_bf16')
                continue
            ws = self._ws(L.evb_bilinear_up_bwd_workspace(c_int(n), c_int(hh // f), c_int(ww // f), c_int(ld), c_int(f)))
            check(L.evb_bilinear_up_bwd_sep(ptr(dlogits), ptr(cls.grad), c_int(n), c_int(hh // f), c_int(ww // f), c_int(ld),
                                            c_int(ld), c_int(64), c_int(f), ptr(ws), c_ll(self._ws_cap()), stream()),
                  'evb_bilinear_up_bwd_sep')

    def _segments(self):
        """[(lo, hi)] tape ranges in the order backward runs them (last segment of the forward first)"""
        cuts = sorted(set(c for c in self._tape_splits if 0 < c < len(self.tape)))
        bounds = [0] + cuts + [len(self.tape)]
        return [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 2, -1, -1)]

    def _bucket_ranges(self):
        """arena ranges of the gradient buckets, in backward order: one per encoder split (ar_split_stages) + the rest.
        Parameters are laid out in forward order, so 'everything from the first parameter of stage s on' is a contiguous
        tail of the arena."""
        if getattr(self, '_buckets', None) is None:
            names = [n for n, _ in self.m.named_parameters()]
            offs = []
            for st in sorted(self.ar_split_stages):
                pfx = 'en.resnet.layer%d.' % (st + 1)
                idx = next((i for i, n in enumerate(names) if n.startswith(pfx)), None)
                if idx is not None and idx > 0:
                    offs.append(self._slots[idx][0])
            bounds = [0] + sorted(set(offs)) + [self.flat_g.numel()]
            self._buckets = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 2, -1, -1)]
        return self._buckets

    def _allreduce_bucket(self, seg_lo):
        """asynchronous NCCL all-reduce (mean) of the bucket whose gradients the segment starting at tape index seg_lo has
        just completed; it runs on NCCL's stream behind the current stream's work and overlaps the next segment"""
        import torch.distributed as dist
        cuts = sorted(set(c for c in self._tape_splits if 0 < c < max(len(self.tape), 1)), reverse=True)
        order = cuts + [0]
        buckets = self._bucket_ranges()
        if len(buckets) != len(order):      # split points and buckets disagree (e.g. frozen stages): single bucket at the end
            if seg_lo == 0:
                self._issue_allreduce(0, self.flat_g.numel())
            return
        a, b = buckets[order.index(seg_lo)]
        self._issue_allreduce(a, b)

    def _issue_allreduce(self, a, b):
        """async all-reduce (mean) of arena[a:b], ordered behind BOTH the main stream (BatchNorm / bias / linear gradients)
        and the weight-gradient side stream -- without stalling the main stream: the side stream waits for an event of the
        main stream and the collective is issued from the side stream's context, so NCCL's stream depends on the side
        stream only; the main chain (dgrad -> BN backward) runs on."""
        import torch.distributed as dist
        if self.side is None:
            self._ar_pending.append(dist.all_reduce(self.flat_g[a:b], op=dist.ReduceOp.AVG, async_op=True))
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self._ar_pending.append(dist.all_reduce(self.flat_g[a:b], op=dist.ReduceOp.AVG, async_op=True))
        self._side_used = True

    def _run_tape(self, lo=0, hi=None):
        """Run the backward closures in reverse order.  Closures of a pyramid-level branch run on that branch's stream:
        entering the branch region the branch streams wait for the main stream (dq is ready), leaving it the main stream
        waits for all of them (dsf_i and inner_i.grad are complete) -- the mirror image of the forward fork/join."""
        tags = {}
        for a, b, st_ in self._tape_tags:
            for j in range(a, b):
                tags[j] = st_
        main = torch.cuda.current_stream()
        fork_idx, join_idx = self._fork_idx, self._join_idx
        use = self.level_streams is not None and fork_idx is not None
        hi = len(self.tape) if hi is None else hi
        for idx in range(hi - 1, lo - 1, -1):
            if use and idx == join_idx - 1:      # about to enter the branch region (from the loss side)
                ev = torch.cuda.Event()
                ev.record(main)
                for st_ in self.level_streams:
                    st_.wait_event(ev)
            st_ = tags.get(idx)
            if st_ is not None:
                with torch.cuda.stream(st_):
                    self.tape[idx]()
            else:
                self.tape[idx]()
            if use and idx == fork_idx:          # all branch closures are enqueued: join
                for st_ in self.level_streams:
                    main.wait_stream(st_)

    def allreduce_grads(self):
        """The one gradient exchange of the step: NCCL all-reduce (mean) of the flat fp32 gradient arena
        (reference: DDP bucketed all-reduce, ever/trainer/th_ddp_trainer.py:25-30)."""
        if self.world > 1:
            import torch.distributed as dist
            if self._reduced_in_graph:   # the replayed graph contains the bucket all-reduces and their joins
                self._reduced_in_graph = False
                return
            if self._ar_pending:     # the buckets were issued during backward: make the current stream wait for them
                for w in self._ar_pending:
                    w.wait()
                self._ar_pending = []
                return
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.AVG)

    def confusion_matrix(self, mask, labels, cm):
        """cm[K,K] (int64, device) += confusion counts of a uint8 prediction mask against int64 labels; labels outside
        [0,K) (ignore_index) are skipped.  GPU version of ConfusionMatrix.forward (ever/metric/confusion_matrix.py:11-25)."""
        labels = labels.contiguous().long()
        check(self.L.evb_confusion_matrix(ptr(mask), ptr(labels), c_ll(labels.numel()), c_int(cm.shape[0]), ptr(cm),
                                          stream()), 'evb_confusion_matrix')
        return cm

    # ------------------------------------------------------------------ fused optimizer
    def ensure_optimizer_state(self):
        """momentum arena (same slots as flat_w / flat_g), device-side lr / norm scalars, norm workspace"""
        if not hasattr(self, '_mom'):
            n = self.flat_w.numel()
            self._mom = torch.zeros_like(self.flat_w)
            self._lr = torch.zeros(1, dtype=torch.float32, device=self.dev)
            self._norm = torch.zeros(2, dtype=torch.float32, device=self.dev)
            self._sgd_ws = torch.empty(self.L.evb_sgd_workspace(c_ll(n)) // 4 + 4, dtype=torch.float32, device=self.dev)
            self._sgd_first = True

    def sgd_step(self, lr, momentum=0.9, weight_decay=1e-4, max_norm=35.0):
        """clip_grad_norm_(max_norm, 2) + torch.optim.SGD step + zero_grad over the flat arenas
        (ever/interface/module.py:83-108, ever/opt/optimizer.py:7-9) as two kernels; lr lives on the device."""
        L = self.L
        n = self.flat_w.numel()
        self.ensure_optimizer_state()
        self._lr.fill_(float(lr))
        check(L.evb_grad_norm(ptr(self.flat_g), c_ll(n), c_float(max_norm if max_norm else 0.0), ptr(self._norm),
                              ptr(self._sgd_ws), stream()), 'evb_grad_norm')
        mask = self.trainable_mask()
        if mask is None:
            check(L.evb_sgd_step(ptr(self.flat_w), ptr(self.flat_g), ptr(self._mom), c_ll(n), ptr(self._lr),
                                 c_float(momentum), c_float(weight_decay), ptr(self._norm),
                                 c_int(1 if self._sgd_first else 0), c_int(1), stream()), 'evb_sgd_step')
        else:   # frozen parameters: no weight decay, no momentum, no update (torch.optim.SGD skips grad-less parameters)
            check(L.evb_sgd_step_masked(ptr(self.flat_w), ptr(self.flat_g), ptr(self._mom), c_ll(n), ptr(self._lr),
                                        c_float(momentum), c_float(weight_decay), ptr(self._norm),
                                        c_int(1 if self._sgd_first else 0), c_int(1), ptr(mask), stream()),
                  'evb_sgd_step_masked')
        self._sgd_first = False
        return self._norm

    # ------------------------------------------------------------------ CUDA-graph step
    def capture_step(self, x, labels):
        """Capture forward + loss + backward for fixed-shape device inputs into CUDA graph(s).
        Returns (replay_fn, losses dict).  NCCL is never captured: with world > 1 the step is a chain of graphs with the
        eager Dice-statistics all-reduce and the asynchronous gradient-bucket all-reduces between them (see below);
        allreduce_grads() then only waits for the buckets, and the optimizer runs after."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        # the warm-up runs real training forwards: keep the BatchNorm buffers (running statistics, num_batches_tracked) as
        # they were, so that capturing a graph never counts as extra training steps
        bufs = [b for b in self.m.buffers()]
        saved = [b.detach().clone() for b in bufs]
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: sets kernel attributes, sizes the workspace (and NCCL's buffers for the buckets)
                self.forward_train(x, labels)
                self.backward(allreduce=self.world > 1, attach=False)
            for b, s_ in zip(bufs, saved):
                b.copy_(s_)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        split = self.world > 1
        pool = torch.cuda.graph_pool_handle()
        cap_stream = torch.cuda.Stream(device=self.dev, priority=self._prio) if self._prio else None
        g1 = torch.cuda.CUDAGraph()
        if not split:
            with torch.cuda.graph(g1, pool=pool, stream=cap_stream):
                out = self.forward_train(x, labels)
                self.backward(allreduce=False)
            return g1.replay, out
        if os.environ.get('EVB_CAPTURE_NCCL', '1') == '1':
            # world > 1, ONE graph: the Dice-statistics all-reduce and the gradient-bucket all-reduces are captured NCCL
            # nodes.  A bucket's node depends on the weight-gradient branch (and through it on the main chain up to the
            # bucket boundary) and is joined only at the end, so it overlaps the rest of backward without cutting the graph.
            # thread_local: NCCL's watchdog thread polls CUDA events while this thread captures
            with torch.cuda.graph(g1, pool=pool, stream=cap_stream, capture_error_mode='thread_local'):
                out = self.forward_train(x, labels)
                self.backward(allreduce=True, attach=False)

            def replay_one():
                g1.replay()
                self._reduced_in_graph = True
            return replay_one, out
        # world > 1: [forward + loss statistics] | eager Dice all-reduce | [loss + backward of bucket 0] | async all-reduce
        # of bucket 0 | [backward of bucket 1] | async all-reduce of bucket 1 | ...  Every bracket is one CUDA graph; the
        # bucket all-reduces run on NCCL's stream and overlap the graphs that follow (allreduce_grads() waits for them).
        with torch.cuda.graph(g1, pool=pool, stream=cap_stream):
            self._forward_part1(x, labels)
        self._dice_allreduce()
        torch.cuda.synchronize()
        seg_graphs = []
        segs = None
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=pool, stream=cap_stream):
            out = self._forward_part2()
            groups, self._saved_for_backward = self._saved_for_backward, None
            self._loss_backward(groups)
            segs = self._segments()
            self._run_tape(*segs[0])
            self._join_side()
        seg_graphs.append(g2)
        for lo, hi in segs[1:]:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=cap_stream):
                self._run_tape(lo, hi)
                self._join_side()
            seg_graphs.append(g)
        self.tape = []
        self._tape_tags, self._fork_idx, self._join_idx = [], None, None
        buckets = self._bucket_ranges()
        bucketed = len(buckets) == len(seg_graphs) and os.environ.get('EVB_AR_OVERLAP', '1') == '1'

        def replay():
            import torch.distributed as dist
            g1.replay()
            self._dice_allreduce()
            for i, g in enumerate(seg_graphs):
                g.replay()
                if bucketed:
                    a, b = buckets[i]
                    self._ar_pending.append(dist.all_reduce(self.flat_g[a:b], op=dist.ReduceOp.AVG, async_op=True))
        return replay, out

    def graph_forward(self, x, labels):
        """Training step through cached CUDA graphs, for callers with static shapes (the plugin's forward when
        config.cuda_graph is on): inputs are copied into static device buffers, the captured forward + loss + backward
        is replayed, the static loss tensors are returned and the gradients are already in the arena (backward() then
        only runs the gradient all-reduce).  One capture per (shape, dtype) signature."""
        lab = labels if isinstance(labels, dict) else dict(cls=labels)
        key = (tuple(x.shape), x.dtype, bool(self.accumulate)) + tuple((k, tuple(v.shape), v.dtype)
                                                                       for k, v in sorted(lab.items()))
        ent = self._graphs.get(key)
        if ent is None:
            sx = torch.empty_like(x, device=self.dev)
            sl = {k: torch.empty_like(v, device=self.dev) for k, v in lab.items()}
            sx.copy_(x)
            for k, v in lab.items():
                sl[k].copy_(v)
            replay, out = self.capture_step(sx, sl if isinstance(labels, dict) else sl['cls'])
            ent = (sx, sl, replay, out)
            self._graphs[key] = ent
        sx, sl, replay, out = ent
        sx.copy_(x, non_blocking=True)
        for k, v in lab.items():
            sl[k].copy_(v, non_blocking=True)
        replay()
        self._saved_for_backward = 'graph'
        return out

    @torch.no_grad()
    def forward_eval(self, x, return_mask=False):
        L = self.L
        self.tape = []
        x = x.contiguous() if x.dtype == torch.uint8 else x.contiguous().float()
        self.pack_weights()
        feats = self._encoder(x, False)
        merged = self._head(feats, False)
        cls, logits = self._classify(merged, self.cls, self.cls_scale, False)
        n, hh, ww, _ = logits.shape
        prob = self._new(n, self.K, hh, ww, dtype=torch.float32)
        mask = self._new(n, hh, ww, dtype=torch.uint8)
        check(L.evb_softmax_nchw(ptr(logits), ptr(prob), ptr(mask), c_ll(n * hh * ww), c_int(hh * ww), c_int(self.K),
                                 c_int(logits.shape[-1]), stream()), 'evb_softmax_nchw')
        self.last_logits = logits
        if return_mask:
            return prob, mask
        return prob


class ChangeStarEngine(FarSegEngine):
    """FarSeg features of both temporal images (one batch of 2N) + ChangeMixin on (f1,f2) and (f2,f1).

    The 16-channel ChangeMixin convolutions run on the tensor-core kernel zero-padded to 64 channels; the first conv
    over cat(fa, fb) is computed as conv(fa, W[:, :C]) + conv(fb, W[:, C:]) (second conv accumulates in its epilogue), so
    the concatenation is never materialised.  Both directions share the parameters: their gradients accumulate."""

    def _collect_extra(self):
        cm = self.m.changemixin.convs
        w0 = cm[0][0].weight
        inner, cdec = w0.shape[0], w0.shape[1] // 2
        ld = 2 * cdec * 9
        self.cm_a = ConvP(w0, None, 1, co=inner, ci=cdec, k=3, cout_pad=64, w_off=0, w_ld=ld)
        self.cm_b = ConvP(w0, None, 1, co=inner, ci=cdec, k=3, cout_pad=64, w_off=cdec * 9, w_ld=ld)
        self.cm_bn0 = self._mk_bn(cm[0][1], cpad=64)
        self.cm_mid = [(ConvP(cm[i][0].weight, None, 1, cout_pad=64, cin_pad=64), self._mk_bn(cm[i][1], cpad=64))
                       for i in (1, 2, 3)]
        self.cm_cls = ConvP(cm[4].weight, cm[4].bias, 1, cout_pad=64, cin_pad=64)
        self.convs += [self.cm_a, self.cm_b] + [c for c, _ in self.cm_mid] + [self.cm_cls]

    def _split(self, merged, n, train):
        """views of the t1 / t2 halves of the feature batch; their gradients land in the halves of one buffer"""
        f1, f2 = Act(merged.data[:n]), Act(merged.data[n:])
        if train:
            g = self._zero(torch.empty_like(merged.data))
            f1.grad, f2.grad = g[:n], g[n:]

            def bwd():
                merged.grad, merged.has_grad = g, True
            self.tape.append(bwd)
        return f1, f2

    def _changemixin(self, fa, fb, train, name):
        u = self.conv(fa, self.cm_a, train=train)
        u = self.conv(fb, self.cm_b, add=u, add_mode=1, train=train)
        y = self.bn_act(u, self.cm_bn0, True, train=train)
        for cp, bp in self.cm_mid:
            y = self.bn_act(self.conv(y, cp, train=train, stats=train), bp, True, train=train)
        return self._classify(y, self.cm_cls, 4, train, name=name)

    @staticmethod
    def _reorder(x):
        if x.dtype == torch.uint8 or x.dim() != 4 or not x.is_floating_point():
            raise ValueError('ChangeStarB200 takes float NCHW pairs [N, 2*Cin, H, W] (t1 channels then t2 channels); raw '
                             'uint8 NHWC tiles are only supported by FarSegB200 -- normalise and stack the pair first')
        n, c2, h, w = x.shape
        if c2 % 2:
            raise ValueError('ChangeStarB200 input needs an even channel count (t1 | t2), got %d' % c2)
        return x.view(n, 2, c2 // 2, h, w).transpose(0, 1).reshape(2 * n, c2 // 2, h, w).contiguous().float()

    def _network_losses(self, x, labels):
        n = x.shape[0]
        feats = self._encoder(self._reorder(x), True)
        merged = self._head(feats, True)
        f1, f2 = self._split(merged, n, True)
        cls, logits = self._classify(f1, self.cls, self.cls_scale, True)
        first = 'bce_loss' if self.K == 1 else 'ce_loss'
        self._loss_stats(cls, logits, labels['cls'], self.K, self.cls_scale, (first, 'dice_loss'))
        for name, (fa, fb) in (('c12', (f1, f2)), ('c21', (f2, f1))):
            c, lg = self._changemixin(fa, fb, True, name)
            self._loss_stats(c, lg, labels['change'], 1, 4, (name + '_bce_loss', name + '_dice_loss'))

    @torch.no_grad()
    def forward_eval(self, x, return_mask=False):
        L = self.L
        self.tape = []
        n = x.shape[0]
        self.pack_weights()
        feats = self._encoder(self._reorder(x), False)
        merged = self._head(feats, False)
        f1, f2 = self._split(merged, n, False)
        cls, logits = self._classify(f1, self.cls, self.cls_scale, False)
        _, clog = self._changemixin(f1, f2, False, 'c12')
        out = {}
        for key, lg, k in (('seg', logits, self.K), ('change', clog, 1)):
            nn_, hh, ww, _ = lg.shape
            prob = self._new(nn_, k, hh, ww, dtype=torch.float32)
            mask = self._new(nn_, hh, ww, dtype=torch.uint8)
            check(L.evb_softmax_nchw(ptr(lg), ptr(prob), ptr(mask), c_ll(nn_ * hh * ww), c_int(hh * ww), c_int(k),
                                     c_int(lg.shape[-1]), stream()), 'evb_softmax_nchw')
            out[key] = prob
            out[key + '_mask'] = mask
        return out
