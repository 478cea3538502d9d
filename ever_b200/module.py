"""The drop-in models: ``FarSegB200`` registered in ``ever.registry.MODEL`` as 'FarSegB200' (and 'FarSeg').

The module owns real ``nn.Parameter``s laid out under exactly the reference's ``state_dict`` keys
(``en.resnet.layer1.0.conv1.weight`` ... ``head.fpn_decoder.classifier.0.bias``), so reference
checkpoints load unchanged and torch optimizers / DDP / ``count_model_parameters`` see ordinary parameters.
The ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects are parameter containers only: their ``forward`` is never
called.  All arithmetic runs in libevb200.so through ``ever_b200.engine.FarSegEngine``.
"""
import math

import torch
import torch.nn as nn

from ._ever_api import MODEL, ERModule
from .engine import ChangeStarEngine, FarSegEngine

RESNETS = {  # block kind, blocks per stage  (ever/module/_resnets.py:241-278)
    'resnet18': ('basic', (2, 2, 2, 2)),
    'resnet34': ('basic', (3, 4, 6, 3)),
    'resnet50': ('bottleneck', (3, 4, 6, 3)),
    'resnet101': ('bottleneck', (3, 4, 23, 3)),
    # deep-stem variants: three 3x3 convs instead of the 7x7 (_resnets.py:137-147, 327-345)
    'resnet50_v1c': ('bottleneck', (3, 4, 6, 3)),
    'resnet101_v1c': ('bottleneck', (3, 4, 23, 3)),
    # ResNeXt: grouped 3x3 in the bottleneck (_resnets.py:80-84, 291-324)
    'resnext50_32x4d': ('bottleneck', (3, 4, 6, 3)),
    'resnext101_32x4d': ('bottleneck', (3, 4, 23, 3)),
    'resnext101_32x8d': ('bottleneck', (3, 4, 23, 3)),
}
RESNEXT = {'resnext50_32x4d': (32, 4), 'resnext101_32x4d': (32, 4), 'resnext101_32x8d': (32, 8)}   # groups, width_per_group


class _Block(nn.Module):
    """Parameter container with the attribute names of BasicBlock / Bottleneck (_resnets.py:32-112)."""

    def __init__(self, kind, cin, planes, stride, down, groups=1, base_width=64):
        super().__init__()
        self.kind, self.stride = kind, stride
        width = int(planes * (base_width / 64.)) * groups   # _resnets.py:80
        if kind == 'basic':
            self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
        else:
            self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(width)
            self.conv2 = nn.Conv2d(width, width, 3, stride, 1, groups=groups, bias=False)
            self.bn2 = nn.BatchNorm2d(width)
            self.conv3 = nn.Conv2d(width, planes * 4, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = down


class _ResNetParams(nn.Module):
    def __init__(self, resnet_type, in_channels=3):
        super().__init__()
        kind, counts = RESNETS[resnet_type]
        groups, base_width = RESNEXT.get(resnet_type, (1, 64))
        exp = 1 if kind == 'basic' else 4
        self.kind = kind
        self.deep_stem = resnet_type.endswith('_v1c')
        if self.deep_stem:   # keys stem.0 / .1 / .3 / .4 / .6 / .7 (the ReLUs hold no parameters)
            self.stem = nn.Sequential(nn.Conv2d(in_channels, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(True),
                                      nn.Conv2d(32, 32, 3, 1, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(True),
                                      nn.Conv2d(32, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(True))
        else:
            self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
            self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), counts), 1):
            blocks = []
            for b in range(n):
                s = 2 if (b == 0 and li > 1) else 1
                down = None
                if b == 0 and (s != 1 or cin != planes * exp):
                    down = nn.Sequential(nn.Conv2d(cin, planes * exp, 1, s, bias=False), nn.BatchNorm2d(planes * exp))
                blocks.append(_Block(kind, cin, planes, s, down, groups, base_width))
                cin = planes * exp
            setattr(self, 'layer%d' % li, nn.Sequential(*blocks))
        self.out_channels = tuple(c * exp for c in (64, 128, 256, 512))
        for m in self.modules():  # _resnets.py:163-169
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')


class _Encoder(nn.Module):
    def __init__(self, resnet_type, in_channels=3):
        super().__init__()
        self.resnet = _ResNetParams(resnet_type, in_channels)


def _fpn_conv(cin, cout, k):  # ConvBlock(bn=False, relu=False), kaiming_uniform(a=1): fpn.py:18-35, ops.py:45-60
    seq = nn.Sequential(nn.Conv2d(cin, cout, k, 1, (k - 1) // 2, bias=False), nn.Identity(), nn.Identity())
    nn.init.kaiming_uniform_(seq[0].weight, a=1)
    return seq


class _FPN(nn.Module):
    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        for i, c in enumerate(in_channels_list, 1):
            self.add_module('fpn_inner%d' % i, _fpn_conv(c, out_channels, 1))
            self.add_module('fpn_layer%d' % i, _fpn_conv(out_channels, out_channels, 3))


class _FSRelation(nn.Module):
    def __init__(self, scene_embedding_channels, in_channels_list, out_channels, scale_aware_proj=True):
        super().__init__()
        self.scale_aware_proj = bool(scale_aware_proj)
        if self.scale_aware_proj:   # one scene MLP per pyramid level (fs_relation.py:22-28)
            self.scene_encoder = nn.ModuleList([
                nn.Sequential(nn.Conv2d(scene_embedding_channels, out_channels, 1), nn.ReLU(True),
                              nn.Conv2d(out_channels, out_channels, 1)) for _ in in_channels_list])
        else:                       # a single MLP shared by all levels (fs_relation.py:29-35)
            self.scene_encoder = nn.Sequential(nn.Conv2d(scene_embedding_channels, out_channels, 1), nn.ReLU(True),
                                               nn.Conv2d(out_channels, out_channels, 1))
        self.content_encoders = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)) for c in in_channels_list])
        self.feature_reencoders = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)) for c in in_channels_list])


class _FSRelationV2(nn.Module):
    """Parameter container of FSRelationV2 (ever/module/fs_relation.py:76-139): scene encoder = conv1x1 -> GroupNorm(32) ->
    ReLU, twice; `project` = conv1x1(2C -> C, no bias) -> BN -> ReLU -> Dropout2d(0.1) on cat([r * p, p])."""
    version = 2
    dropout_p = 0.1

    def __init__(self, scene_embedding_channels, in_channels_list, out_channels, scale_aware_proj=True):
        super().__init__()
        self.scale_aware_proj = bool(scale_aware_proj)

        def scene():
            return nn.Sequential(nn.Conv2d(scene_embedding_channels, out_channels, 1), nn.GroupNorm(32, out_channels),
                                 nn.ReLU(True), nn.Conv2d(out_channels, out_channels, 1), nn.GroupNorm(32, out_channels),
                                 nn.ReLU(True))

        def project():
            return nn.Sequential(nn.Conv2d(out_channels * 2, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels),
                                 nn.ReLU(True), nn.Dropout2d(p=self.dropout_p))
        if self.scale_aware_proj:
            self.scene_encoder = nn.ModuleList([scene() for _ in in_channels_list])
            self.project = nn.ModuleList([project() for _ in in_channels_list])
        else:
            self.scene_encoder = scene()
            self.project = project()
        self.content_encoders = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)) for c in in_channels_list])
        self.feature_reencoders = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)) for c in in_channels_list])


class _Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, in_feat_output_strides=(4, 8, 16, 32), out_feat_output_stride=4,
                 classifier_config=None):
        super().__init__()
        cc = dict(classifier_config or {})
        self.num_classes = int(cc.get('num_classes', 1))
        self.scale_factor = int(cc.get('scale_factor', 1))
        ks = int(cc.get('kernel_size', 1))
        if ks not in (1, 3):
            raise NotImplementedError('classifier kernel_size 1 or 3 (got %d)' % ks)
        self.dropout_rate = float(cc.get('dropout_rate', -1))   # nn.Dropout on the merged features (fpn.py:175-176,190)
        self.num_upsample = []
        self.blocks = nn.ModuleList()
        for os_ in in_feat_output_strides:
            nup = int(math.log2(int(os_))) - int(math.log2(int(out_feat_output_stride)))
            self.num_upsample.append(nup)
            nl = nup if nup != 0 else 1
            self.blocks.append(nn.Sequential(*[
                nn.Sequential(nn.Conv2d(in_channels if j == 0 else out_channels, out_channels, 3, 1, 1, bias=False),
                              nn.BatchNorm2d(out_channels), nn.ReLU(True), nn.Identity()) for j in range(nl)]))
        self.dropout = nn.Dropout(self.dropout_rate) if self.dropout_rate > 0 else nn.Identity()
        self.classifier = nn.Sequential(nn.Conv2d(out_channels, self.num_classes, ks, padding=(ks - 1) // 2), nn.Identity())


class _Head(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.fpn = _FPN(tuple(cfg.fpn.in_channels_list), int(cfg.fpn.out_channels))
        fs = cfg.fs_relation
        # fs_relation.version = 2 selects FSRelationV2 (the reference's FarSegHead hard-wires FSRelation, fs_relation.py:171;
        # heads built on FSRelationV2 -- FarSeg++ -- construct it with the same keyword arguments)
        ver = int(fs.get('version', 1)) if hasattr(fs, 'get') else 1
        if ver not in (1, 2):
            raise ValueError('head.fs_relation.version must be 1 (FSRelation) or 2 (FSRelationV2)')
        rel_cls = _FSRelationV2 if ver == 2 else _FSRelation
        self.fs_relation = rel_cls(int(fs.scene_embedding_channels), tuple(fs.in_channels_list), int(fs.out_channels),
                                   bool(fs.scale_aware_proj))
        d = cfg.fpn_decoder
        # AssymetricDecoder(norm_fn=nn.BatchNorm2d, num_groups_gn=None) (fpn.py:145-170): only the BatchNorm + ReLU branch
        nf = d.get('norm_fn', None) if hasattr(d, 'get') else None
        if nf is not None and nf is not nn.BatchNorm2d:
            raise NotImplementedError('fpn_decoder.norm_fn other than nn.BatchNorm2d (GroupNorm / None + GELU branch of '
                                      'AssymetricDecoder, ever/module/fpn.py:166-167) is not built')
        self.fpn_decoder = _Decoder(int(d.in_channels), int(d.out_channels), tuple(d.in_feat_output_strides),
                                    int(d.out_feat_output_stride), d.classifier_config)


class _StepFn(torch.autograd.Function):
    """The whole native training step as ONE autograd node whose inputs are the trainable parameters.

    forward  = engine forward + loss (or the cached CUDA-graph replay); the returned losses carry a grad_fn, so the
    reference's stock path works unchanged: ``Launcher.compute_loss_gradient`` scales them (``v / forward_times``,
    ever/core/launcher.py:193-200), ``ERModule.backward`` sums and calls ``total_loss.backward()``
    (ever/interface/module.py:76-81), and under ``THDDPTrainer`` (ever/trainer/th_ddp_trainer.py:25-30) the
    ``DistributedDataParallel`` reducer sees every parameter's AccumulateGrad hook fire.
    backward = the engine's native backward with the upstream d(total)/d(loss) applied on the device, then one copy of
    the flat gradient arena whose per-parameter views are returned to autograd (which accumulates them into ``p.grad``:
    gradient accumulation over ``forward_times`` micro-batches is autograd's, the engine overwrites its arena)."""

    @staticmethod
    def forward(ctx, model, x, labels, *params):
        eng = model._engine()
        if bool(model.config.cuda_graph):
            out = eng.graph_forward(x, labels)
        else:
            out = eng.forward_train(x, labels)
        eng._step_token = getattr(eng, '_step_token', 0) + 1
        eng._last_keys = list(out.keys())
        ctx.model, ctx.keys, ctx.token = model, list(out.keys()), eng._step_token
        return tuple(out[k] for k in ctx.keys)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.model._engine()
        if eng._step_token != ctx.token or eng._saved_for_backward is None:
            raise RuntimeError('FarSegB200: backward through a stale step (the engine keeps the activations of the '
                               'latest training forward only)')
        for p in eng.params:   # a previous native step may have aliased .grad to the arena: autograd must not add into it
            if p.grad is not None and p.grad is eng._gv[id(p)]:
                p.grad = None
        if isinstance(eng._saved_for_backward, str):
            # graph mode: forward + loss + backward were replayed with unit loss weights; every loss is assumed to carry
            # the same upstream factor (what Launcher / GradScaler produce)
            eng.backward(allreduce=False, attach=False)
            s = next((g for g in gouts if g is not None), None)
            buf = eng.flat_g * s.to(torch.float32) if s is not None else torch.zeros_like(eng.flat_g)
        else:
            eng.backward(allreduce=False, attach=False, upstream=dict(zip(ctx.keys, gouts)))
            buf = eng.flat_g.clone()
        grads = tuple(buf[o:o + n].view_as(p) for p, (o, n) in zip(eng.params, eng._slots) if p.requires_grad)
        return (None, None, None) + grads


class NativeStepMixin:
    """forward / backward plumbing shared by the plugin models: the training forward is one autograd node (_StepFn) and
    ``backward`` takes the native or the autograd path (see FarSegB200.backward)"""

    def _train_forward(self, eng, x, labels):
        if torch.is_grad_enabled():
            tp = [p for p in eng.params if p.requires_grad]
            keys_vals = _StepFn.apply(self, x, labels, *tp)
            out = dict(zip(eng._last_keys, keys_vals))
        elif bool(self.config.cuda_graph):
            out = eng.graph_forward(x, labels)
        else:
            out = eng.forward_train(x, labels)
        object.__setattr__(self, '_last_out', out)
        return out

    def backward(self, loss_dict=None, amp=None, scaler=None, **kwargs):
        """ERModule.backward hook (ever/interface/module.py:76-81).

        Native path: ``loss_dict`` is None or the very dict values the last training forward returned -> the engine's
        backward runs with unit loss weights, the gradient all-reduce follows (world > 1) and every ``p.grad`` aliases its
        slot of the flat arena (StepLoop, bench, and any caller that wants the fused optimizer).
        Autograd path: anything else (Launcher hands over ``v / forward_times``; a GradScaler may scale) -> the reference's
        own ``sum(loss_dict.values()).backward()``, which reaches the engine through ``_StepFn.backward`` with the upstream
        factors; gradients accumulate into ordinary ``p.grad`` tensors and a DDP wrapper does its own all-reduce."""
        eng = self._engine()
        last = getattr(self, '_last_out', None)
        native = loss_dict is None or (last is not None and len(loss_dict) == len(last)
                                       and all(loss_dict.get(k) is v for k, v in last.items()))
        if native:
            eng.backward()
            return
        total_loss = sum([e for e in loss_dict.values()])
        if amp and scaler is not None:
            scaler.scale(total_loss).backward()
        else:
            total_loss.backward()

    def clip_grad_info(self):
        return dict()

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        object.__setattr__(self, 'engine', None)  # parameter storage moved: rebuild arenas lazily
        return r


@MODEL.register('FarSegB200')
class FarSegB200(NativeStepMixin, ERModule):
    """FarSeg (ResNetEncoder -> FarSegHead -> CE + Dice), glue model of SURVEY.md Appendix E, computed by the
    sm_100a engine.  forward(x, y): training -> {'ce_loss','dice_loss'}; eval -> softmax probabilities."""

    def __init__(self, config=None):
        super().__init__(config)
        enc = self.config.encoder
        self.en = _Encoder(enc.resnet_type, int(enc.in_channels))
        chans = self.en.resnet.out_channels
        head = self.config.head
        if tuple(head.fpn.in_channels_list) != tuple(chans):
            head.fpn.in_channels_list = tuple(chans)
            head.fs_relation.scene_embedding_channels = chans[-1]
        self.head = _Head(head)
        self.engine = None
        if int(enc.output_stride) != 32:
            raise NotImplementedError('FarSeg needs output_stride 32 (the FPN top-down x2 adds assume a /2 pyramid)')
        if not bool(enc.include_conv5):
            raise NotImplementedError('include_conv5=False: FarSegHead needs the four-level pyramid c2..c5')
        nl = enc.get('norm_layer', None) if hasattr(enc, 'get') else None
        if nl is not None and nl is not nn.BatchNorm2d:
            raise NotImplementedError('encoder.norm_layer other than nn.BatchNorm2d (got %r)' % (nl,))
        if bool(enc.pretrained):
            self._load_pretrained(enc.resnet_type, int(enc.in_channels))
        self._freeze()

    # ImageNet weights, same files as the reference (ever/module/_resnets.py:7-18)
    PRETRAINED_URLS = {
        'resnet18': 'https://download.pytorch.org/models/resnet18-5c106cde.pth',
        'resnet34': 'https://download.pytorch.org/models/resnet34-333f7ec4.pth',
        'resnet50': 'https://download.pytorch.org/models/resnet50-19c8e357.pth',
        'resnet101': 'https://download.pytorch.org/models/resnet101-5d3b4d8f.pth',
        'resnet50_v1c': 'https://download.openmmlab.com/pretrain/third_party/resnet50_v1c-2cccc1ad.pth',
        'resnet101_v1c': 'https://download.openmmlab.com/pretrain/third_party/resnet101_v1c-e67eebb6.pth',
        'resnext50_32x4d': 'https://download.pytorch.org/models/resnext50_32x4d-7cdf4587.pth',
        'resnext101_32x8d': 'https://download.pytorch.org/models/resnext101_32x8d-8ba56ff5.pth',
        'resnext101_32x4d': 'https://s3.ap-northeast-2.amazonaws.com/open-mmlab/pretrain/third_party/resnext101_32x4d-a5af3160.pth',
    }

    def _load_pretrained(self, resnet_type, in_channels, state_dict=None):
        """encoder.pretrained=True (ever/module/_resnets.py:230-238): fetch the reference's ImageNet checkpoint through the
        torch hub cache (raises without network / cache -- never a silent random init) and load it non-strictly (the
        checkpoint's fc.* keys have no counterpart).  in_channels != 3: the first conv is tiled over the new channels
        and rescaled, ResNetEncoder.patch_first_conv (ever/module/resnet.py:55-69)."""
        if state_dict is None:
            from torch.utils.model_zoo import load_url
            state_dict = load_url(self.PRETRAINED_URLS[resnet_type], progress=False)
        if 'state_dict' in state_dict:
            state_dict = state_dict['state_dict']
        state_dict = dict(state_dict)
        first = 'stem.0.weight' if self.en.resnet.deep_stem else 'conv1.weight'
        if in_channels != 3 and first in state_dict:
            w = state_dict[first]
            new = torch.stack([w[:, i % 3] for i in range(in_channels)], dim=1) * (3.0 / in_channels)
            state_dict[first] = new
        res = self.en.resnet.load_state_dict(state_dict, strict=False)
        missing = [k for k in res.missing_keys]
        if missing:
            raise RuntimeError('pretrained checkpoint lacks %d encoder tensors, e.g. %s' % (len(missing), missing[:3]))

    def _freeze(self):
        """ResNetEncoder._frozen_res_bn / _freeze_at (ever/module/resnet.py:155-173,227-234): frozen BN layers run on their
        running statistics and have frozen affine parameters; freeze_at >= k freezes the stem (1) and layer1..4 (2..5)."""
        enc, r = self.config.encoder, self.en.resnet
        if not bool(enc.batchnorm_trainable):
            for m in r.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    for p in m.parameters():
                        p.requires_grad = False
                    m.eval()
        if int(enc.freeze_at) >= 1 and r.deep_stem:
            # the reference's _freeze_at touches resnet.conv1 / bn1, which a deep-stem ResNet does not have (resnet.py:162-165)
            raise AttributeError("ResNetEncoder(freeze_at >= 1) with a deep-stem (v1c) ResNet: 'ResNet' object has no "
                                 "attribute 'conv1' (same failure as the reference, ever/module/resnet.py:164)")
        groups = [[] if r.deep_stem else [r.conv1, r.bn1], [r.layer1], [r.layer2], [r.layer3], [r.layer4]]
        for i, g in enumerate(groups, 1):
            if int(enc.freeze_at) >= i:
                for m in g:
                    for p in m.parameters():
                        p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self._freeze()
        return self

    def set_default_config(self):
        self.config.update(dict(
            # with_cp (activation checkpointing flags per stage, resnet.py:189-208) is accepted and has no effect: it
            # never changes results, and the engine keeps every activation of the step resident in HBM by design
            encoder=dict(resnet_type='resnet50', in_channels=3, pretrained=False, batchnorm_trainable=True, freeze_at=0,
                         output_stride=32, include_conv5=True, with_cp=(False, False, False, False), norm_layer=None),
            head=dict(
                fpn=dict(in_channels_list=(256, 512, 1024, 2048), out_channels=256),
                fs_relation=dict(scene_embedding_channels=2048, in_channels_list=(256, 256, 256, 256), out_channels=256,
                                 scale_aware_proj=True, version=1),
                fpn_decoder=dict(in_channels=256, out_channels=256, in_feat_output_strides=(4, 8, 16, 32),
                                 out_feat_output_stride=4,
                                 classifier_config=dict(scale_factor=4.0, num_classes=1, kernel_size=1))),
            loss=dict(ce=dict(weight=1.0), dice=dict(weight=1.0, smooth=1.0, sync_statistics=True), ignore_index=255),
            # uint8 HWC inputs are normalised on the fly (th_mean_std_normalize defaults, ever/preprocess/function.py:9)
            input=dict(mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375)),
            # static-shape training: forward() replays a cached CUDA graph of forward + loss + backward (one capture per
            # input signature); backward() then only all-reduces.  Off by default (eager launches, any shape).
            cuda_graph=False,
        ))

    # --------------------------------------------------------------------------------------------------
    def _engine(self):
        if self.engine is None:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('FarSegB200 computes only on a CUDA (sm_100a) device: move the model with .cuda(); '
                                   'there is no CPU path')
            # bypass nn.Module.__setattr__ bookkeeping for a plain python object
            object.__setattr__(self, 'engine', FarSegEngine(self))
        return self.engine

    def _labels(self, y):
        if y is None:
            raise ValueError('training forward needs y (dict with key "cls" or a label tensor)')
        return y['cls'] if isinstance(y, dict) else y

    def forward(self, x, y=None):
        eng = self._engine()
        if self.training:
            return self._train_forward(eng, x, self._labels(y))
        return eng.forward_eval(x)


MODEL.register('FarSeg', FarSegB200, override=True) if hasattr(MODEL, 'register') else None


class _ChangeMixin(nn.Module):
    """Parameter container of ChangeMixin (Z-Zheng/ChangeStar; not in the reference tree, see oracle/changestar_oracle.py)."""

    def __init__(self, in_channels, inner_channels=16, num_convs=4):
        super().__init__()
        layers = [nn.Sequential(nn.Conv2d(in_channels, inner_channels, 3, 1, 1, bias=False), nn.BatchNorm2d(inner_channels),
                                nn.ReLU(True))]
        layers += [nn.Sequential(nn.Conv2d(inner_channels, inner_channels, 3, 1, 1, bias=False),
                                 nn.BatchNorm2d(inner_channels), nn.ReLU(True)) for _ in range(num_convs - 1)]
        layers += [nn.Conv2d(inner_channels, 1, 3, 1, 1), nn.Identity()]
        self.convs = nn.Sequential(*layers)


@MODEL.register('ChangeStarB200')
class ChangeStarB200(FarSegB200):
    """ChangeStar = FarSeg feature extractor on both temporal images + ChangeMixin (bitemporal, temporally symmetric)
    change head.  forward(x[N, 2*Cin, H, W], y={'cls': t1 labels, 'change': binary change labels}) ->
    {'ce_loss'|'bce_loss', 'dice_loss', 'c12_bce_loss', 'c12_dice_loss', 'c21_bce_loss', 'c21_dice_loss'};
    eval -> {'seg': probabilities of t1, 'change': sigmoid(c12)}."""

    def __init__(self, config=None):
        super().__init__(config)
        cdec = int(self.config.head.fpn_decoder.out_channels)
        cm = self.config.changemixin
        if int(cm.num_convs) != 4 or float(cm.scale_factor) != 4.0:
            raise NotImplementedError('ChangeMixin with num_convs=4, scale_factor=4')
        self.changemixin = _ChangeMixin(2 * cdec, int(cm.inner_channels), int(cm.num_convs))

    def set_default_config(self):
        super().set_default_config()
        self.config.update(dict(changemixin=dict(inner_channels=16, num_convs=4, scale_factor=4.0)))

    def _engine(self):
        if self.engine is None:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('ChangeStarB200 computes only on a CUDA (sm_100a) device; there is no CPU path')
            object.__setattr__(self, 'engine', ChangeStarEngine(self))
        return self.engine

    def _labels(self, y):
        if not isinstance(y, dict) or 'cls' not in y or 'change' not in y:
            raise ValueError("training forward needs y = {'cls': ..., 'change': ...}")
        return y


MODEL.register('ChangeStar', ChangeStarB200, override=True) if hasattr(MODEL, 'register') else None


from . import freenet  # noqa: E402,F401  (registers 'FreeNetB200' / 'FreeNet'; imports NativeStepMixin from this module)
