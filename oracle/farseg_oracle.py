"""ORACLE (test infrastructure, NOT product code).

Plain-PyTorch restatement of the reference hot path: ResNet encoder -> FPN ->
FS-Relation -> asymmetric decoder -> CE + Dice, i.e. the "glue" FarSeg model of
SURVEY.md Appendix E.  The arithmetic of the reference for this path lives in
PyTorch itself (the reference ships no kernels), so the restatement is a set of
torch modules issuing the same ATen ops in the same order, with the same
``state_dict`` keys.  Every class cites the reference file:line it follows.

Pinned by ``tests/test_oracle_vs_reference.py`` (bit-exact against the real
reference imported from /root/reference, run in the build container) and by the
committed fixtures ``tests/golden/*.pt`` produced from the real reference by
``tests/golden/make_golden.py``.  The reference has no tests / golden vectors of
its own, so those fixtures are the pin.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# ResNet (reference: ever/module/_resnets.py)
# ----------------------------------------------------------------------------


class _Basic(nn.Module):
    """BasicBlock, reference ever/module/_resnets.py:32-69."""
    expansion = 1

    def __init__(self, cin, planes, stride, down):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = down

    def forward(self, x):
        idt = x
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        if self.downsample is not None:
            idt = self.downsample(x)
        y += idt
        return self.relu(y)


class _Bottle(nn.Module):
    """Bottleneck (stride on the 3x3), reference ever/module/_resnets.py:72-112."""
    expansion = 4

    def __init__(self, cin, planes, stride, down, groups=1, base_width=64):
        super().__init__()
        width = int(planes * (base_width / 64.)) * groups   # ResNeXt: _resnets.py:80-84
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = down

    def forward(self, x):
        idt = x
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        if self.downsample is not None:
            idt = self.downsample(x)
        y += idt
        return self.relu(y)


RESNET_SPECS = {
    'resnet18': (_Basic, (2, 2, 2, 2)),
    'resnet34': (_Basic, (3, 4, 6, 3)),
    'resnet50': (_Bottle, (3, 4, 6, 3)),
    'resnet101': (_Bottle, (3, 4, 23, 3)),
    'resnet50_v1c': (_Bottle, (3, 4, 6, 3)),    # deep stem, reference ever/module/_resnets.py:327-345
    'resnet101_v1c': (_Bottle, (3, 4, 23, 3)),
    'resnext50_32x4d': (_Bottle, (3, 4, 6, 3)),   # reference ever/module/_resnets.py:291-324
    'resnext101_32x4d': (_Bottle, (3, 4, 23, 3)),
    'resnext101_32x8d': (_Bottle, (3, 4, 23, 3)),
}
RESNEXT_SPECS = {'resnext50_32x4d': (32, 4), 'resnext101_32x4d': (32, 4), 'resnext101_32x8d': (32, 8)}   # groups, width_per_group


class _ResNetTrunk(nn.Module):
    """ResNet without fc, reference ever/module/_resnets.py:115-227 (ctor, init :163-169,
    _make_layer :181-203, stem_forward :205-212)."""

    def __init__(self, kind, in_channels=3):
        super().__init__()
        block, counts = RESNET_SPECS[kind]
        self.deep_stem = kind.endswith('_v1c')
        if self.deep_stem:   # three 3x3 convs, reference ever/module/_resnets.py:137-147
            self.stem = nn.Sequential(
                nn.Conv2d(in_channels, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
                nn.Conv2d(32, 32, 3, 1, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
                nn.Conv2d(32, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True))
        else:
            self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
            self.bn1 = nn.BatchNorm2d(64)
            self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), counts), 1):
            stride = 1 if li == 1 else 2
            blocks = []
            for b in range(n):
                s = stride if b == 0 else 1
                down = None
                if b == 0 and (s != 1 or cin != planes * block.expansion):
                    down = nn.Sequential(nn.Conv2d(cin, planes * block.expansion, 1, s, bias=False),
                                         nn.BatchNorm2d(planes * block.expansion))
                blocks.append(block(cin, planes, s, down, *RESNEXT_SPECS[kind]) if kind in RESNEXT_SPECS
                              else block(cin, planes, s, down))
                cin = planes * block.expansion
            setattr(self, 'layer%d' % li, nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))  # parameter-free, kept for module parity
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class ResNetEncoderOracle(nn.Module):
    """ResNetEncoder.forward, reference ever/module/resnet.py:183-211 (defaults :213-225:
    output_stride 32, include_conv5, trainable BN, freeze_at 0)."""

    def __init__(self, resnet_type='resnet50', in_channels=3, freeze_at=0, batchnorm_trainable=True):
        super().__init__()
        self.resnet = _ResNetTrunk(resnet_type, in_channels)
        self.out_channels = tuple(c * RESNET_SPECS[resnet_type][0].expansion for c in (64, 128, 256, 512))
        self.freeze_at, self.batchnorm_trainable = freeze_at, batchnorm_trainable
        self._freeze()

    def _freeze(self):
        """_frozen_res_bn / _freeze_at, reference ever/module/resnet.py:155-173 (param_util.freeze_params sets
        requires_grad=False; frozen BN additionally runs in eval mode, :227-234)."""
        r = self.resnet
        if not self.batchnorm_trainable:
            for m in r.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad = False
                    m.eval()
        groups = [[] if r.deep_stem else [r.conv1, r.bn1], [r.layer1], [r.layer2], [r.layer3], [r.layer4]]
        for i, g in enumerate(groups, 1):
            if self.freeze_at >= i:
                for m in g:
                    for p in m.parameters():
                        p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self._freeze()
        return self

    def forward(self, x):
        r = self.resnet
        x = r.stem(x) if r.deep_stem else r.relu(r.bn1(r.conv1(x)))   # stem_forward, _resnets.py:205-212
        x = r.maxpool(x)
        c2 = r.layer1(x)
        c3 = r.layer2(c2)
        c4 = r.layer3(c3)
        c5 = r.layer4(c4)
        return [c2, c3, c4, c5]


# ----------------------------------------------------------------------------
# FPN / decoder (reference: ever/module/fpn.py, ever/module/ops.py)
# ----------------------------------------------------------------------------


def _plain_conv(cin, cout, k):
    """ConvBlock(bn=False, relu=False, bias=False) with kaiming_uniform(a=1):
    reference ever/module/ops.py:45-60, ever/module/fpn.py:18-35."""
    seq = nn.Sequential(nn.Conv2d(cin, cout, k, 1, (k - 1) // 2, bias=False), nn.Identity(), nn.Identity())
    nn.init.kaiming_uniform_(seq[0].weight, a=1)
    return seq


class _Fp32Around(nn.Module):
    """Bf16compatible, reference ever/module/ops.py:152-166."""

    def __init__(self, module):
        super().__init__()
        self._inner_module = module

    def forward(self, x):
        dt = x.dtype
        if dt == torch.bfloat16:
            x = x.float()
        x = self._inner_module(x)
        return x.to(dt) if dt == torch.bfloat16 else x


class FPNOracle(nn.Module):
    """FPN.forward, reference ever/module/fpn.py:80-115 (nearest x2 top-down, done in fp32 for
    bf16 inputs :96-102)."""

    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        self.n = len(in_channels_list)
        for i, c in enumerate(in_channels_list, 1):
            # registration order inner_i, layer_i matters for parameter ordering (fpn.py:66-76)
            self.add_module('fpn_inner%d' % i, _plain_conv(c, out_channels, 1))
            self.add_module('fpn_layer%d' % i, _plain_conv(out_channels, out_channels, 3))

    def forward(self, feats):
        last = getattr(self, 'fpn_inner%d' % self.n)(feats[-1])
        outs = [getattr(self, 'fpn_layer%d' % self.n)(last)]
        for i in range(self.n - 1, 0, -1):
            dt = last.dtype
            up = last.float() if dt == torch.bfloat16 else last
            up = F.interpolate(up, scale_factor=2, mode='nearest')
            if dt == torch.bfloat16:
                up = up.to(dt)
            lat = getattr(self, 'fpn_inner%d' % i)(feats[i - 1])
            last = lat + up
            outs.insert(0, getattr(self, 'fpn_layer%d' % i)(last))
        return tuple(outs)


class FSRelationOracle(nn.Module):
    """FSRelation, reference ever/module/fs_relation.py:14-73: one scene MLP per pyramid level when
    scale_aware_proj (:22-28, the FarSegHead default :193), else a single shared MLP (:29-35)."""

    def __init__(self, scene_embedding_channels, in_channels_list, out_channels, scale_aware_proj=True):
        super().__init__()
        self.scale_aware_proj = scale_aware_proj
        if scale_aware_proj:
            self.scene_encoder = nn.ModuleList([
                nn.Sequential(nn.Conv2d(scene_embedding_channels, out_channels, 1), nn.ReLU(True),
                              nn.Conv2d(out_channels, out_channels, 1)) for _ in in_channels_list])
        else:
            self.scene_encoder = nn.Sequential(nn.Conv2d(scene_embedding_channels, out_channels, 1), nn.ReLU(True),
                                               nn.Conv2d(out_channels, out_channels, 1))
        self.content_encoders = nn.ModuleList()
        self.feature_reencoders = nn.ModuleList()
        for c in in_channels_list:
            self.content_encoders.append(nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)))
            self.feature_reencoders.append(nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)))

    def forward(self, scene, feats):
        cfs = [enc(p) for enc, p in zip(self.content_encoders, feats)]
        if self.scale_aware_proj:
            sfs = [enc(scene) for enc in self.scene_encoder]
            rel = [torch.sigmoid((sf * cf).sum(dim=1, keepdim=True)) for sf, cf in zip(sfs, cfs)]
        else:
            sf = self.scene_encoder(scene)
            rel = [torch.sigmoid((sf * cf).sum(dim=1, keepdim=True)) for cf in cfs]
        pfs = [enc(p) for enc, p in zip(self.feature_reencoders, feats)]
        return [r * p for r, p in zip(rel, pfs)]


class FSRelationV2Oracle(nn.Module):
    """FSRelationV2, reference ever/module/fs_relation.py:76-163: scene encoder = (1x1 conv, GroupNorm(32), ReLU) x 2
    (:86-96 / :107-114), relation as in FSRelation (:142-152), refined = cat([r * p, o]) (:156), `project` = 1x1 conv
    (2C -> C, no bias) + BN + ReLU + Dropout2d(0.1) (:97-104 / :115-120, applied :158-161)."""

    def __init__(self, scene_embedding_channels, in_channels_list, out_channels, scale_aware_proj=True):
        super().__init__()
        self.scale_aware_proj = scale_aware_proj

        def scene():
            return nn.Sequential(nn.Conv2d(scene_embedding_channels, out_channels, 1), nn.GroupNorm(32, out_channels),
                                 nn.ReLU(True), nn.Conv2d(out_channels, out_channels, 1), nn.GroupNorm(32, out_channels),
                                 nn.ReLU(True))

        def project():
            return nn.Sequential(nn.Conv2d(out_channels * 2, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels),
                                 nn.ReLU(True), nn.Dropout2d(p=0.1))
        if scale_aware_proj:
            self.scene_encoder = nn.ModuleList([scene() for _ in in_channels_list])
            self.project = nn.ModuleList([project() for _ in in_channels_list])
        else:
            self.scene_encoder = scene()
            self.project = project()
        self.content_encoders = nn.ModuleList()
        self.feature_reencoders = nn.ModuleList()
        for c in in_channels_list:
            self.content_encoders.append(nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)))
            self.feature_reencoders.append(nn.Sequential(nn.Conv2d(c, out_channels, 1), nn.BatchNorm2d(out_channels), nn.ReLU(True)))

    def forward(self, scene, feats):
        cfs = [enc(p) for enc, p in zip(self.content_encoders, feats)]
        if self.scale_aware_proj:
            sfs = [enc(scene) for enc in self.scene_encoder]
            rel = [torch.sigmoid((sf * cf).sum(dim=1, keepdim=True)) for sf, cf in zip(sfs, cfs)]
        else:
            sf = self.scene_encoder(scene)
            rel = [torch.sigmoid((sf * cf).sum(dim=1, keepdim=True)) for cf in cfs]
        pfs = [enc(p) for enc, p in zip(self.feature_reencoders, feats)]
        refined = [torch.cat([r * p, o], dim=1) for r, p, o in zip(rel, pfs, feats)]
        if self.scale_aware_proj:
            return [op(x) for op, x in zip(self.project, refined)]
        return [self.project(x) for x in refined]


class DecoderOracle(nn.Module):
    """AssymetricDecoder, reference ever/module/fpn.py:144-193."""

    def __init__(self, in_channels, out_channels, in_feat_output_strides=(4, 8, 16, 32), out_feat_output_stride=4,
                 num_classes=1, scale_factor=4.0, kernel_size=1, dropout_rate=-1):
        super().__init__()
        self.blocks = nn.ModuleList()
        for os_ in in_feat_output_strides:
            nup = int(math.log2(int(os_))) - int(math.log2(int(out_feat_output_stride)))
            nl = nup if nup != 0 else 1
            self.blocks.append(nn.Sequential(*[
                nn.Sequential(
                    nn.Conv2d(in_channels if j == 0 else out_channels, out_channels, 3, 1, 1, bias=False),
                    nn.BatchNorm2d(out_channels),
                    nn.ReLU(True),
                    _Fp32Around(nn.UpsamplingBilinear2d(scale_factor=2)) if nup != 0 else nn.Identity())
                for j in range(nl)]))
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else nn.Identity()   # fpn.py:175-176
        self.classifier = nn.Sequential(
            nn.Conv2d(out_channels, num_classes, kernel_size, padding=(kernel_size - 1) // 2),
            _Fp32Around(nn.UpsamplingBilinear2d(scale_factor=scale_factor)) if scale_factor > 1 else nn.Identity())

    def forward(self, feats):
        inner = [blk(f) for blk, f in zip(self.blocks, feats)]
        out = sum(inner) / len(inner)
        return self.classifier(self.dropout(out))


class FarSegHeadOracle(nn.Module):
    """FarSegHead.forward, reference ever/module/fs_relation.py:166-206."""

    def __init__(self, in_channels_list=(256, 512, 1024, 2048), fpn_channels=256, decoder_channels=256, num_classes=1,
                 scale_aware_proj=True, classifier_kernel_size=1, fs_version=1, classifier_dropout=-1):
        super().__init__()
        self.fpn = FPNOracle(in_channels_list, fpn_channels)
        rel_cls = FSRelationV2Oracle if fs_version == 2 else FSRelationOracle
        self.fs_relation = rel_cls(in_channels_list[-1], (fpn_channels,) * 4, fpn_channels, scale_aware_proj)
        self.fpn_decoder = DecoderOracle(fpn_channels, decoder_channels, num_classes=num_classes,
                                         kernel_size=classifier_kernel_size, dropout_rate=classifier_dropout)

    def forward(self, feats):
        ps = self.fpn(feats)
        scene = F.adaptive_avg_pool2d(feats[-1], 1)
        return self.fpn_decoder(self.fs_relation(scene, ps))


# ----------------------------------------------------------------------------
# Losses (reference: ever/module/loss.py)
# ----------------------------------------------------------------------------


def dice_loss_oracle(logit, target, smooth=1.0, ignore_index=255, all_reduce=None):
    """dice_loss_with_logits, reference ever/module/loss.py:54-75 with select :26-37 and
    dice_coeff :40-51.  ``all_reduce`` (callable) stands for all_reduce_sum :20-23."""
    c = logit.size(1)
    flat = logit.permute(0, 2, 3, 1).reshape(-1, c)
    t = target.reshape(-1)
    valid = t != ignore_index
    flat, t = flat[valid, :], t[valid]
    if c == 1:
        prob = flat.sigmoid()
        onehot = t.reshape(-1, 1)
    else:
        prob = flat.log_softmax(dim=1).exp()
        onehot = F.one_hot(t.long(), num_classes=c).type_as(flat)
    inter = torch.sum(prob * onehot, dim=0)
    z = prob.sum(dim=0) + onehot.sum(dim=0)
    if all_reduce is not None:
        inter, z = all_reduce(inter), all_reduce(z)
    z = z + smooth
    return 1. - ((2 * inter + smooth) / z).mean()


def bce_loss_oracle(logit, target, ignore_index=255):
    """binary_cross_entropy_with_logits, reference ever/module/loss.py:229-235 (+ :10-17)."""
    p, t = logit.reshape(-1), target.reshape(-1)
    valid = t != ignore_index
    return F.binary_cross_entropy_with_logits(p.masked_select(valid).float(), t.masked_select(valid).float())


class FarSegOracle(nn.Module):
    """The glue ERModule of SURVEY.md Appendix E: encoder -> FarSegHead -> {ce_loss, dice_loss}
    in training, softmax probabilities in eval."""

    def __init__(self, resnet_type='resnet50', num_classes=15, decoder_channels=256, in_channels=3, freeze_at=0,
                 batchnorm_trainable=True, scale_aware_proj=True, classifier_kernel_size=1, fs_version=1,
                 classifier_dropout=-1):
        super().__init__()
        self.en = ResNetEncoderOracle(resnet_type, in_channels, freeze_at, batchnorm_trainable)
        self.head = FarSegHeadOracle(self.en.out_channels, 256, decoder_channels, num_classes, scale_aware_proj,
                                     classifier_kernel_size, fs_version, classifier_dropout)
        self.dice_all_reduce = None

    def logits(self, x):
        return self.head(self.en(x))

    def forward(self, x, y=None):
        logit = self.logits(x)
        if self.training:
            cls = y['cls']
            return dict(ce_loss=F.cross_entropy(logit, cls.long(), ignore_index=255),
                        dice_loss=dice_loss_oracle(logit, cls, ignore_index=255, all_reduce=self.dice_all_reduce))
        return logit.softmax(dim=1)


# ----------------------------------------------------------------------------
# Synthetic inputs
# ----------------------------------------------------------------------------


def synthetic_batch(n, h, w, num_classes, in_channels=3, ignore_frac=0.05, seed_offset=0):
    g = torch.Generator().manual_seed(1234 + seed_offset)
    x = torch.randn(n, in_channels, h, w, generator=g)
    g = torch.Generator().manual_seed(4321 + seed_offset)
    y = torch.randint(0, num_classes, (n, h, w), generator=g)
    g = torch.Generator().manual_seed(99 + seed_offset)
    y[torch.rand(n, h, w, generator=g) < ignore_frac] = 255
    return x, y


def build_oracle(resnet_type, num_classes, decoder_channels=256, seed=0):
    torch.manual_seed(seed)
    return FarSegOracle(resnet_type, num_classes, decoder_channels)


def deterministic_fill(model, seed=0):
    """Fill every parameter/buffer from a per-key seeded generator so that the real reference
    and this restatement get identical weights independently of construction order (the reference
    ResNet also builds an ``fc`` layer that consumes RNG, ever/module/_resnets.py:161)."""
    import zlib
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) % (2 ** 31))
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(v)
        elif k.endswith('running_mean'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('running_var'):
            out[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif v.dim() == 4:  # conv weight: kaiming-normal-like, fan_in scaling keeps activations O(1)
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            out[k] = torch.randn(v.shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif k.endswith('weight'):  # BN gamma
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:  # BN beta / conv bias
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
    model.load_state_dict(out, strict=True)
    return model
