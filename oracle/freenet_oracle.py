"""ORACLE (test infrastructure, NOT product code) -- FreeNet.

PARITY UNPINNED for the network as a whole: FreeNet is NOT in /root/reference (only linked from its README.md:55, SURVEY.md
row a14).  This is a restatement, from the published description of Z-Zheng/FreeNet ("FPGA: Fast Patch-Free Global Learning
Framework for Fully End-to-End Hyperspectral Image Classification", TGRS 2020), of its encoder-decoder:

  encoder   conv3x3_gn_relu(Cin -> c1); per stage: [SEBlock(c, r) -> conv3x3_gn_relu(c, c)] x num_blocks, tapped;
            between stages downsample2x = 3x3 stride-2 conv (+bias) -> ReLU;  block_channels (96, 128, 192, 256), r = 16
  decoder   reduce_1x1convs (c_i -> inner_dim, +bias); top-down: inner_i = reduce(feat_i) + nearest_x2(out_{i+1});
            out_i = fuse_3x3convs(inner_i) (3x3, +bias);  cls_pred_conv 1x1 (inner_dim -> K) at full resolution
  loss      sum(CE(logit, y - 1, ignore -1) * w) / sum(w)   (y: 1..K, 0 = unlabelled; w: the training-pixel mask)

Its building blocks ARE in the reference tree and are used as the pins: ``SEBlock`` restates ever/module/se_block.py:9-24
(checked bit-exact against the real class in tests/test_oracle.py), GroupNorm / Conv2d / nearest interpolation are torch's.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / incumbent / cpu_baseline legs may import this.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SEBlockOracle(nn.Module):
    """SEBlock, reference ever/module/se_block.py:9-24: GAP -> Linear(c, c // r) -> ReLU -> Linear -> Sigmoid -> scale."""

    def __init__(self, in_channels, reduction):
        super().__init__()
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.seq = nn.Sequential(nn.Linear(in_channels, in_channels // reduction), nn.ReLU(inplace=True),
                                 nn.Linear(in_channels // reduction, in_channels), nn.Sigmoid())

    def forward(self, x):
        v = self.gap(x)
        score = self.seq(v.view(v.size(0), v.size(1)))
        return x * score.view(score.size(0), score.size(1), 1, 1)


def conv3x3_gn_relu(cin, cout, groups):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1), nn.GroupNorm(groups, cout), nn.ReLU(inplace=True))


def downsample2x(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1), nn.ReLU(inplace=True))


def repeat_block(c, r, n, se_cls=SEBlockOracle):
    return nn.Sequential(*[nn.Sequential(se_cls(c, r), conv3x3_gn_relu(c, c, r)) for _ in range(n)])


class FreeNetOracle(nn.Module):
    def __init__(self, in_channels=200, num_classes=9, block_channels=(96, 128, 192, 256), num_blocks=(1, 1, 1, 1),
                 inner_dim=128, reduction_ratio=1.0, se_cls=SEBlockOracle):
        super().__init__()
        r = int(16 * reduction_ratio)
        ch = [int(c * reduction_ratio / r) * r for c in block_channels]
        ops = [conv3x3_gn_relu(in_channels, ch[0], r), repeat_block(ch[0], r, num_blocks[0], se_cls), nn.Identity()]
        for i in range(1, 4):
            ops += [downsample2x(ch[i - 1], ch[i]), repeat_block(ch[i], r, num_blocks[i], se_cls), nn.Identity()]
        self.feature_ops = nn.ModuleList(ops)
        inner = int(inner_dim * reduction_ratio)
        self.reduce_1x1convs = nn.ModuleList([nn.Conv2d(c, inner, 1) for c in ch])
        self.fuse_3x3convs = nn.ModuleList([nn.Conv2d(inner, inner, 3, 1, 1) for _ in ch])
        self.cls_pred_conv = nn.Conv2d(inner, num_classes, 1)

    def logits(self, x):
        feats = []
        for op in self.feature_ops:
            x = op(x)
            if isinstance(op, nn.Identity):
                feats.append(x)
        inner = [conv(f) for conv, f in zip(self.reduce_1x1convs, feats)]
        inner.reverse()
        out = self.fuse_3x3convs[0](inner[0])
        for i in range(len(inner) - 1):
            top2x = F.interpolate(out, scale_factor=2.0, mode='nearest')
            out = self.fuse_3x3convs[i + 1](inner[i + 1] + top2x)
        return self.cls_pred_conv(out)

    def forward(self, x, y=None, w=None):
        logit = self.logits(x)
        if self.training:
            if isinstance(y, dict):
                y, w = y['cls'], y['w']
            losses = F.cross_entropy(logit, y.long() - 1, ignore_index=-1, reduction='none')
            return dict(cls_loss=(losses * w).sum() / w.sum())
        return logit.softmax(dim=1)


def synthetic_cube(n, c, h, w, num_classes, seed=0, labelled_frac=0.05):
    """a hyperspectral-cube-shaped batch: randn spectra, labels 1..K on a sparse set of pixels (0 = unlabelled) and the
    training mask w that selects them (FreeNet trains on a few hundred labelled pixels of ONE image)"""
    g = torch.Generator().manual_seed(2020 + seed)
    x = torch.randn(n, c, h, w, generator=g)
    y = torch.randint(1, num_classes + 1, (n, h, w), generator=g)
    wmask = (torch.rand(n, h, w, generator=g) < labelled_frac).float()
    y = y * (wmask > 0).long()
    return x, y, wmask
