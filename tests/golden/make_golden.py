"""Generate golden fixtures from the REAL reference (/root/reference), in the build container.

    PYTHONPATH=/root/reference:tests/golden/_stubs python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these fixtures are outputs of the
unmodified reference modules (ever.module.ResNetEncoder + ever.module.FarSegHead +
F.cross_entropy + ever.module.loss.dice_loss_with_logits) on seeded synthetic tiles with
deterministic per-key weights (oracle.farseg_oracle.deterministic_fill).  They are small (slices,
norms, masks) so they can be committed; tests rebuild inputs/weights from the same seeds.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

import ever as er  # noqa: E402  (the real reference)
import ever.module as erm  # noqa: E402
from ever.module.loss import dice_loss_with_logits  # noqa: E402
from oracle.farseg_oracle import deterministic_fill, synthetic_batch  # noqa: E402

CASES = {
    # name: (resnet, K, N, H, W, decoder_channels)
    'r18_k5_2x64': ('resnet18', 5, 2, 64, 64, 128),
    'r50_k15_1x64': ('resnet50', 15, 1, 64, 64, 256),
    'r18_k1_2x64': ('resnet18', 1, 2, 64, 64, 128),
    # ResNetEncoder(in_channels=8) (resnet.py:100-117: fresh 7x7 conv) + FSRelation(scale_aware_proj=False) (fs_relation.py:29-35)
    # deep-stem ResNet (three 3x3 convs, _resnets.py:137-147)
    'r50v1c_k5_1x64': ('resnet50_v1c', 5, 1, 64, 64, 128),
    'r18_k5_c8_shared_2x64': ('resnet18', 5, 2, 64, 64, 128, dict(in_channels=8, scale_aware_proj=False)),
    # FSRelationV2 (fs_relation.py:76-163) swapped into the head; Dropout2d draws from torch.manual_seed(DROP_SEED)
    'r18_k5_v2_2x64': ('resnet18', 5, 2, 64, 64, 128, dict(fs_version=2)),
    # ResNeXt: grouped 3x3 (32 groups x 4 channels) in every bottleneck (_resnets.py:80-84,291-300)
    'rx50_k5_1x64': ('resnext50_32x4d', 5, 1, 64, 64, 128),
}
DROP_SEED = 20240


class RefFarSeg(er.ERModule):
    def __init__(self, config):
        super().__init__(config)
        self.en = erm.ResNetEncoder(self.config.encoder)
        self.head = erm.FarSegHead(self.config.head)
        if int(self.config.get('fs_version', 1)) == 2:   # the reference's own FSRelationV2 in place of FSRelation
            from ever.module.fs_relation import FSRelationV2
            self.head.fs_relation = FSRelationV2(**self.head.config.fs_relation)

    def forward(self, x, y=None):
        logit = self.head(self.en(x))
        if self.training:
            return dict(ce_loss=F.cross_entropy(logit, y['cls'].long(), ignore_index=255),
                        dice_loss=dice_loss_with_logits(logit, y['cls'], ignore_index=255)), logit
        return logit.softmax(dim=1)

    def set_default_config(self):
        self.config.update(dict(encoder=dict(), head=dict()))


def ref_config(resnet, k, dec, in_channels=3, scale_aware_proj=True, fs_version=1):
    chans = (64, 128, 256, 512) if resnet in ('resnet18', 'resnet34') else (256, 512, 1024, 2048)
    return dict(fs_version=fs_version, encoder=dict(resnet_type=resnet, in_channels=in_channels),
                head=dict(fpn=dict(in_channels_list=chans, out_channels=256),
                          fs_relation=dict(scene_embedding_channels=chans[-1], scale_aware_proj=scale_aware_proj),
                          fpn_decoder=dict(out_channels=dec, classifier_config=dict(num_classes=k, scale_factor=4.0, kernel_size=1))))


def run_case(name):
    resnet, k, n, h, w, dec = CASES[name][:6]
    opts = CASES[name][6] if len(CASES[name]) > 6 else {}
    m = RefFarSeg(ref_config(resnet, k, dec, **opts))
    deterministic_fill(m, seed=0)
    x, y = synthetic_batch(n, h, w, max(k, 2), in_channels=opts.get('in_channels', 3))
    if k == 1:
        # binary case exercises the sigmoid branch of dice (loss.py:66-68); CE is replaced by masked BCE
        from ever.module.loss import binary_cross_entropy_with_logits
        m.train()
        logit = m.head(m.en(x))
        losses = dict(bce_loss=binary_cross_entropy_with_logits(logit, y, ignore_index=255),
                      dice_loss=dice_loss_with_logits(logit, y, ignore_index=255))
    else:
        m.train()
        torch.manual_seed(DROP_SEED)
        losses, logit = m(x, dict(cls=y))
    sum(losses.values()).backward()
    gsum = {kk: float(p.grad.double().sum()) for kk, p in m.named_parameters()}
    gnorm = {kk: float(p.grad.double().norm()) for kk, p in m.named_parameters()}
    bn_after = {kk: v.clone() for kk, v in m.state_dict().items()
                if 'running' in kk and (('bn1.' in kk or 'stem.' in kk) and 'layer' not in kk)}
    m.eval()
    with torch.no_grad():
        prob = m(x)
    out = dict(case=CASES[name], losses={kk: float(v) for kk, v in losses.items()},
               logit_slice=logit.detach()[:, :, ::4, ::4].clone(), grad_sum=gsum, grad_norm=gnorm,
               stem_bn_running=bn_after,
               eval_mask=(prob.argmax(dim=1).to(torch.uint8) if k > 1 else (prob > 0.5).to(torch.uint8)),
               eval_prob_slice=prob[:, :, ::4, ::4].clone(), torch_version=torch.__version__)
    torch.save(out, os.path.join(HERE, name + '.pt'))
    print(name, out['losses'], 'params', len(gsum))


if __name__ == '__main__':
    torch.set_num_threads(8)
    for c in (sys.argv[1:] or CASES):
        run_case(c)
