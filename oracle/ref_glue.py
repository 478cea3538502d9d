"""ORACLE / reference-arm support (test infrastructure, NOT product code): import the UNMODIFIED reference package and
compose its own classes into the glue FarSeg model of SURVEY.md Appendix E.

The reference is installed once, offline, with
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``
(git-ignored; travels to the GPU box with the snapshot); in the build container /root/reference itself is the fallback.
``prettytable`` / ``albumentations`` (import-time dependencies that are not installed here) are satisfied by the two stubs
under tests/golden/_stubs.  Only tests/, __graft_entry__.smoke() and bench.py's reference / incumbent /
cpu_baseline legs may import this module.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, 'baseline', '_ref'), '/root/reference']
STUBS = os.path.join(ROOT, 'tests', 'golden', '_stubs')


def reference_available():
    return any(os.path.isdir(os.path.join(d, 'ever')) for d in REF_DIRS)


def import_reference():
    """put the vendored reference (and the two import stubs) on sys.path and import it; returns the ``ever`` package"""
    for d in REF_DIRS:
        if os.path.isdir(os.path.join(d, 'ever')):
            if d not in sys.path:
                sys.path.insert(0, d)
            break
    else:
        raise ImportError('reference not installed: baseline/_ref/ever missing')
    if STUBS not in sys.path:
        sys.path.insert(0, STUBS)
    import ever
    return ever


def make_reference_farseg(resnet='resnet50', num_classes=15, decoder_channels=256, **encoder_opts):
    """The glue ERModule of SURVEY.md Appendix E, composed of the reference's OWN classes: ResNetEncoder
    (ever/module/resnet.py), FarSegHead (ever/module/fs_relation.py), F.cross_entropy + dice_loss_with_logits
    (ever/module/loss.py:54-75)."""
    er = import_reference()
    import torch.nn.functional as F
    import ever.module as erm
    from ever.module.loss import dice_loss_with_logits

    class RefFarSeg(er.ERModule):
        def __init__(self, config):
            super().__init__(config)
            self.en = erm.ResNetEncoder(self.config.encoder)
            self.head = erm.FarSegHead(self.config.head)

        def forward(self, x, y=None):
            logit = self.head(self.en(x))
            if self.training:
                return dict(ce_loss=F.cross_entropy(logit, y['cls'].long(), ignore_index=255),
                            dice_loss=dice_loss_with_logits(logit, y['cls'], ignore_index=255))
            return logit.softmax(dim=1)

        def set_default_config(self):
            self.config.update(dict(encoder=dict(), head=dict()))

    chans = (64, 128, 256, 512) if resnet in ('resnet18', 'resnet34') else (256, 512, 1024, 2048)
    enc = dict(resnet_type=resnet)
    enc.update(encoder_opts)
    cfg = dict(encoder=enc,
               head=dict(fpn=dict(in_channels_list=chans, out_channels=256),
                         fs_relation=dict(scene_embedding_channels=chans[-1], in_channels_list=(256,) * 4, out_channels=256,
                                          scale_aware_proj=True),
                         fpn_decoder=dict(in_channels=256, out_channels=decoder_channels, in_feat_output_strides=(4, 8, 16, 32),
                                          out_feat_output_stride=4,
                                          classifier_config=dict(scale_factor=4.0, num_classes=num_classes, kernel_size=1))))
    return RefFarSeg(cfg)
