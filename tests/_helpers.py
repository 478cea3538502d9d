"""Shared helpers of the GPU tests (tests/ is on sys.path under pytest's default import mode)."""
import torch


def eval_r18_model(k=5, calibrate_hw=(128, 128)):
    """eval-mode R18 FarSegB200 with BN running statistics calibrated on a random batch (well-scaled outputs)"""
    from ever_b200.module import FarSegB200
    from oracle.farseg_oracle import FarSegOracle, deterministic_fill, synthetic_batch
    ora = deterministic_fill(FarSegOracle('resnet18', k, 128), 0).cuda().train()
    for m_ in ora.modules():
        if isinstance(m_, torch.nn.BatchNorm2d):
            m_.momentum = 1.0
    x, _ = synthetic_batch(2, calibrate_hw[0], calibrate_hw[1], k)
    with torch.no_grad():
        ora.logits(x.cuda())
    mine = FarSegB200(dict(encoder=dict(resnet_type='resnet18'),
                           head=dict(fpn_decoder=dict(out_channels=128, classifier_config=dict(num_classes=k)))))
    mine.load_state_dict(ora.state_dict(), strict=True)
    return mine.cuda().eval()
