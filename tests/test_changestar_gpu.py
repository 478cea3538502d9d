"""ChangeStar (FarSeg features + ChangeMixin) training step and eval against oracle/changestar_oracle.py (parity of this
row is UNPINNED with respect to upstream: ChangeStar is not in the reference tree -- see the oracle's header)."""
import copy
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _data(n, h, w, k):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, 6, h, w, generator=g)
    cls = torch.randint(0, max(k, 2), (n, h, w), generator=g)
    chg = torch.randint(0, 2, (n, h, w), generator=g)
    cls[torch.rand(n, h, w, generator=g) < 0.05] = 255
    chg[torch.rand(n, h, w, generator=g) < 0.05] = 255
    return x.cuda(), dict(cls=cls.cuda(), change=chg.cuda())


@pytest.mark.parametrize('case', [('resnet18', 1, 128, 2, 128, 128), ('resnet50', 5, 256, 2, 128, 128)])
def test_changestar_train_step(case):
    from ever_b200.module import ChangeStarB200
    from oracle.changestar_oracle import ChangeStarOracle
    from oracle.farseg_oracle import deterministic_fill
    resnet, k, dec, n, h, w = case
    ora = deterministic_fill(ChangeStarOracle(resnet, k, dec), 0)
    mine = ChangeStarB200(dict(encoder=dict(resnet_type=resnet),
                               head=dict(fpn_decoder=dict(out_channels=dec, classifier_config=dict(num_classes=k)))))
    mine.load_state_dict(ora.state_dict(), strict=True)
    x, y = _data(n, h, w, k)
    torch.backends.cudnn.allow_tf32 = False
    ora = ora.cuda().train()
    ora32 = copy.deepcopy(ora)
    mine = mine.cuda().train()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        lb = ora(x, y)
    sum(lb.values()).backward()
    l32 = ora32(x, y)
    sum(l32.values()).backward()
    out = mine(x, y)
    mine.backward(out, None, None)
    torch.cuda.synchronize()
    got = {kk: float(v) for kk, v in out.items()}
    ref = {kk: float(v) for kk, v in lb.items()}
    assert set(got) == set(ref)
    for kk in ref:
        assert abs(got[kk] - ref[kk]) <= 1e-2 * abs(ref[kk]), (kk, got, ref)
    pm, pb, p32 = dict(mine.named_parameters()), dict(ora.named_parameters()), dict(ora32.named_parameters())
    gmax = max(float(p_.grad.norm()) for p_ in pb.values())
    rep, bad = {}, {}
    for name in pm:
        if float(pb[name].grad.norm()) < 1e-6 * gmax:
            continue
        e, noise = _rel(pm[name].grad, pb[name].grad), _rel(pb[name].grad, p32[name].grad)
        rep[name] = (e, noise)
        if e > max(4 * noise, 5e-2) and not (name.endswith('0.bias') and 'encoders' in name):
            bad[name] = (e, noise)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(dict(losses=got, ref=ref, grads={k_: v for k_, v in rep.items() if 'changemixin' in k_}),
              open('gpurun_out/parity_changestar_%s.json' % resnet, 'w'), indent=1)
    assert not bad, list(bad.items())[:8]
    # BN running statistics of the (zero-padded) ChangeMixin BNs are written back to the module buffers
    for name, buf in mine.named_buffers():
        if 'changemixin' in name and 'running' in name:
            assert _rel(buf, dict(ora.named_buffers())[name]) < 2e-2, name


def test_changestar_eval():
    from ever_b200.module import ChangeStarB200
    from oracle.changestar_oracle import ChangeStarOracle
    from oracle.farseg_oracle import deterministic_fill
    resnet, k, dec, n, h, w = 'resnet18', 5, 128, 2, 128, 128
    ora = deterministic_fill(ChangeStarOracle(resnet, k, dec), 0).cuda().train()
    x, _ = _data(n, h, w, k)
    for m_ in ora.modules():   # calibrate running statistics to the batch statistics
        if isinstance(m_, torch.nn.BatchNorm2d):
            m_.momentum = 1.0
    with torch.no_grad():
        ora(x, dict(cls=torch.zeros(n, h, w, dtype=torch.long, device='cuda'), change=torch.zeros(n, h, w, dtype=torch.long, device='cuda')))
    mine = ChangeStarB200(dict(encoder=dict(resnet_type=resnet),
                               head=dict(fpn_decoder=dict(out_channels=dec, classifier_config=dict(num_classes=k)))))
    mine.load_state_dict(ora.state_dict(), strict=True)
    ora.eval()
    mine = mine.cuda().eval()
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        ref = ora(x)
    got = mine(x)
    torch.cuda.synchronize()
    assert _rel(got['seg'], ref['seg'].float()) < 3e-2
    assert _rel(got['change'], ref['change'].float()) < 3e-2
    agree = float((got['change_mask'].bool() == (ref['change'][:, 0] > 0.5)).float().mean())
    assert agree > 0.97, agree
