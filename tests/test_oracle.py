"""The oracle (oracle/farseg_oracle.py) is pinned two ways:
  * against the committed fixtures produced from the REAL reference (tests/golden/make_golden.py);
  * when /root/reference is present (build container only), bit-exact against the reference itself.
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from oracle.farseg_oracle import (FarSegOracle, bce_loss_oracle, deterministic_fill, dice_loss_oracle,
                                  synthetic_batch)

CASES = ['r18_k5_2x64', 'r50_k15_1x64', 'r18_k1_2x64', 'r18_k5_c8_shared_2x64', 'r50v1c_k5_1x64', 'r18_k5_v2_2x64', 'rx50_k5_1x64']
DROP_SEED = 20240   # tests/golden/make_golden.py


def _run_oracle(case):
    resnet, k, n, h, w, dec = case[:6]
    opts = case[6] if len(case) > 6 else {}
    m = deterministic_fill(FarSegOracle(resnet, k, dec, **opts), 0)
    x, y = synthetic_batch(n, h, w, max(k, 2), in_channels=opts.get('in_channels', 3))
    m.train()
    if k == 1:
        logit = m.logits(x)
        losses = dict(bce_loss=bce_loss_oracle(logit, y), dice_loss=dice_loss_oracle(logit, y))
    else:
        torch.manual_seed(DROP_SEED)
        logit = m.logits(x)
        losses = dict(ce_loss=F.cross_entropy(logit, y.long(), ignore_index=255), dice_loss=dice_loss_oracle(logit, y))
    sum(losses.values()).backward()
    return m, x, y, logit, losses


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_golden(name, golden_dir):
    g = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    torch.set_num_threads(8)
    m, x, y, logit, losses = _run_oracle(g['case'])
    for k, v in g['losses'].items():
        assert abs(float(losses[k]) - v) <= 1e-6 * max(1, abs(v)), (k, float(losses[k]), v)
    torch.testing.assert_close(logit.detach()[:, :, ::4, ::4], g['logit_slice'], rtol=1e-5, atol=1e-6)
    named = dict(m.named_parameters())
    assert set(named) == set(g['grad_norm'])
    for k, v in g['grad_norm'].items():
        got = float(named[k].grad.double().norm())
        assert abs(got - v) <= 1e-4 * max(v, 1e-6), (k, got, v)
    for k, v in g['stem_bn_running'].items():
        torch.testing.assert_close(m.state_dict()[k], v, rtol=1e-5, atol=1e-6)
    m.eval()
    with torch.no_grad():
        prob = m(x)
    mask = prob.argmax(dim=1).to(torch.uint8) if g['case'][1] > 1 else (prob > 0.5).to(torch.uint8)
    # fp32 CPU oracle vs fp32 CPU reference: same ops in the same order -> identical masks
    assert torch.equal(mask, g['eval_mask'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/ever'), reason='reference tree only exists in the build container')
@pytest.mark.parametrize('name', ['r18_k5_2x64', 'r18_k5_c8_shared_2x64', 'r50v1c_k5_1x64', 'r18_k5_v2_2x64', 'rx50_k5_1x64'])
def test_oracle_bit_exact_vs_reference(name, golden_dir):
    for p_ in ('/root/reference', os.path.join(golden_dir, '_stubs'), golden_dir):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    import make_golden as mg
    resnet, k, n, h, w, dec = mg.CASES[name][:6]
    opts = mg.CASES[name][6] if len(mg.CASES[name]) > 6 else {}
    ref = deterministic_fill(mg.RefFarSeg(mg.ref_config(resnet, k, dec, **opts)), 0)
    ora = deterministic_fill(FarSegOracle(resnet, k, dec, **opts), 0)
    assert list(ref.state_dict().keys()) == list(ora.state_dict().keys())
    x, y = synthetic_batch(n, h, w, k, in_channels=opts.get('in_channels', 3))
    ref.train(), ora.train()
    torch.set_num_threads(8)
    torch.manual_seed(DROP_SEED)
    lr, _ = ref(x, dict(cls=y))
    torch.manual_seed(DROP_SEED)
    lo = ora(x, dict(cls=y))
    # ResNet-18 cases are bit-reproducible on the CPU; the ResNet-50 case is not even reference-vs-reference (a fresh copy
    # of the real reference differs from itself by ~3e-6 in 165 gradients: threaded MKL-DNN reductions), so it is compared
    # at 5e-5 of each tensor's max instead
    exact = 'r18' in name
    for kk in lr:
        if exact:
            assert torch.equal(lr[kk], lo[kk]), kk
        else:
            assert abs(float(lr[kk]) - float(lo[kk])) <= 1e-6 * max(1.0, abs(float(lr[kk]))), kk
    sum(lr.values()).backward()
    sum(lo.values()).backward()
    for (ka, pa), (kb, pb) in zip(ref.named_parameters(), ora.named_parameters()):
        assert ka == kb
        if exact:
            assert torch.equal(pa.grad, pb.grad), ka
        else:
            assert float((pa.grad - pb.grad).abs().max()) <= 5e-5 * float(pa.grad.abs().max()) + 1e-12, ka
