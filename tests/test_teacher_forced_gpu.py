"""Teacher-forced whole-model parity (the north_star 1e-2 bf16 gate, made meaningful).

End-to-end gradients of a ~50-layer bf16 ReLU/BatchNorm network are chaotic: the reference's own bf16-autocast run
differs from its fp32 run by O(1) relative L2 on most tensors.  Here the composed model -- the
engine's real forward schedule and its real backward tape, every kernel, every fusion, every stream branch -- is checked
op by op on the REFERENCE'S OWN tensors: the oracle runs one bf16-autocast step with hooks recording the output and
output-gradient of every module; the engine then runs its step with ``engine.tf`` set, which compares each activation
(forward) and each completed activation gradient (backward) with the oracle's tensor of the same module path and
overwrites it before the next op consumes it.  Every y, dx, dW, d-gamma, d-beta of the network is therefore produced from
the reference's inputs by exactly the code the product runs, and must agree to <= 1e-2 relative L2.
"""
import json
import os
import re

import pytest
import torch
import torch.nn.functional as F

from _helpers import RefCapture, TeacherForcing, rel_l2

pytestmark = pytest.mark.gpu

GATE = 1e-2
# Conv2d(bias=True) -> training-mode BatchNorm2d of the FS-Relation encoders (ever/module/fs_relation.py:41-52)
ZERO_BIAS = re.compile(r'head\.fs_relation\.(content_encoders|feature_reencoders)\.\d+\.0\.bias$')
# (conv, BN, ReLU) module paths of the BatchNorm + ReLU in front of the max-pool
STEM_BN = dict(plain=('en.resnet.conv1', 'en.resnet.bn1', 'en.resnet.relu#0'),
               v1c=('en.resnet.stem.6', 'en.resnet.stem.7', 'en.resnet.stem.8'))


def _build(resnet, k, dec, **opts):
    from ever_b200.module import FarSegB200
    from oracle.farseg_oracle import FarSegOracle, deterministic_fill
    ora = deterministic_fill(FarSegOracle(resnet, k, dec, **opts), 0)
    mine = FarSegB200(dict(encoder=dict(resnet_type=resnet, in_channels=opts.get('in_channels', 3)),
                           head=dict(fs_relation=dict(scale_aware_proj=opts.get('scale_aware_proj', True),
                                                      version=opts.get('fs_version', 1)),
                                     fpn_decoder=dict(out_channels=dec, classifier_config=dict(
                                         num_classes=k, kernel_size=opts.get('classifier_kernel_size', 1),
                                         dropout_rate=opts.get('classifier_dropout', -1))))))
    mine.load_state_dict(ora.state_dict(), strict=True)
    return ora.cuda().train(), mine.cuda().train()


def bn_truth(cap, params, bn_name):
    """fp64 (d-gamma, d-beta) of the BatchNorm `bn_name` recomputed from the reference's captured tensors: x = output of the
    conv in front of it, g = gradient arriving at the BN output (= the gradient of the ReLU output behind it masked by
    output > 0; for the BN of a residual branch the block's last ReLU; a down-sample BN has no ReLU).  None if the layer
    does not follow one of those patterns."""
    head, idx = bn_name.rsplit('.', 1)
    if idx in ('bn1', 'bn2', 'bn3'):                      # ResNet blocks and the plain stem
        conv = head + '.conv' + idx[2]
        relu = head + '.relu#%d' % (int(idx[2]) - 1)
    elif idx.isdigit() and '.downsample' in head:         # Sequential(conv, BN)
        conv, relu = '%s.%d' % (head, int(idx) - 1), None
    elif idx.isdigit():                                   # Sequential(conv, BN, ReLU, ...): deep stem, decoder, FS-Relation
        conv, relu = '%s.%d' % (head, int(idx) - 1), '%s.%d' % (head, int(idx) + 1)
    else:
        return None
    if conv not in cap.fwd or (relu is not None and relu not in cap.bwd) or (relu is None and bn_name not in cap.bwd):
        return None
    x = cap.fwd[conv].double()
    g = cap.bwd[bn_name].double() if relu is None else cap.bwd[relu].double() * (cap.fwd[relu] > 0)
    dims = (0, 2, 3)
    mean = x.mean(dims, keepdim=True)
    xhat = (x - mean) / torch.sqrt(x.var(dims, unbiased=False, keepdim=True) + 1e-5)
    return (g * xhat).sum(dims), g.sum(dims)


def oracle_step_captured(ora, x, y, all_reduce=None):
    from oracle.farseg_oracle import bce_loss_oracle, dice_loss_oracle
    cap = RefCapture(ora)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logit = ora.logits(x)
        if logit.shape[1] == 1:
            losses = dict(bce_loss=bce_loss_oracle(logit, y), dice_loss=dice_loss_oracle(logit, y, all_reduce=all_reduce))
        else:
            losses = dict(ce_loss=F.cross_entropy(logit, y.long(), ignore_index=255),
                          dice_loss=dice_loss_oracle(logit, y, all_reduce=all_reduce))
    sum(losses.values()).backward()
    cap.remove()
    torch.cuda.synchronize()
    return cap, {k: float(v) for k, v in losses.items()}


def summarize(tf, grads, tag):
    rep = dict(fwd=tf.err['fwd'], bwd=tf.err['bwd'], param_grads=grads, missing=tf.missing)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(rep, open('gpurun_out/teacher_forced_%s.json' % tag, 'w'), indent=1)

    def stat(d):
        v = sorted(d.values())
        return dict(n=len(v), median=v[len(v) // 2], max=v[-1]) if v else dict(n=0)
    s = dict(case=tag, fwd=stat(tf.err['fwd']), bwd=stat(tf.err['bwd']), param_grads=stat(grads),
             worst_fwd=sorted(tf.err['fwd'].items(), key=lambda kv: -kv[1])[:3],
             worst_bwd=sorted(tf.err['bwd'].items(), key=lambda kv: -kv[1])[:3],
             worst_grad=sorted(grads.items(), key=lambda kv: -kv[1])[:3])
    print(json.dumps(s))
    return s


CASES = [
    # BASELINE configs[1] model at two full-size tiles (every layer shape class of C2; BN sees 2 x 16 x 16 = 512 samples at c5)
    ('resnet50', 15, 256, 2, 512, 512, {}),
    # configs[0] model (BasicBlock path) at its own tile size
    ('resnet18', 5, 128, 2, 256, 256, {}),
    # deep stem + shared scene encoder + binary head share the remaining op variants
    ('resnet50_v1c', 5, 128, 2, 256, 256, {}),
    ('resnet18', 1, 128, 2, 128, 128, dict(scale_aware_proj=False)),
    # 3x3 classifier (classifier_config.kernel_size, ever/module/fpn.py:172-181) and more than 16 classes (32- / 64-wide logit rows)
    ('resnet18', 21, 128, 2, 128, 128, dict(classifier_kernel_size=3)),
    ('resnet18', 40, 128, 2, 128, 128, {}),
    # FSRelationV2 (ever/module/fs_relation.py:76-163): GroupNorm scene encoder, concat + project conv, Dropout2d (both runs
    # draw their channel masks from the same torch seed, so the same channels are dropped)
    ('resnet18', 5, 128, 2, 128, 128, dict(fs_version=2)),
    ('resnet50', 7, 256, 2, 256, 256, dict(fs_version=2)),
    # nn.Dropout on the merged decoder features (classifier_config.dropout_rate, ever/module/fpn.py:175-176,190), same seed
    ('resnet18', 5, 128, 2, 128, 128, dict(classifier_dropout=0.3)),
    # ResNeXt: grouped 3x3 (32 groups) in every bottleneck (ever/module/_resnets.py:80-84,291-300), dense block-diagonal on the GPU
    ('resnext50_32x4d', 5, 128, 2, 256, 256, {}),
]
DROP_SEED = 4242


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s_k%d_%dx%d%s' % (c[0], c[1], c[3], c[4], '_v2' if c[6].get('fs_version') == 2 else '_drop' if c[6].get('classifier_dropout') else '') if isinstance(c, tuple) else None)
def test_teacher_forced_step(case):
    from oracle.farseg_oracle import synthetic_batch
    resnet, k, dec, n, h, w, opts = case
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ora, mine = _build(resnet, k, dec, **opts)
    x, y = synthetic_batch(n, h, w, max(k, 2))
    x, y = x.cuda(), y.cuda()
    torch.manual_seed(DROP_SEED)
    cap, loss_ref = oracle_step_captured(ora, x, y)
    eng = mine._engine()
    tf = TeacherForcing(cap, force=True)
    eng.tf = tf
    torch.manual_seed(DROP_SEED)
    out = mine(x, dict(cls=y))
    mine.backward(out, None, None)
    torch.cuda.synchronize()
    eng.tf = None
    pm, pr = dict(mine.named_parameters()), dict(ora.named_parameters())
    gmax = max(float(p.grad.norm()) for p in pr.values() if p.grad is not None)
    grads = {}
    for name, p in pm.items():
        g_ref = pr[name].grad
        if ZERO_BIAS.match(name):
            # conv bias directly before a training-mode BN: exactly zero in exact arithmetic (BN backward removes the
            # per-channel mean of its input gradient).  The reference holds bf16 rounding noise there (1e-3 .. 1e-6 of the
            # other gradients), the engine writes exact zeros
            assert float(g_ref.norm()) < 2e-3 * gmax, (name, float(g_ref.norm()), gmax)
            assert float(p.grad.norm()) == 0.0, name
            continue
        grads[name] = rel_l2(p.grad, g_ref)
    # BatchNorm d-gamma / d-beta are sums over all pixels; where the summands cancel (factor ~400 for the BN in front of the
    # max-pool, whose input gradient is the SPARSE max-pool gradient; 10-50 elsewhere) torch's own bf16 batch-norm backward is
    # 1-5 % away from the exact sums (tools/diag_tf.py: reference vs fp64 1.8e-2, engine vs fp64 2e-8).  A BN parameter
    # gradient that misses the 1e-2 gate against the reference is therefore held to the fp64 recomputation from the
    # reference's own captured tensors instead: engine within 1e-3 of fp64 AND the distance explained by the reference's error.
    bn_rep = {}
    for nm in [n_ for n_, e in grads.items() if e > GATE]:
        mod_name, leaf = nm.rsplit('.', 1)
        truth = bn_truth(cap, pr, mod_name) if leaf in ('weight', 'bias') else None
        if truth is None:
            continue
        t = truth[0 if leaf == 'weight' else 1]
        bn_rep[nm] = dict(engine_vs_fp64=rel_l2(pm[nm].grad, t), reference_vs_fp64=rel_l2(pr[nm].grad, t),
                          engine_vs_reference=grads[nm])
        assert bn_rep[nm]['engine_vs_fp64'] <= 1e-3, bn_rep
        assert bn_rep[nm]['reference_vs_fp64'] > 0.5 * grads[nm], bn_rep
        grads.pop(nm)
    # the max-pool BN is always reported (the extreme case)
    for leaf, idx in (('weight', 0), ('bias', 1)):
        nm = STEM_BN['v1c' if resnet.endswith('_v1c') else 'plain'][1] + '.' + leaf
        t = bn_truth(cap, pr, nm.rsplit('.', 1)[0])[idx]
        bn_rep.setdefault(nm, dict(engine_vs_fp64=rel_l2(pm[nm].grad, t), reference_vs_fp64=rel_l2(pr[nm].grad, t)))
        assert bn_rep[nm]['engine_vs_fp64'] <= 1e-3, bn_rep
    print(json.dumps(dict(bn_param_grads_vs_fp64=bn_rep)))
    tag = '%s_k%d_%dx%dx%d' % (resnet, k, n, h, w)
    s = summarize(tf, grads, tag)
    assert not tf.missing, tf.missing[:5]
    # every conv output / fused BN-ReLU(-residual) output / pooled, up-sampled, merged, relation tensor was compared
    n_convs = len([c for c in eng.convs if c.name])
    assert len(tf.err['fwd']) >= 2 * n_convs - 8 and len(tf.err['bwd']) >= 2 * n_convs - 10, s
    for kk, v in loss_ref.items():
        assert abs(float(out[kk]) - v) <= 2e-3 * abs(v), (kk, float(out[kk]), v)
    bad = [(kind, nm, e) for kind in ('fwd', 'bwd') for nm, e in tf.err[kind].items() if not e <= GATE]
    bad += [('grad', nm, e) for nm, e in grads.items() if not e <= GATE]
    assert not bad, bad[:12]


def test_compare_only_run_is_the_plain_step():
    """The hook in compare-only mode (force=False) does not alter the step: losses and gradients are bit-identical to a
    run without it (so the teacher-forced run exercises the same schedule as the product)."""
    from oracle.farseg_oracle import synthetic_batch
    ora, mine = _build('resnet18', 5, 128)
    x, y = synthetic_batch(2, 128, 128, 5)
    x, y = x.cuda(), y.cuda()
    cap, _ = oracle_step_captured(ora, x, y)
    state0 = {kk: v.clone() for kk, v in mine.state_dict().items()}
    out = mine(x, dict(cls=y))
    mine.backward(out, None, None)
    torch.cuda.synchronize()
    l0, g0 = {kk: float(v) for kk, v in out.items()}, mine.engine.flat_g.clone()
    mine.load_state_dict(state0)
    mine.engine.tf = TeacherForcing(cap, force=False)
    out = mine(x, dict(cls=y))
    mine.backward(out, None, None)
    torch.cuda.synchronize()
    assert {kk: float(v) for kk, v in out.items()} == l0
    assert torch.equal(mine.engine.flat_g, g0)
