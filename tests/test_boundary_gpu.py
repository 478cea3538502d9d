"""The drop-in boundary under the reference's OWN training stack: the unmodified ``ever`` package from
baseline/_ref drives the plugin -- ``Launcher.compute_loss_gradient`` (losses / forward_times, ever/core/launcher.py:193-200),
``ERModule.backward`` (sum + ``.backward()``, ever/interface/module.py:76-81), ``apply_gradients`` (clip + stock optimizer),
and the ``DistributedDataParallel`` wrapper ``THDDPTrainer.make_model`` always applies (ever/trainer/th_ddp_trainer.py:25-30).
"""
import os
import tempfile

import pytest
import torch

from _helpers import rel_l2

pytestmark = pytest.mark.gpu


def _pair(resnet='resnet18', k=5, dec=128):
    from ever_b200.module import FarSegB200
    from oracle.farseg_oracle import FarSegOracle, deterministic_fill
    ora = deterministic_fill(FarSegOracle(resnet, k, dec), 0)
    mine = FarSegB200(dict(encoder=dict(resnet_type=resnet),
                           head=dict(fpn_decoder=dict(out_channels=dec, classifier_config=dict(num_classes=k)))))
    mine.load_state_dict(ora.state_dict(), strict=True)
    return ora, mine


def _batches(nb, n, h, w, k):
    from oracle.farseg_oracle import synthetic_batch
    out = []
    for i in range(nb):
        x, y = synthetic_batch(n, h, w, k, seed_offset=i)
        out.append((x, dict(cls=y)))
    return out


def test_autograd_path_scales_and_accumulates():
    """What Launcher does with forward_times = 2: two micro-batches, losses / 2, ``backward(loss_dict)`` each, gradients
    accumulate in p.grad.  Result == 0.5 * (g1 + g2) of two native unit-weight steps; a different factor per loss is applied
    per loss (checked against native runs with only that loss' gradient)."""
    _, mine = _pair()
    mine = mine.cuda().train()
    (x1, y1), (x2, y2) = [(x.cuda(), {k: v.cuda() for k, v in y.items()}) for x, y in _batches(2, 2, 128, 128, 5)]
    state0 = {kk: v.clone() for kk, v in mine.state_dict().items()}

    def native(x, y):
        mine.load_state_dict(state0)
        out = mine(x, y)
        mine.backward(out, None, None)
        torch.cuda.synchronize()
        return mine.engine.flat_g.clone()
    g1, g2 = native(x1, y1), native(x2, y2)
    mine.load_state_dict(state0)
    mine.zero_grad(set_to_none=True)
    for x, y in ((x1, y1), (x2, y2)):
        out = mine(x, y)
        assert all(v.requires_grad for v in out.values())
        mine.backward({k: v / 2 for k, v in out.items()}, False, None)
    torch.cuda.synchronize()
    eng = mine.engine
    got = torch.zeros_like(eng.flat_g)
    for p, (o, n) in zip(eng.params, eng._slots):
        assert p.grad is not None and p.grad.data_ptr() != eng.flat_g[o:o + n].data_ptr()   # ordinary autograd-owned grads
        got[o:o + n] = p.grad.flatten()
    want = 0.5 * (g1 + g2)
    assert rel_l2(got, want) < 5e-3, rel_l2(got, want)
    # per-loss factors: total = 3 * ce + 0 * dice  ->  gradient == 3 x (native step with the Dice weight switched off)
    mine.load_state_dict(state0)
    mine.zero_grad(set_to_none=True)
    out = mine(x1, y1)
    (3.0 * out['ce_loss'] + 0.0 * out['dice_loss']).backward()
    torch.cuda.synchronize()
    got = torch.cat([p.grad.flatten() for p in eng.params])
    eng.dice_w = 0.0
    mine.load_state_dict(state0)
    o2 = mine(x1, y1)
    mine.backward(o2, None, None)
    torch.cuda.synchronize()
    eng.dice_w = 1.0
    want = 3.0 * torch.cat([eng.flat_g[o:o + n] for (o, n) in eng._slots])
    # a factor that is not a power of two re-rounds the bf16 logit gradient (2^-9 per element), which the BatchNorm
    # backward passes amplify to ~1e-2 on the whole arena; the power-of-two case above is exact to 5e-3
    assert rel_l2(got, want) < 3e-2, rel_l2(got, want)


@pytest.mark.parametrize('forward_times', [1, 2])
def test_real_launcher_trains_plugin_like_the_reference(forward_times):
    """``Launcher.train_iters`` of the unmodified reference (baseline/_ref) for 5 iterations: once with the reference's own
    modules (glue model of SURVEY.md Appendix E), once with FarSegB200, same initial weights, same batches, bf16 mixed
    precision, the reference's SGD / poly-LR factories and grad clipping.  The per-forward losses follow each other."""
    from _launcher_harness import make_reference_farseg, reference_available, run_launcher
    if not reference_available():
        pytest.skip('baseline/_ref not installed')
    ref = make_reference_farseg('resnet18', 5, 128)
    ora, mine = _pair()
    ref.load_state_dict(ora.state_dict(), strict=True)    # same keys: the state_dict contract
    batches = _batches(2, 2, 128, 128, 5)
    iters = 5
    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2:
        seen_ref, last_ref = run_launcher(ref.cuda(), batches, iters, d1, forward_times=forward_times)
        seen_mine, last_mine = run_launcher(mine.cuda(), batches, iters, d2, forward_times=forward_times)
        assert os.path.exists(os.path.join(d2, 'checkpoint-%d.pth' % iters))   # SaveCheckpointCallback ran on the plugin
    assert len(seen_ref) == len(seen_mine) == iters * forward_times
    print('ref ', [round(d['ce_loss'] + d['dice_loss'], 4) for d in seen_ref])
    print('mine', [round(d['ce_loss'] + d['dice_loss'], 4) for d in seen_mine])
    for a, b in zip(seen_mine, seen_ref):
        for kk in b:
            assert abs(a[kk] - b[kk]) <= 2e-2 * abs(b[kk]) + 1e-3, (kk, seen_mine, seen_ref)
    assert seen_mine[-1]['ce_loss'] < seen_mine[0]['ce_loss']
    # the global gradient norm of a bf16 step is conditioning-limited: same order of magnitude only
    assert 0.4 * last_ref['grad_norm'] <= last_mine['grad_norm'] <= 2.5 * last_ref['grad_norm'], (last_mine, last_ref)
    # the stock optimizer really updated the plugin's parameters (views into the engine's arena)
    assert rel_l2(mine.state_dict()['head.fpn_decoder.classifier.0.weight'].cpu(),
                  ora.state_dict()['head.fpn_decoder.classifier.0.weight']) > 1e-4


def test_ddp_wrapped_plugin_survives_the_stock_trainer():
    """THDDPTrainer.make_model wraps every model in DistributedDataParallel (th_ddp_trainer.py:25-30) and Launcher calls
    the wrapper's forward, then the UNWRAPPED model's backward: the DDP reducer must see every parameter's gradient hook
    fire each iteration (it raises on the 2nd forward otherwise).  world_size 1 NCCL group; 4 iterations; losses identical
    to the un-wrapped Launcher run (all-reduce over one rank is the identity)."""
    import torch.distributed as dist
    from _launcher_harness import reference_available, run_launcher
    if not reference_available():
        pytest.skip('baseline/_ref not installed')
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', str(29600 + os.getpid() % 2000))
    created = not dist.is_initialized()
    if created:
        dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    try:
        batches = _batches(2, 2, 128, 128, 5)
        _, plain = _pair()
        _, wrapped = _pair()
        with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2:
            seen_plain, _ = run_launcher(plain.cuda(), batches, 4, d1)
            seen_ddp, _ = run_launcher(wrapped.cuda(), batches, 4, d2,
                                       wrap=lambda m: torch.nn.parallel.DistributedDataParallel(m, device_ids=[0], output_device=0))
        assert seen_ddp == seen_plain, (seen_ddp, seen_plain)
    finally:
        if created:
            dist.destroy_process_group()
