"""CPU-side tests (no GPU): the C-ABI library loads and exports everything include/evb200.h declares, the plugin
module mirrors the reference's state_dict / registry contract, the product path refuses to run without CUDA, and the
multi-rank loss rule used by the engine (global Dice statistics, Dice gradient x world, gradient all-reduce AVG)
reproduces the reference's DDP semantics on 2 gloo ranks."""
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import ctypes
    from ever_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, 'include', 'evb200.h')).read()
    names = sorted(set(re.findall(r'\b(evb_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.evb_version() >= 100
    # and nothing exported that the header does not declare
    out = subprocess.check_output(['nm', '-D', '--defined-only', path], text=True)
    exported = sorted(set(re.findall(r' T (evb_[a-z0-9_]+)$', out, flags=re.M)))
    assert set(exported) <= set(names), set(exported) - set(names)


def test_sass_has_tcgen05_and_tma():
    """The built library really contains Blackwell tensor-core / TMA instructions (UTC*MMA, UTMALDG, LDTM)."""
    from ever_b200 import build
    path = build.build()
    cuobjdump = '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', path], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in sass or 'UTCMMA' in sass
    assert 'UTMALDG' in sass
    assert 'LDTM' in sass


@pytest.mark.parametrize('resnet,k,dec,opts', [('resnet18', 5, 128, {}), ('resnet50', 15, 256, {}), ('resnet101', 7, 256, {}),
                                               ('resnet18', 5, 128, dict(in_channels=8, scale_aware_proj=False)),
                                               ('resnet50_v1c', 5, 128, {}),
                                               ('resnext50_32x4d', 5, 128, {}), ('resnext101_32x8d', 5, 128, {}),
                                               # FSRelationV2 (per-level and shared), keys as ever/module/fs_relation.py:76-139
                                               ('resnet18', 5, 128, dict(fs_version=2)),
                                               ('resnet18', 5, 128, dict(fs_version=2, scale_aware_proj=False))])
def test_state_dict_contract(resnet, k, dec, opts):
    from ever_b200.module import FarSegB200
    from oracle.farseg_oracle import FarSegOracle
    m = FarSegB200(dict(encoder=dict(resnet_type=resnet, in_channels=opts.get('in_channels', 3), with_cp=(True, True, False, False)),
                        head=dict(fs_relation=dict(scale_aware_proj=opts.get('scale_aware_proj', True),
                                                   version=opts.get('fs_version', 1)),
                                  fpn_decoder=dict(out_channels=dec, classifier_config=dict(num_classes=k)))))
    o = FarSegOracle(resnet, k, dec, **opts)
    a = [(n, tuple(v.shape), v.dtype) for n, v in m.state_dict().items()]
    b = [(n, tuple(v.shape), v.dtype) for n, v in o.state_dict().items()]
    assert a == b
    assert [n for n, _ in m.named_parameters()] == [n for n, _ in o.named_parameters()]
    m.load_state_dict(o.state_dict(), strict=True)


def test_registry_and_config_merge():
    from ever_b200._ever_api import MODEL, ERModule
    from ever_b200.module import FarSegB200
    assert MODEL['FarSegB200'] is FarSegB200 and MODEL['FarSeg'] is FarSegB200
    assert issubclass(FarSegB200, ERModule)
    m = FarSegB200(dict(encoder=dict(resnet_type='resnet18'),
                        head=dict(fpn_decoder=dict(classifier_config=dict(num_classes=5)))))
    # recursive merge keeps defaults the user did not override (ever/core/config.py:76-89)
    assert m.config.head.fpn_decoder.out_channels == 256
    assert m.config.head.fpn_decoder.classifier_config.scale_factor == 4.0
    assert tuple(m.config.head.fpn.in_channels_list) == (64, 128, 256, 512)
    assert 'GLOBAL' in m.config


def test_no_cpu_fallback():
    from ever_b200.module import FarSegB200
    m = FarSegB200(dict(encoder=dict(resnet_type='resnet18'),
                        head=dict(fpn_decoder=dict(classifier_config=dict(num_classes=5))))).train()
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(torch.zeros(1, 3, 64, 64), dict(cls=torch.zeros(1, 64, 64, dtype=torch.long)))


@pytest.mark.skipif(not os.path.isdir('/root/reference/ever'), reason='reference tree only exists in the build container')
def test_make_model_through_real_ever():
    """With the real reference on the path the plugin registers into ever.registry.MODEL and
    ever.core.builder.make_model builds it from a reference-style config dict."""
    code = r'''
import sys
sys.path.insert(0, "/root/reference"); sys.path.insert(0, "%s/tests/golden/_stubs"); sys.path.insert(0, "%s")
import ever as er
from ever.core.builder import make_model
import ever_b200.module as mod
from ever_b200._ever_api import HAVE_EVER
assert HAVE_EVER and issubclass(mod.FarSegB200, er.ERModule)
cfg = dict(type="FarSegB200", params=dict(encoder=dict(resnet_type="resnet18"),
           head=dict(fpn_decoder=dict(out_channels=128, classifier_config=dict(num_classes=5)))))
m = make_model(cfg)
assert isinstance(m, mod.FarSegB200)
assert len(m.custom_param_groups()) == 1
assert "en.resnet.layer1.0.conv1.weight" in m.state_dict()
print("OK")
''' % (ROOT, ROOT)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert 'OK' in out.stdout, out.stderr[-2000:]


# ---------------------------------------------------------------------------------------------- 2-rank gloo
def _engine_loss_rule(logit, labels, k, world, dice_all_reduce):
    """torch restatement of evb_loss_stats / evb_loss_finalize / evb_loss_grad (head_loss.cu): returns the per-rank
    logit gradient the engine forms BEFORE the gradient all-reduce (AVG)."""
    flat = logit.permute(0, 2, 3, 1).reshape(-1, k).float()
    t = labels.reshape(-1)
    valid = t != 255
    p = flat.softmax(dim=1)
    onehot = F.one_hot(t.clamp(max=k - 1), k).float() * valid[:, None]
    pv = p * valid[:, None]
    inter, sp, sy = (pv * onehot).sum(0), pv.sum(0), onehot.sum(0)
    stats = torch.cat([inter, sp, sy])
    stats = dice_all_reduce(stats)
    inter, sp, sy = stats[:k], stats[k:2 * k], stats[2 * k:]
    z = sp + sy + 1.0
    a = world * 2.0 / (k * z)
    b = world * (2 * inter + 1.0) / (k * z * z)
    g = b[None, :] - onehot * a[None, :]
    dot = (p * g).sum(1, keepdim=True)
    d = (p - onehot) / valid.sum() + p * (g - dot)
    d = d * valid[:, None]
    return d.reshape(logit.shape[0], logit.shape[2], logit.shape[3], k).permute(0, 3, 1, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.distributed.nn import all_reduce as ar
    from oracle.farseg_oracle import dice_loss_oracle
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    k = 4
    conv = torch.nn.Conv2d(8, k, 1)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 8, 16, 16, generator=g)
    y = torch.randint(0, k, (2, 16, 16), generator=g)
    y[torch.rand(2, 16, 16, generator=g) < (0.05 + 0.2 * rank)] = 255   # different ignore counts per rank
    # reference semantics: DDP (grad mean over ranks) + differentiable all-reduce inside Dice (loss.py:20-23,46-48)
    logit = conv(x)
    loss = F.cross_entropy(logit, y, ignore_index=255) + dice_loss_oracle(logit, y, all_reduce=lambda v: ar(v))
    loss.backward()
    ref = [p.grad.clone() for p in conv.parameters()]
    for r in ref:
        dist.all_reduce(r)
        r /= world
    # engine rule
    conv.zero_grad()
    logit = conv(x)

    def red(v):
        v = v.clone()
        dist.all_reduce(v)
        return v
    d = _engine_loss_rule(logit.detach(), y, k, world, red)
    logit.backward(d)
    mine = [p.grad.clone() for p in conv.parameters()]
    for m_ in mine:
        dist.all_reduce(m_)
        m_ /= world
    err = max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(mine, ref))
    if rank == 0:
        q.put(err)
    dist.destroy_process_group()


def test_two_rank_loss_rule_matches_ddp_reference():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-5


def test_step_loop_lr_timing_matches_reference_launcher():
    """The reference sets the LR for the NEXT iteration from the not-yet-incremented global step
    (ever/core/launcher.py:224-237): iteration 1 at base_lr, iteration k >= 2 at schedule(k - 2).  SURVEY.md a15 records the
    Launcher log for poly(base 0.007, power 0.9, max_iters 3): lr after steps 1, 2, 3 = 0.007, 0.00486, 0.002604."""
    from ever_b200.trainer import StepLoop, poly_lr

    class _Eng:
        def set_distributed(self, *a):
            pass

    class _Cfg:
        cuda_graph = False

    class _Model:
        config = _Cfg()

        def _engine(self):
            return _Eng()
    loop = StepLoop(_Model(), poly_lr(0.007, 0.9, 3), base_lr=0.007)
    logged, used = [], []
    for _ in range(3):
        used.append(loop.lr)
        loop._update_lr()
        loop.global_step += 1
        logged.append(loop.lr)
    assert [round(v, 6) for v in logged] == [0.007, 0.00486, 0.002604]
    assert [round(v, 6) for v in used] == [0.007, 0.007, 0.00486]


def test_c_abi_rejects_bad_arguments_without_touching_the_device():
    """Every entry point validates its arguments first and reports EVB_ERR_ARG (1) as a status code -- nothing throws across
    the C boundary and no kernel is launched (this runs on a box without a GPU)."""
    import ctypes
    from ever_b200._lib import lib
    L = lib()
    c_int, c_ll, c_f = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    null = ctypes.c_void_p(0)
    ERR_ARG = 1
    # convolution: only 1x1 / 3x3, stride 1 / 2
    assert L.evb_conv2d_fwd(null, c_int(1), c_int(32), c_int(32), c_int(64), null, c_int(64), c_int(5), c_int(1), null,
                            c_int(64), null, null, c_int(0), c_int(0), null) == ERR_ARG
    assert L.evb_conv2d_dgrad(null, c_int(1), c_int(16), c_int(16), c_int(64), null, c_int(64), c_int(3), c_int(3), null,
                              c_int(32), c_int(32), c_int(64), c_int(0), c_int(0), null) == ERR_ARG
    # wgrad: channel counts must be multiples of 64
    assert L.evb_conv2d_wgrad(null, c_int(1), c_int(32), c_int(32), c_int(48), null, c_int(64), c_int(3), c_int(1), null,
                              c_int(0), null, c_ll(0), c_int(0), c_int(0), null) == ERR_ARG
    # epilogue statistics need their output buffers
    assert L.evb_conv2d_fwd_stats(null, c_int(1), c_int(32), c_int(32), c_int(64), null, c_int(64), c_int(3), c_int(1), null,
                                  c_int(64), null, null, null) == ERR_ARG
    # BatchNorm: C % 8, mask mode, partial-column count
    assert L.evb_bn_apply(null, null, null, null, null, c_ll(16), c_int(12), c_int(1), null) == ERR_ARG
    assert L.evb_bn_apply_mask(null, null, null, null, null, null, c_ll(16), c_int(64), null) == ERR_ARG   # needs res + mask
    assert L.evb_bn_bwd(null, null, null, null, null, null, null, c_int(4), c_int(0), null, null, c_int(0), null, null,
                        c_int(0), c_ll(16), c_int(64), null, null) == ERR_ARG
    assert L.evb_bn_finalize(null, c_int(0), c_ll(16), c_int(64), null, null, null, null, c_f(0.1), c_f(1e-5), null, null, null,
                             null, null) == ERR_ARG
    # im2col window, gather element size, canvas bounding box, loss class count
    assert L.evb_im2col_nchw(null, null, c_int(1), c_int(3), c_int(32), c_int(32), c_int(20), c_int(3), c_int(2), c_int(1),
                             null) == ERR_ARG
    assert L.evb_pixel_gather(null, c_int(1), c_int(8), c_int(8), c_int(3), c_int(3), c_ll(0), null, null, c_int(1), c_int(8),
                              c_int(8), null) == ERR_ARG
    assert L.evb_canvas_accumulate(null, c_int(1), c_int(5), c_int(8), c_int(8), null, null, null, c_int(1), c_int(16),
                                   c_int(16), c_int(0), c_int(32), c_int(0), c_int(16), null) == ERR_ARG
    assert L.evb_loss_stats(null, null, c_ll(16), c_int(17), c_int(16), c_int(255), null, null, null) == ERR_ARG
    # tuning switches validate their ranges
    assert L.evb_set_bn_reduce_blocks(c_int(9)) == ERR_ARG
    assert L.evb_version() >= 101


def test_scale_transform_output_size_matches_interpolate():
    """Scale (ever/magic/transform/segm.py:71-88): the size our TTA path resizes to == the shape F.interpolate produces, for
    the factors the reference's own unit test sweeps (segm.py:100-105)"""
    import numpy as np
    import torch.nn.functional as F
    from ever_b200.infer import Scale, _scaled_size
    x = torch.zeros(1, 1, 90, 130)
    factors = [float(f) for f in np.linspace(0.25, 2.0, num=8)] + [0.49, 1.3, (0.5, 1.5)]
    for f in factors:
        want = F.interpolate(x, scale_factor=f, mode='bilinear', align_corners=True).shape[2:]
        assert _scaled_size(Scale(scale_factor=f), 90, 130) == tuple(want), f
    assert _scaled_size(Scale(size=(894, 896)), 90, 130) == (894, 896)
    assert _scaled_size(Scale(size=64), 90, 130) == (64, 64)
    with pytest.raises(ValueError):
        Scale()
    with pytest.raises(ValueError):
        Scale(size=4, scale_factor=2.0)


@pytest.mark.parametrize('splits', [(3, 2, 1), (3,), ()])
def test_gradient_buckets_tile_the_arena_in_backward_order(splits):
    """the per-stage gradient buckets of the multi-GPU step: contiguous arena tails starting at the first
    parameter of an encoder stage, listed in the order backward completes them, covering every gradient exactly once;
    _allreduce_bucket maps a finished tape segment to its bucket (host logic only: the engine object is a shell here)"""
    from ever_b200.engine import FarSegEngine, _ceil
    from ever_b200.module import FarSegB200
    m = FarSegB200(dict(encoder=dict(resnet_type='resnet18'),
                        head=dict(fpn_decoder=dict(out_channels=128, classifier_config=dict(num_classes=5)))))
    e = FarSegEngine.__new__(FarSegEngine)
    e.m, e.ar_split_stages, e._buckets, e.side = m, splits, None, None
    off, e._slots = 0, []
    for p in m.parameters():
        e._slots.append((off, p.numel()))
        off += _ceil(p.numel(), 4)
    e.flat_g = torch.zeros(off)
    b = e._bucket_ranges()
    assert len(b) == len(splits) + 1 and b[0][1] == off and b[-1][0] == 0
    assert all(b[i][0] == b[i + 1][1] for i in range(len(b) - 1))          # adjacent, descending: backward order
    names = [n for n, _ in m.named_parameters()]
    starts = {e._slots[names.index('en.resnet.layer%d.0.conv1.weight' % (s + 1))][0] for s in splits}
    assert {lo for lo, _ in b[:-1]} == starts
    # a finished tape segment -> its bucket: tape split points are recorded in forward order (stage 1 first)
    e.tape = list(range(100))
    e._tape_splits = [10 * (i + 1) for i in range(len(splits))]
    issued = []
    e._issue_allreduce = lambda lo, hi: issued.append((lo, hi))
    for lo, _ in e._segments():
        e._allreduce_bucket(lo)
    assert issued == b


@pytest.mark.skipif(not os.path.isdir('/root/reference/ever'), reason='reference tree only exists in the build container')
def test_backbone_names_equal_the_reference_registry():
    """every `resnet_type` the reference's ResNetEncoder accepts (registry.MODEL.register calls of ever/module/resnet.py:26-34)
    is a backbone of the plugin, and nothing else is"""
    import re
    from ever_b200.module import RESNETS
    src = open('/root/reference/ever/module/resnet.py').read()
    names = set(re.findall(r"registry\.MODEL\.register\('(res\w+)'", src))
    assert names and names == set(RESNETS)


def test_accounting_table_names_are_c_abi_entry_points():
    """every rule of the step-accounting proxy (ever_b200/acct.py, bench.py's hbm_bytes_per_step / flops_per_step) is keyed by
    a function include/evb200.h declares: a renamed entry point cannot silently drop out of the accounting"""
    import re
    from ever_b200.acct import TABLE
    hdr = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'evb200.h')).read()
    declared = set(re.findall(r'\b(evb_\w+)\s*\(', hdr))
    assert set(TABLE) <= declared, sorted(set(TABLE) - declared)
