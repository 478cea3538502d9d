#!/usr/bin/env python
"""Benchmark of the hot path.  Default workload = BASELINE.json configs[1]: FarSeg-R50, 15 classes, 8 x 3 x 512 x 512
synthetic tiles per GPU; one step = forward + CE/Dice loss + backward (+ NCCL gradient all-reduce at N>1) + fused clip/SGD.

    python bench.py --gpus N --steps K --warmup W [--config c1|c2|c3|c4|c5]     # B200 engine, prints ONE JSON line
    python bench.py --impl reference --steps K --warmup W [--config ...]        # the reference's own PyTorch-CPU path

value  = whole-job tiles/s with inputs resident in HBM (CUDA-graph replay of the step, CUDA events, max over ranks)
e2e    = the same metric through the plugin call model(x, y) / model.backward() with pinned HOST inputs:
         H2D copies of the tile batch + labels and a D2H read of the losses inside the timed region.
Other keys: roofline (dominant tensor-core kernel), roofline_hbm (largest HBM-bound kernel class), cpu_baseline (the
reference on the host cores, N=1 only), gpu_incumbent (the reference's modules on the same GPU through torch/cuDNN),
hbm_bytes_per_step / flops_per_step (algorithmic, counted from the C-ABI calls of one step), clocks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs (SURVEY.md section 8 shorthand C1..C5).  per_gpu: tiles (or bitemporal pairs) per GPU;
# scaling 'weak' = per-GPU batch fixed, 'strong' = `total` units split over the ranks.
CONFIGS = {
    'c1': dict(model='FarSeg', resnet='resnet18', k=5, dec=256, hw=(256, 256), per_gpu=2, scaling='weak', cin=3,
               workload='FarSeg-R18 5-class, 2x3x256x256 synthetic tiles per GPU (BASELINE configs[0])', gflop_per_tile=65.6),
    'c2': dict(model='FarSeg', resnet='resnet50', k=15, dec=256, hw=(512, 512), per_gpu=8, scaling='weak', cin=3,
               workload='FarSeg-R50 (ResNet-50 + FPN + FS-Relation + asymmetric decoder, 256-ch), 15-class, 8x3x512x512 '
                        'synthetic tiles per GPU (BASELINE configs[1])', gflop_per_tile=343.1),
    'c2s': dict(model='FarSeg', resnet='resnet50', k=15, dec=256, hw=(512, 512), total=8, scaling='strong', cin=3,
                workload='FarSeg-R50 15-class, 8x3x512x512 synthetic tiles in TOTAL split over the GPUs (strong scaling of '
                         'BASELINE configs[1]: 1 tile per GPU at 8 GPUs)', gflop_per_tile=343.1),
    'c2x': dict(model='FarSeg', resnet='resnext50_32x4d', k=15, dec=256, hw=(512, 512), per_gpu=8, scaling='weak', cin=3,
                workload='FarSeg-ResNeXt50-32x4d 15-class, 8x3x512x512 synthetic tiles per GPU (configs[1] with the grouped-conv '
                         'backbone of ever/module/_resnets.py:291-300)', gflop_per_tile=None),
    'c3': dict(model='ChangeStar', resnet='resnet50', k=1, dec=256, hw=(512, 512), total=8, scaling='strong', cin=3,
               workload='ChangeStar (FarSeg-R50 + ChangeMixin), 8 bitemporal pairs of 3x512x512 sharded by batch over the '
                        'GPUs (BASELINE configs[2]); a pair counts as 2 tiles', gflop_per_tile=None),
    'c4': dict(model='FarSeg', resnet='resnet101', k=7, dec=256, hw=(1024, 1024), per_gpu=4, scaling='weak', cin=3,
               workload='FarSeg-R101 7-class, 4x3x1024x1024 synthetic tiles per GPU (BASELINE configs[3])',
               gflop_per_tile=1837.0),
    'c5': dict(model='FreeNet', resnet='freenet', k=9, dec=128, hw=(624, 352), per_gpu=1, scaling='weak', cin=200,
               workload='FreeNet (patch-free hyperspectral classification: conv3x3+GroupNorm+ReLU blocks with squeeze-excitation, '
                        'top-down fusion), one 200-band 610x340 cube zero-padded to 624x352 per GPU (BASELINE configs[4]), masked '
                        'CE on the labelled pixels', gflop_per_tile=None),
}
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu --set full capture
# summarised in profiles/ (None until captured)
NCU_TRAFFIC_BYTES = 68308480 + 23763968  # profiles/r02_ncu_full_igemm2_3x3_256_128.txt (dram__bytes_read + dram__bytes_write)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace('.', '', 1).isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == 'active'})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm), power_w_max=max(pw) if pw else None)


def model_config(cfg):
    if cfg['model'] == 'FreeNet':
        return dict(in_channels=cfg['cin'], num_classes=cfg['k'])
    return dict(encoder=dict(resnet_type=cfg['resnet'], in_channels=cfg['cin']),
                head=dict(fpn_decoder=dict(out_channels=cfg['dec'], classifier_config=dict(num_classes=cfg['k']))))


def units_per_gpu(cfg, world):
    if cfg['scaling'] == 'weak':
        return cfg['per_gpu']
    return max(1, cfg['total'] // world)


def synthetic(cfg, n, seed=0):
    """randn images (seed 1234), randint labels (seed 4321) with 5 % ignore (seed 99); bitemporal: two draws
    (1234 / 1235) stacked on the channel axis, binary change labels."""
    h, w = cfg['hw']
    kk = max(cfg['k'], 2)
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(n, cfg['cin'], h, w, generator=g)
    if cfg['model'] == 'FreeNet':   # labels 1..K on 5 % of the pixels (0 = unlabelled) + the training mask that selects them
        g = torch.Generator().manual_seed(4321 + seed)
        y = torch.randint(1, kk + 1, (n, h, w), generator=g)
        wm = (torch.rand(n, h, w, generator=g) < 0.05).float()
        return x, dict(cls=y * (wm > 0).long(), w=wm)
    g = torch.Generator().manual_seed(4321 + seed)
    y = torch.randint(0, kk, (n, h, w), generator=g)
    g = torch.Generator().manual_seed(99 + seed)
    y[torch.rand(n, h, w, generator=g) < 0.05] = 255
    if cfg['model'] == 'ChangeStar':
        g = torch.Generator().manual_seed(1235 + seed)
        x = torch.cat([x, torch.randn(n, cfg['cin'], h, w, generator=g)], dim=1)
        g = torch.Generator().manual_seed(777 + seed)
        return x, dict(cls=y, change=torch.randint(0, 2, (n, h, w), generator=g))
    return x, dict(cls=y)


def tiles_of(cfg, n):
    return n * (2 if cfg['model'] == 'ChangeStar' else 1)


def metric_name(cfg):
    h, w = cfg['hw']
    if cfg['model'] == 'FreeNet':
        return 'FreeNet %dx%dx%d cubes/s fwd+bwd' % (cfg['cin'], h, w)
    return '%s-%s %dx%d tiles/s fwd+bwd' % (cfg['model'], cfg['resnet'].replace('resnet', 'R'), h, w)


# ------------------------------------------------------------------------------------------------ reference models
def reference_model(cfg):
    """(model, kind): the UNMODIFIED reference's own modules composed into the glue model (kind 'reference', from
    baseline/_ref) when they exist for this config, else the oracle port (kind 'port')."""
    torch.manual_seed(0)
    if cfg['model'] == 'ChangeStar':
        from oracle.changestar_oracle import ChangeStarOracle
        return ChangeStarOracle(cfg['resnet'], cfg['k'], cfg['dec']), 'port'
    if cfg['model'] == 'FreeNet':   # not in the reference tree either: the restated network (its SEBlock is the in-tree one)
        from oracle.freenet_oracle import FreeNetOracle
        return FreeNetOracle(cfg['cin'], cfg['k']), 'port'
    try:
        from oracle.ref_glue import make_reference_farseg, reference_available
        if reference_available():
            return make_reference_farseg(cfg['resnet'], cfg['k'], cfg['dec'], in_channels=cfg['cin']), 'reference'
    except Exception as e:  # pragma: no cover
        sys.stderr.write('[bench] reference import failed (%s): falling back to the oracle port\n' % e)
    from oracle.farseg_oracle import FarSegOracle
    return FarSegOracle(cfg['resnet'], cfg['k'], cfg['dec'], in_channels=cfg['cin']), 'port'


def cpu_reference(cfg, n_units, iters, warmup):
    """The reference's PyTorch-CPU path (fp32; Launcher cannot autocast on CPU, ever/core/launcher.py:194) on all host
    cores.  Returns (tiles/s, seconds per step, cores, kind, steps actually timed)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m, kind = reference_model(cfg)
    m = m.train()
    x, y = synthetic(cfg, n_units)
    ts = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        m.zero_grad(set_to_none=True)
        losses = m(x, y)
        sum(v for k_, v in losses.items() if k_.endswith('loss')).backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    sec = sum(ts) / len(ts)
    return tiles_of(cfg, n_units) / sec, sec, cores, kind, len(ts)


CPU_SAMPLE = dict(c1=2, c2=4, c2s=4, c3=1, c4=1, c5=1)   # units per CPU step (bounded sample of the per-GPU batch)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    n_units = CPU_SAMPLE[args.config]
    # bounded: at most 5 timed steps and 1 warm-up (a C2 step of 4 tiles is ~1.3 s on 16 cores, a C4 tile ~10 s);
    # the line reports the counts that were actually run
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    val, sec, cores, kind, ran = cpu_reference(cfg, n_units, steps, warmup)
    sample = ('%d of the %d %s of the per-GPU batch per step, %d warm-up + %d timed fwd+loss+bwd steps, fp32, torch CPU, '
              '%d threads' % (n_units, units_per_gpu(cfg, 1), 'pairs' if cfg['model'] == 'ChangeStar' else 'tiles', warmup, ran,
                              cores))
    line = dict(metric=metric_name(cfg), value=val, unit='tiles/s', n_gpus=args.gpus, steps=ran, warmup=warmup,
                steps_requested=args.steps, warmup_requested=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling=cfg['scaling'], vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload=cfg['workload'] + ', fwd+loss+bwd', config=args.config, tiles_per_step=tiles_of(cfg, n_units)),
                cpu_baseline=dict(value=val, unit='tiles/s', cores=cores, kind=kind, sample=sample),
                e2e=dict(value=val, unit='tiles/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU incumbent
def gpu_incumbent(cfg, n_units, variants=('nchw', 'channels_last'), iters=15):
    """The reference's own GPU path on this box: its modules through torch/cuDNN, bf16 autocast, fwd + loss + bwd +
    clip_grad_norm_ + torch.optim.SGD, exactly what Launcher runs (NCHW eager), plus channels_last and torch.compile
    (ever/trainer/trainer.py:241-244)."""
    res = {}
    x, y = synthetic(cfg, n_units)
    x = x.cuda()
    y = {k_: v.cuda() for k_, v in y.items()}
    for fmt in variants:
        try:
            m, kind = reference_model(cfg)
            m = m.cuda().train()
            xx = x
            if fmt == 'channels_last':
                m = m.to(memory_format=torch.channels_last)
                xx = x.contiguous(memory_format=torch.channels_last)
            fwd = torch.compile(m) if fmt == 'compile' else m
            opt = torch.optim.SGD(m.parameters(), lr=0.007, momentum=0.9, weight_decay=1e-4)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    losses = fwd(xx, y)
                    loss = sum(v for k_, v in losses.items() if k_.endswith('loss'))
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 35.0)
                opt.step()
            for _ in range(5):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(iters):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            res[fmt] = dict(ms_per_step=ms, tiles_per_s=tiles_of(cfg, n_units) / ms * 1e3, kind=kind)
            del m, opt, fwd
        except Exception as e:  # pragma: no cover
            res[fmt] = dict(unavailable=str(e)[:200])
        torch.cuda.empty_cache()
    res['what'] = ('reference modules on the same GPU, torch %s + cuDNN, bf16 autocast, %d units per step, %d timed steps'
                   % (torch.__version__, n_units, iters))
    return res


# ------------------------------------------------------------------------------------------------ dominant-kernel roofline
def conv_roofline(pk):
    """3x3 256->256 conv on 8 x 128 x 128 (33.7 % of forward FLOPs), igemm2_kernel<256>, timed alone
    with CUDA events on the launch stream; operands rotate through > L2-size buffers."""
    from ever_b200 import ops
    n, h, w, c = 8, 128, 128, 256
    nbuf = 6  # 6 x (67 MB in + 67 MB out) > 126 MB L2
    xs = [torch.randn(n, h, w, c, device='cuda').bfloat16() for _ in range(nbuf)]
    ys = [torch.empty(n, h, w, c, device='cuda', dtype=torch.bfloat16) for _ in range(nbuf)]
    wt = torch.randn(c, c, 3, 3, device='cuda') * 0.02
    wf, _ = ops.pack_conv_weight_torch(wt)
    for i in range(nbuf):
        ops.conv2d_fwd(xs[i], wf, 3, 1, c, out=ys[i])
    torch.cuda.synchronize()
    iters = 30
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(iters):
        ops.conv2d_fwd(xs[i % nbuf], wf, 3, 1, c, out=ys[i % nbuf])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 2.0 * n * h * w * c * c * 9
    ach = flop / (ms * 1e-3) / 1e12
    return dict(bound='tensor', kernel='igemm2_kernel<256,false> conv3x3 256->256 @ 8x128x128 (persistent, TMEM double-buffered, TMA store)', achieved=ach, peak=pk['tf_burst'],
                unit='TFLOP/s', frac=ach / pk['tf_burst'], traffic=NCU_TRAFFIC_BYTES, peak_source=pk['src'] + ' bf16_tflops (burst)',
                ms_per_launch=ms, flop_per_launch=flop)


def bn_roofline(pk):
    """The largest HBM-bound kernel class of the step (in-situ cost, profiles/r01_knockout_insitu_cost.json): BatchNorm
    backward (reduce + finalize + apply, cp.async-staged) on the 8 x 128 x 128 x 256 layer shape.  Algorithmic bytes =
    read dy, x twice + write dx = 5 x M x C x 2 B; rotating buffer sets > L2, CUDA events on the launch stream."""
    import ctypes
    from ever_b200._lib import check, lib, ptr, stream
    L = lib()
    c_int, c_ll = ctypes.c_int, ctypes.c_longlong
    m, c, nset = 131072, 256, 3
    xs, dys, dxs = [[torch.randn(m, c, device='cuda').bfloat16() for _ in range(nset)] for _ in range(3)]
    st = torch.rand(4, c, device='cuda') + 0.5
    dg, db = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
    ws = torch.empty(L.evb_bn_workspace(c_ll(m), c_int(c)) // 4, device='cuda')

    def run(i):
        check(L.evb_bn_bwd(ptr(dys[i]), ptr(xs[i]), None, ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), c_int(2), c_int(0),
                           ptr(dxs[i]), None, c_int(0), ptr(dg), ptr(db), c_int(0), c_ll(m), c_int(c), ptr(ws), stream()),
              'evb_bn_bwd')
    for i in range(nset):
        run(i)
    torch.cuda.synchronize()
    iters = 30
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(iters):
        run(i % nset)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = 5.0 * m * c * 2
    ach = nbytes / (ms * 1e-3) / 1e9
    return dict(bound='hbm', kernel='evb_bn_bwd = bn_bwd_reduce_ca<2> + bn_bwd_finalize2 + bn_bwd_apply_ca<2> on [131072, 256] bf16',
                achieved=ach, peak=pk['hbm'], unit='GB/s', frac=ach / pk['hbm'], traffic=134321664 + 4555776 + 134238976 + 28960768,
                peak_source=pk['src'] + ' hbm_gbs (copy)', ms_per_launch=ms, bytes_per_launch=nbytes)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from ever_b200 import _lib
    from ever_b200.module import ChangeStarB200, FarSegB200
    T0 = time.time()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)

    def log(msg):
        if os.environ.get('EVB_BENCH_VERBOSE', '1') == '1':
            sys.stderr.write('[bench rank %d %.1fs] %s\n' % (rank, time.time() - T0, msg))
            sys.stderr.flush()

    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.manual_seed(0)
    from ever_b200.freenet import FreeNetB200
    cls = dict(ChangeStar=ChangeStarB200, FreeNet=FreeNetB200).get(cfg['model'], FarSegB200)
    model = cls(model_config(cfg)).cuda().train()
    eng = model._engine()
    eng.set_distributed(rank, world)
    if os.environ.get('EVB_SYNC_DICE', '1') == '0':
        eng.sync_dice = False
    if world > 1:  # same initial weights everywhere (DDP broadcasts rank 0's at construction)
        dist.broadcast(eng.flat_w, 0)
        torch.cuda.synchronize()
    log('model built, nccl ok')
    n_units = units_per_gpu(cfg, world)
    n_tiles = tiles_of(cfg, n_units)
    xh, yh = synthetic(cfg, n_units, seed=rank)
    xh = xh.pin_memory()
    yh = {k_: v.pin_memory() for k_, v in yh.items()}
    x = xh.cuda()
    y = {k_: v.cuda() for k_, v in yh.items()}
    labels = y['cls'] if cfg['model'] == 'FarSeg' else y
    lr = 0.007

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic work of one step, counted from the C-ABI calls of one eager step
    acct = eng.accounting(True)
    eng.forward_train(x, labels)
    eng.backward(allreduce=False)
    eng.sgd_step(lr)
    work = eng.accounting(False).totals()
    torch.cuda.synchronize()

    # ---- value: CUDA-graph replay, inputs resident in HBM
    use_graph = os.environ.get('EVB_NO_GRAPH', '0') != '1'
    graph = None
    l0 = _lib.launches[0]
    if use_graph:
        log('capturing step graph')
        graph, out = eng.capture_step(x, labels)
        per_step_launches = (_lib.launches[0] - l0) // 3 + 3
    else:
        out = eng.forward_train(x, labels)
        eng.backward(allreduce=False)
        per_step_launches = _lib.launches[0] - l0 + 3

    def step():
        nonlocal out
        if graph is not None:
            graph()
        else:
            out = eng.forward_train(x, labels)
            eng.backward(allreduce=False)
        if os.environ.get('EVB_BENCH_NO_AR', '0') == '1':     # diagnostics only: lower bound without the gradient exchange
            for w_ in eng._ar_pending:
                w_.wait()
            eng._ar_pending = []
        else:
            eng.allreduce_grads()
        eng.sgd_step(lr)

    log('graph ready, warm-up')
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    log('timing')
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * n_tiles / (ms * 1e-3)
    loss_now = {k: float(v) for k, v in out.items()}

    log('value done: %.2f ms/step' % ms)
    # ---- e2e: plugin API with host buffers
    e2e_steps = max(3, min(args.steps, 50))

    # double-buffered input pipeline: the H2D copy of step i+1 (pinned host -> device staging, copy stream) overlaps the
    # compute of step i; every step still pays its own H2D + a D2H read of the losses inside the timed region
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(x), {k_: torch.empty_like(v) for k_, v in y.items()}) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = dict(i=0)
    loss_host = [torch.empty(len(out), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            stage[slot][0].copy_(xh, non_blocking=True)
            for k_, v in yh.items():
                stage[slot][1][k_].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    for ev in consumed:
        ev.record()
    prefetch(0)

    def e2e_step():
        slot = state['i'] & 1
        state['i'] += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[slot])
        prefetch(slot ^ 1)
        # the plugin call a user makes: model(x, y) -> loss dict, model.backward(...); with config.cuda_graph the forward
        # copies the batch into static buffers, replays the cached CUDA graph of forward + loss + backward, and
        # backward() runs the gradient all-reduce
        o = model(stage[slot][0], stage[slot][1])
        consumed[slot].record(cur)
        model.backward(o, None, None)
        eng.sgd_step(lr)
        # D2H read of this step's losses into pinned memory, every step; the host blocks on the PREVIOUS step's copy, so the
        # read-back is pipelined one step deep (the host launches step i+1 while step i's losses travel) instead of
        # draining the GPU at every step
        loss_host[slot].copy_(torch.stack([v.detach() for v in o.values()]), non_blocking=True)
        loss_ready[slot].record(cur)
        if state['i'] > 1:
            loss_ready[slot ^ 1].synchronize()
            state['last_loss'] = loss_host[slot ^ 1].tolist()

    model.config.cuda_graph = use_graph
    for _ in range(3):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    loss_ready[(state['i'] - 1) & 1].synchronize()     # the last step's losses have arrived on the host
    state['last_loss'] = loss_host[(state['i'] - 1) & 1].tolist()
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1) / e2e_steps
    t = torch.tensor([ms2], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms2 = float(t)
    h2d = int(xh.numel() * xh.element_size() + sum(v.numel() * v.element_size() for v in yh.values()))
    e2e = dict(value=world * n_tiles / (ms2 * 1e-3), unit='tiles/s', h2d_bytes_per_step=h2d,
               d2h_bytes_per_step=4 * len(out), ms_per_step=ms2, steps=e2e_steps,
               api='%s.forward(x, y) + .backward() via libevb200.so C ABI; pinned host inputs, double-buffered H2D on a copy '
                   'stream, D2H read of the losses every step (pinned, host waits one step behind)' % cls.__name__)

    if rank == 0:
        pk = peaks()
        roof = conv_roofline(pk)
        roof_hbm = bn_roofline(pk)
        line = dict(metric=metric_name(cfg), value=value, unit='tiles/s', n_gpus=world, steps=args.steps,
                    warmup=max(3, args.warmup), ms_per_step=ms, higher_is_better=True, scaling=cfg['scaling'], vs_baseline=None,
                    dtype='bf16', data='synthetic',
                    config=dict(workload=cfg['workload'] + ', fwd + CE/Dice loss + bwd + grad all-reduce + clip/SGD',
                                config=args.config, global_batch=world * n_tiles, tiles_per_gpu=n_tiles, tile=list(cfg['hw']),
                                parallelism='dp%d' % world, cuda_graph=graph is not None,
                                l2='step working set (GBs of activations) >> 126 MB L2; no explicit flush'),
                    e2e=e2e, gpu_launches=int(per_step_launches * args.steps), clocks=clocks, roofline=roof,
                    roofline_hbm=roof_hbm, losses=loss_now)
        # step-level accounting: algorithmic HBM bytes and conv FLOPs of one step, as scheduled by the engine
        step_tf = work['flops'] / (ms * 1e-3) / 1e12
        line['flops_per_step'] = work['flops']
        line['hbm_bytes_per_step'] = work['bytes']
        line['step_tflops_per_gpu'] = step_tf
        line['step_frac_of_sustained_tensor_peak'] = step_tf / pk['tf_sust']
        line['step_hbm_gbs'] = work['bytes'] / (ms * 1e-3) / 1e9
        line['step_frac_of_hbm_peak'] = line['step_hbm_gbs'] / pk['hbm']
        line['step_roofline_ms'] = dict(tensor=work['flops'] / (pk['tf_sust'] * 1e12) * 1e3,
                                        hbm=work['bytes'] / (pk['hbm'] * 1e9) * 1e3)
        if cfg['gflop_per_tile']:
            tfs = value / world * cfg['gflop_per_tile'] / 1e3
            line['model_tflops_per_gpu'] = tfs
            line['model_frac_of_sustained_peak'] = tfs / pk['tf_sust']
        if world == 1 and not args.no_cpu:
            log('cpu baseline')
            cpu_val, cpu_sec, cores, kind, ran = cpu_reference(cfg, CPU_SAMPLE[args.config], 4, 1)
            line['cpu_baseline'] = dict(value=cpu_val, unit='tiles/s', cores=cores, kind=kind,
                                        sample='%d units of the per-GPU batch per step, 1 warm-up + %d timed fwd+loss+bwd '
                                               'steps (%.1f s of CPU work), fp32 torch CPU, all host cores'
                                               % (CPU_SAMPLE[args.config], ran, cpu_sec * (ran + 1)))
        if world == 1 and not args.no_incumbent:
            log('gpu incumbent')
            del graph
            variants = ['nchw', 'channels_last'] + ([] if args.no_incumbent_compile else ['compile'])
            line['gpu_incumbent'] = gpu_incumbent(cfg, n_units, variants)
        print(json.dumps(line))
    if world > 1:
        # the captured step graph holds NCCL kernel nodes: release every graph before the communicator goes away, and never
        # let a stuck teardown (seen with live captured collectives) keep the job from exiting after the line is printed
        sys.stdout.flush()
        graph = None
        eng._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


def main():
    if os.environ.get('EVB_BENCH_WATCHDOG'):   # diagnostics: dump every thread's stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['EVB_BENCH_WATCHDOG']), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-incumbent', action='store_true', help='skip the gpu_incumbent leg')
    ap.add_argument('--no-incumbent-compile', action='store_true',
                    help='skip the torch.compile variant of the gpu_incumbent leg (it adds ~40 s of compilation)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
