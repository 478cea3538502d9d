#!/usr/bin/env python
"""Benchmark of the hot path: FarSeg-R50, 15 classes, 8 x 3 x 512 x 512 synthetic tiles per GPU (BASELINE.json
configs[1]), one step = forward + CE/Dice loss + backward (+ NCCL gradient all-reduce at N>1) + fused clip/SGD.

    python bench.py --gpus N --steps K --warmup W            # B200 engine (libevb200.so), prints ONE JSON line
    python bench.py --impl reference --steps K --warmup W    # the reference's PyTorch-CPU path (oracle port)

value  = whole-job tiles/s with inputs resident in HBM (CUDA-graph replay of the step, CUDA events, max over ranks)
e2e    = the same metric through the plugin call model(x, y) / model.backward() with pinned HOST inputs:
         H2D copies of the tile batch + labels and a D2H read of the losses inside the timed region.
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

K_CLASSES, TILE, PER_GPU_BATCH = 15, 512, 8
GFLOP_PER_TILE = 343.1  # fwd 114.8 + bwd 228.3, FlopCounterMode on the reference modules
METRIC = 'FarSeg-R50 512x512 tiles/s fwd+bwd'
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu --set full capture
# summarised in profiles/ (None until captured)
NCU_TRAFFIC_BYTES = 68309248 + 24216576  # profiles/r01_ncu_full_prof_igemm2_3x3.txt


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
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == 'active'})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def farseg_config():
    return dict(encoder=dict(resnet_type='resnet50'),
                head=dict(fpn_decoder=dict(out_channels=256, classifier_config=dict(num_classes=K_CLASSES))))


def synthetic(n, seed=0):
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(n, 3, TILE, TILE, generator=g)
    g = torch.Generator().manual_seed(4321 + seed)
    y = torch.randint(0, K_CLASSES, (n, TILE, TILE), generator=g)
    g = torch.Generator().manual_seed(99 + seed)
    y[torch.rand(n, TILE, TILE, generator=g) < 0.05] = 255
    return x, y


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference(n_tiles, iters, warmup):
    """The reference's PyTorch-CPU path (fp32; Launcher cannot autocast on CPU, ever/core/launcher.py:194),
    restated in oracle/farseg_oracle.py, on all host cores."""
    from oracle.farseg_oracle import FarSegOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = FarSegOracle('resnet50', K_CLASSES, 256).train()
    x, y = synthetic(n_tiles)
    ts = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        m.zero_grad(set_to_none=True)
        losses = m(x, dict(cls=y))
        sum(losses.values()).backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    sec = sum(ts) / len(ts)
    return n_tiles / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_tiles = 4
    val, sec, cores = cpu_reference(n_tiles, max(1, min(args.steps, 5)), max(1, min(args.warmup, 1)))
    sample = '%d tiles of 3x512x512 per step (bounded sample of the 8-tile batch), fp32, torch CPU, %d threads' % (n_tiles, cores)
    line = dict(metric=METRIC, value=val, unit='tiles/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference',
                config=dict(workload='FarSeg-R50 15-class 3x512x512 synthetic tiles, fwd+loss+bwd', tiles_per_step=n_tiles),
                cpu_baseline=dict(value=val, unit='tiles/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=val, unit='tiles/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ dominant-kernel roofline
def conv_roofline(pk):
    """3x3 256->256 conv on 8 x 128 x 128 (33.7 % of forward FLOPs), igemm_kernel<256>, timed alone
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
    from ever_b200.module import FarSegB200
    T0 = time.time()
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
    model = FarSegB200(farseg_config()).cuda().train()
    eng = model._engine()
    eng.set_distributed(rank, world)
    if world > 1:  # same initial weights everywhere (DDP broadcasts rank 0's at construction)
        dist.broadcast(eng.flat_w, 0)
        torch.cuda.synchronize()
    log('model built, nccl ok')
    xh, yh = synthetic(PER_GPU_BATCH, seed=rank)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    x, y = xh.cuda(), yh.cuda()
    lr = 0.007

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: CUDA-graph replay, inputs resident in HBM
    use_graph = os.environ.get('EVB_NO_GRAPH', '0') != '1'
    graph = None
    l0 = _lib.launches[0]
    if use_graph:
        log('capturing step graph')
        graph, out = eng.capture_step(x, y)
        per_step_launches = (_lib.launches[0] - l0) // 3 + 3
    else:
        out = model(x, dict(cls=y))
        model.backward(out, None, None)
        per_step_launches = _lib.launches[0] - l0 + 3

    def step():
        nonlocal out
        if graph is not None:
            graph()
        else:
            out = eng.forward_train(x, y)
            eng.backward(allreduce=False)
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
    value = world * PER_GPU_BATCH / (ms * 1e-3)
    loss_now = {k: float(v) for k, v in out.items()}

    log('value done: %.2f ms/step' % ms)
    # ---- e2e: plugin API with host buffers
    e2e_steps = max(3, min(args.steps, 10))

    # double-buffered input pipeline: the H2D copy of step i+1 (pinned host -> device staging, copy stream) overlaps the
    # compute of step i; every step still pays its own H2D + a D2H read of the losses inside the timed region
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = dict(i=0)

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            stage[slot][0].copy_(xh, non_blocking=True)
            stage[slot][1].copy_(yh, non_blocking=True)
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
        o = model(stage[slot][0], dict(cls=stage[slot][1]))
        consumed[slot].record(cur)
        model.backward(o, None, None)
        eng.sgd_step(lr)
        return torch.stack([o['ce_loss'], o['dice_loss']]).cpu()

    model.config.cuda_graph = use_graph
    for _ in range(3):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1) / e2e_steps
    t = torch.tensor([ms2], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms2 = float(t)
    e2e = dict(value=world * PER_GPU_BATCH / (ms2 * 1e-3), unit='tiles/s',
               h2d_bytes_per_step=int(xh.numel() * 4 + yh.numel() * 8), d2h_bytes_per_step=8, ms_per_step=ms2,
               api='FarSegB200.forward(x, y) + .backward() via libevb200.so C ABI; pinned host inputs, double-buffered H2D on a copy '
                   'stream, D2H read of the losses every step')

    if rank == 0:
        pk = peaks()
        roof = conv_roofline(pk)
        roof_hbm = bn_roofline(pk)
        cpu_val, cpu_sec, cores = cpu_reference(4, 4, 1) if world == 1 and not args.no_cpu else (None, None, None)
        tfs = value / world * GFLOP_PER_TILE / 1e3
        line = dict(metric=METRIC, value=value, unit='tiles/s', n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
                    config=dict(workload='FarSeg-R50 (ResNet-50 + FPN + FS-Relation + asymmetric decoder, 256-ch), 15-class, '
                                         '8x3x512x512 synthetic tiles per GPU, fwd + CE/Dice loss + bwd + grad all-reduce + '
                                         'clip/SGD', global_batch=world * PER_GPU_BATCH, tile=TILE,
                                parallelism='dp%d' % world, cuda_graph=graph is not None,
                                l2='step working set (~10 GB of activations) >> 126 MB L2; no explicit flush'),
                    e2e=e2e, gpu_launches=int(per_step_launches * args.steps), clocks=clocks, roofline=roof,
                    roofline_hbm=roof_hbm,
                    model_tflops_per_gpu=tfs, model_frac_of_sustained_peak=tfs / pk['tf_sust'], losses=loss_now)
        if cpu_val is not None:
            line['cpu_baseline'] = dict(value=cpu_val, unit='tiles/s', cores=cores, kind='port',
                                        sample='4 tiles of 3x512x512 per step (half the 8-tile batch), 1 warm-up + 4 timed '
                                               'fwd+loss+bwd steps (~7 s of CPU work), fp32 torch CPU, all host cores')
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
