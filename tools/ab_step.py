"""A/B timing of the captured FarSeg-R50 8x512^2 step under library tuning knobs, in ONE process on one GPU
(box-to-box variance is larger than most single-kernel gains).  usage: python tools/ab_step.py name=setter:value,... ..."""
import ctypes
import json
import sys

sys.path.insert(0, '.')
import torch  # noqa: E402
from bench import PER_GPU_BATCH, farseg_config, synthetic  # noqa: E402
from ever_b200._lib import lib  # noqa: E402
from ever_b200.module import FarSegB200  # noqa: E402


def run(knobs, iters=30):
    import os
    L = lib()
    for fn, v in knobs:
        if fn.startswith('EVB_'):      # engine-level switch read from the environment when the engine is built
            os.environ[fn] = str(v)
        else:
            getattr(L, fn)(ctypes.c_int(v))
    torch.manual_seed(0)
    m = FarSegB200(farseg_config()).cuda().train()
    eng = m._engine()
    x, y = synthetic(PER_GPU_BATCH)
    replay, out = eng.capture_step(x.cuda(), y.cuda())
    for _ in range(5):
        replay(); eng.sgd_step(0.007)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            replay(); eng.sgd_step(0.007)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    loss = {k: float(v) for k, v in out.items()}
    del m, eng, replay
    torch.cuda.empty_cache()
    return best, loss


if __name__ == '__main__':
    configs = []
    for a in sys.argv[1:]:
        name, _, spec = a.partition('=')
        knobs = [(kv.split(':')[0], int(kv.split(':')[1])) for kv in spec.split(',') if kv]
        configs.append((name, knobs))
    res = {}
    for rep in range(2):
        for name, knobs in configs:
            ms, loss = run(knobs)
            res.setdefault(name, []).append(round(ms, 4))
            print(json.dumps(dict(config=name, rep=rep, ms=ms, tiles_per_s=PER_GPU_BATCH / ms * 1e3, loss=loss)), flush=True)
    print(json.dumps(res))
