"""Eager (no CUDA graph) training steps of a bench.py config for profiling under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <N> --csv --log-file gpurun_out/launches.csv \
        python tools/ncu_step.py [steps] [config]"""
import sys
sys.path.insert(0, '.')
import torch
from bench import CONFIGS, model_config, synthetic, units_per_gpu
from ever_b200 import _lib
from ever_b200.freenet import FreeNetB200
from ever_b200.module import ChangeStarB200, FarSegB200
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = CONFIGS[sys.argv[2] if len(sys.argv) > 2 else 'c2']
torch.manual_seed(0)
m = dict(ChangeStar=ChangeStarB200, FreeNet=FreeNetB200).get(cfg['model'], FarSegB200)(model_config(cfg)).cuda().train()
x, y = synthetic(cfg, units_per_gpu(cfg, 1))
x, y = x.cuda(), {k: v.cuda() for k, v in y.items()}
for i in range(nsteps):
    l0 = _lib.launches[0]
    out = m(x, y)
    m.backward(out, None, None)
    m.engine.sgd_step(0.007)
    torch.cuda.synchronize()
    print('step', i, 'launches', _lib.launches[0] - l0)
print({k: float(v) for k, v in out.items()})
