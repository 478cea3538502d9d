"""One eager (no CUDA graph) FarSeg-R50 training step for profiling under ncu."""
import sys
sys.path.insert(0, '.')
import torch
from bench import farseg_config, synthetic, PER_GPU_BATCH
from ever_b200.module import FarSegB200
torch.manual_seed(0)
m = FarSegB200(farseg_config()).cuda().train()
x, y = synthetic(PER_GPU_BATCH)
x, y = x.cuda(), y.cuda()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(nsteps):
    out = m(x, dict(cls=y))
    m.backward(out, None, None)
    m.engine.sgd_step(0.007)
torch.cuda.synchronize()
print({k: float(v) for k, v in out.items()})
