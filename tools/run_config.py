"""Run one training step of another BASELINE.json config on the GPU (memory / generality check)."""
import sys, time
sys.path.insert(0, '.')
import torch
from ever_b200.module import FarSegB200
resnet, k, n, hw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torch.manual_seed(0)
m = FarSegB200(dict(encoder=dict(resnet_type=resnet), head=dict(fpn_decoder=dict(classifier_config=dict(num_classes=k))))).cuda().train()
x = torch.randn(n, 3, hw, hw, device='cuda'); y = torch.randint(0, k, (n, hw, hw), device='cuda')
for i in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    out = m(x, dict(cls=y)); m.backward(out, None, None); m.engine.sgd_step(0.007)
    torch.cuda.synchronize()
    print(resnet, n, hw, {kk: round(float(v), 4) for kk, v in out.items()}, 'step %.1f ms' % ((time.time() - t0) * 1e3), 'mem %.1f GB' % (torch.cuda.max_memory_allocated() / 2**30))
