"""Step time (CUDA-graph replay, CUDA events) of the other BASELINE.json configs on one B200 -- not bench lines,
evidence that the engine covers them: C1 R18 2x256^2 K=5; C3 ChangeStar-R50 pairs of 512^2; C4 R101 4x1024^2 K=7;
C5-shape: a 1x200x624x352 hyperspectral cube (PaviaU 610x340 padded to /16 as FreeNet does, then /32) through the in-tree
high-channel entry ResNetEncoder(in_channels=200) (ever/module/resnet.py:100-117) + FarSegHead -- FreeNet itself is not in
the reference tree."""
import json
import sys
sys.path.insert(0, '.')
import torch
from ever_b200.module import ChangeStarB200, FarSegB200


def timed(replay, eng, iters=15):
    for _ in range(3):
        replay(); eng.sgd_step(0.007)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        replay(); eng.sgd_step(0.007)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
only = sys.argv[1:]
for name, cls, resnet, k, n, hw, dec in [('C1 FarSeg-R18 2x256^2 K=5', FarSegB200, 'resnet18', 5, 2, 256, 128),
                                          ('C5-shape FarSeg-R50 in_channels=200 1x200x640x352 K=9', FarSegB200, 'resnet50', 9, 1, (640, 352), 256),
                                          ('C2 FarSeg-R50 1x512^2 K=15 (strong-scaling shard)', FarSegB200, 'resnet50', 15, 1, 512, 256),
                                          ('C3 ChangeStar-R50 1 pair 512^2 (per-GPU shard at 8 GPUs)', ChangeStarB200, 'resnet50', 1, 1, 512, 256),
                                          ('C3 ChangeStar-R50 8 pairs 512^2', ChangeStarB200, 'resnet50', 1, 8, 512, 256),
                                          ('C4 FarSeg-R101 4x1024^2 K=7', FarSegB200, 'resnet101', 7, 4, 1024, 256)]:
    if only and not any(name.startswith(o) for o in only):
        continue
    torch.manual_seed(0)
    cin = 200 if 'in_channels=200' in name else 3
    m = cls(dict(encoder=dict(resnet_type=resnet, in_channels=cin),
                 head=dict(fpn_decoder=dict(out_channels=dec, classifier_config=dict(num_classes=k))))).cuda().train()
    eng = m._engine()
    cs = cls is ChangeStarB200
    hh, ww = hw if isinstance(hw, tuple) else (hw, hw)
    x = torch.randn(n, 2 * cin if cs else cin, hh, ww, device='cuda')
    y = torch.randint(0, max(k, 2), (n, hh, ww), device='cuda')
    labels = dict(cls=y, change=torch.randint(0, 2, (n, hh, ww), device='cuda')) if cs else y
    replay, out = eng.capture_step(x, labels)
    ms = timed(replay, eng)
    row = dict(config=name, ms_per_step=ms, tiles_per_s=n * (2 if cs else 1) / ms * 1e3, images_per_step=n * (2 if cs else 1),
               mem_gb=torch.cuda.max_memory_allocated() / 2**30, losses={kk: float(v) for kk, v in out.items()})
    rows.append(row)
    print(json.dumps(row))
    del m, eng, replay
    torch.cuda.empty_cache()
json.dump(rows, open('gpurun_out/bench_configs.json', 'w'), indent=1)
