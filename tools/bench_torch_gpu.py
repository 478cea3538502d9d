"""The incumbent on the same box: the oracle (= the reference's PyTorch modules restated) on the GPU through
torch/cuDNN, eager, bf16 autocast as Launcher runs it (NCHW), plus channels_last -- FarSeg-R50, 8x3x512x512."""
import json
import sys
sys.path.insert(0, '.')
import torch
import torch.nn.functional as F
from bench import K_CLASSES, PER_GPU_BATCH, synthetic
from oracle.farseg_oracle import FarSegOracle, dice_loss_oracle

res = {}
for fmt in ('nchw', 'channels_last'):
    torch.manual_seed(0)
    m = FarSegOracle('resnet50', K_CLASSES, 256).cuda().train()
    x, y = synthetic(PER_GPU_BATCH)
    x, y = x.cuda(), y.cuda()
    if fmt == 'channels_last':
        m = m.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    opt = torch.optim.SGD(m.parameters(), lr=0.007, momentum=0.9, weight_decay=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            lg = m.logits(x)
            loss = F.cross_entropy(lg, y, ignore_index=255) + dice_loss_oracle(lg, y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 35.0)
        opt.step()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(15):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 15
    res[fmt] = dict(ms_per_step=ms, tiles_per_s=PER_GPU_BATCH / ms * 1e3)
print(json.dumps(dict(impl='torch+cuDNN eager bf16 autocast (oracle modules) on the same B200', **res)))
