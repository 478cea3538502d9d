import sys, copy, statistics
sys.path.insert(0, '.')
import torch, torch.nn.functional as F
from oracle.farseg_oracle import FarSegOracle, deterministic_fill, dice_loss_oracle, synthetic_batch
def rel(a, b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-12))
torch.backends.cudnn.allow_tf32 = False
def run(o, x, y, ac):
    o.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=ac):
        lg = o.logits(x)
        loss = F.cross_entropy(lg, y, ignore_index=255) + dice_loss_oracle(lg, y)
    loss.backward()
    return {k: p.grad.clone() for k, p in o.named_parameters()}
for resnet, k, n, hw in [('resnet18', 5, 4, 256), ('resnet50', 15, 4, 256)]:
    o = deterministic_fill(FarSegOracle(resnet, k, 128), 0).cuda().train()
    x, y = synthetic_batch(n, hw, hw, k); x, y = x.cuda(), y.cuda()
    # structured labels: class = argmax of K fixed random projections of the 9x9-blurred image
    g = torch.Generator(device='cuda').manual_seed(5)
    proj = torch.randn(k, 3, 1, 1, device='cuda', generator=g)
    ys = F.conv2d(F.avg_pool2d(x, 9, 1, 4), proj).argmax(1)
    ys[y == 255] = 255
    for name, lab in (('random', y), ('structured', ys)):
        for steps in (0, 20):
            oo = copy.deepcopy(o)
            opt = torch.optim.SGD(oo.parameters(), lr=0.01, momentum=0.9)
            for _ in range(steps):   # a few SGD steps: gradients stop being noise-dominated
                opt.zero_grad(); lg = oo.logits(x); (F.cross_entropy(lg, lab, ignore_index=255) + dice_loss_oracle(lg, lab)).backward(); opt.step()
            gb = run(oo, x, lab, True); g32 = run(oo, x, lab, False)
            v = sorted(rel(gb[kk], g32[kk]) for kk in gb if not (kk.endswith('0.bias') and 'encoders' in kk))
            print(resnet, name, 'sgd_steps', steps, 'median %.3f p90 %.3f max %.3f' % (statistics.median(v), v[int(.9*len(v))], v[-1]))
