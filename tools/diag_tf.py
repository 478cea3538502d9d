"""Diagnostics behind the teacher-forced parity test: for the tensors where engine and bf16 oracle differ by more than
1e-2, recompute the op in fp64 from the oracle's own captured inputs and report |engine - truth| and |oracle - truth|
(who is noisy?), and probe torch's conv+bias rounding.  Run on the GPU box: python tools/diag_tf.py"""
import json
import sys

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import torch
import torch.nn.functional as F

from _helpers import TeacherForcing, rel_l2
from test_teacher_forced_gpu import _build, oracle_step_captured
from oracle.farseg_oracle import synthetic_batch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
out = {}

# ---- 1. what does torch's bf16 conv + bias round?
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(2, 256, 32, 32, device='cuda', generator=g).bfloat16()
w = (torch.randn(256, 256, 1, 1, device='cuda', generator=g) * 0.05).bfloat16()
b = torch.randn(256, device='cuda', generator=g)
y = F.conv2d(x, w, b.bfloat16())
acc = F.conv2d(x.float(), w.float())
cand = dict(single_round_fp32_bias=(acc + b.view(1, -1, 1, 1)).bfloat16(),
            single_round_bf16_bias=(acc + b.bfloat16().float().view(1, -1, 1, 1)).bfloat16(),
            double_round=(acc.bfloat16().float() + b.bfloat16().float().view(1, -1, 1, 1)).bfloat16())
out['conv_bias_rounding'] = {k: dict(equal_frac=float((v == y).float().mean()), rel=rel_l2(v.float(), y.float()))
                             for k, v in cand.items()}

resnet, k, dec, n, h, w_ = 'resnet18', 5, 128, 2, 256, 256
ora, mine = _build(resnet, k, dec)
xx, yy = synthetic_batch(n, h, w_, k)
xx, yy = xx.cuda(), yy.cuda()
cap, _ = oracle_step_captured(ora, xx, yy)


class Keep(TeacherForcing):
    def __init__(self, cap):
        super().__init__(cap, True)
        self.mine = dict(fwd={}, bwd={})

    def __call__(self, kind, name, t):
        self.mine[kind][name] = t.detach().clone()
        super().__call__(kind, name, t)


tf = Keep(cap)
eng = mine._engine()
eng.tf = tf
dbg = {}
eng.debug = dbg
o = mine(xx, dict(cls=yy))
mine.backward(o, None, None)
torch.cuda.synchronize()
pm, po = dict(mine.named_parameters()), dict(ora.named_parameters())


def nhwc(t):
    return t.permute(0, 3, 1, 2)


# ---- 2. BN + ReLU in front of the max-pool: fp64 truth from the oracle's captured tensors
xs = cap.fwd['en.resnet.conv1'].double().requires_grad_(True)
gam = po['en.resnet.bn1.weight'].detach().double().requires_grad_(True)
bet = po['en.resnet.bn1.bias'].detach().double().requires_grad_(True)
ys = F.relu(F.batch_norm(xs, None, None, gam, bet, True, 0.1, 1e-5))
dy = cap.bwd['en.resnet.relu#0'].double()
ys.backward(dy)
rep = {}
for nm, truth, mine_t, ora_t in (('dbeta', bet.grad, pm['en.resnet.bn1.bias'].grad, po['en.resnet.bn1.bias'].grad),
                                 ('dgamma', gam.grad, pm['en.resnet.bn1.weight'].grad, po['en.resnet.bn1.weight'].grad),
                                 ('dx', xs.grad, nhwc(tf.mine['bwd']['en.resnet.conv1']), cap.bwd['en.resnet.conv1'])):
    rep[nm] = dict(engine_vs_truth=rel_l2(mine_t, truth), oracle_vs_truth=rel_l2(ora_t, truth), engine_vs_oracle=rel_l2(mine_t, ora_t))
rep['sum_abs_over_abs_sum'] = float(dy.abs().sum() / dy.sum(dim=(0, 2, 3)).abs().sum())
out['stem_bn'] = rep

# ---- 3. FS-Relation level i: fp64 truth of (du1, du2, dsf) from the oracle's captured inputs
for i in (0, 3):
    pfx = 'head.fs_relation.'
    u1 = cap.fwd[pfx + 'content_encoders.%d.0' % i].double().requires_grad_(True)
    u2 = cap.fwd[pfx + 'feature_reencoders.%d.0' % i].double().requires_grad_(True)
    sf = cap.fwd[pfx + 'scene_encoder.%d.2' % i].double().requires_grad_(True)
    dz = cap.bwd['head.fpn_decoder.blocks.%d.0.0:in' % i].double()

    def bn(u, key):
        return F.relu(F.batch_norm(u, None, None, po[pfx + key + '.1.weight'].detach().double(),
                                   po[pfx + key + '.1.bias'].detach().double(), True, 0.1, 1e-5))
    cf, p = bn(u1, 'content_encoders.%d' % i), bn(u2, 'feature_reencoders.%d' % i)
    r = torch.sigmoid((sf * cf).sum(1, keepdim=True))
    z = r * p
    z.backward(dz)
    r_ora = torch.sigmoid((cap.fwd[pfx + 'scene_encoder.%d.2' % i] * cap.fwd[pfx + 'content_encoders.%d.2' % i]).float().sum(1, keepdim=True))
    rep = dict(r_engine_vs_truth=rel_l2(dbg['rel%d' % i].view_as(r), r), r_oracle_vs_truth=rel_l2(r_ora, r),
               z_engine_vs_truth=rel_l2(nhwc(tf.mine['fwd']['head.fpn_decoder.blocks.%d.0.0:in' % i]), z),
               z_oracle_vs_truth=rel_l2(cap.fwd['head.fpn_decoder.blocks.%d.0.0:in' % i], z))
    for nm, truth, key in (('du1', u1.grad, pfx + 'content_encoders.%d.0' % i), ('du2', u2.grad, pfx + 'feature_reencoders.%d.0' % i),
                           ('dsf', sf.grad, pfx + 'scene_encoder.%d.2' % i)):
        m_t = tf.mine['bwd'][key]
        m_t = nhwc(m_t) if m_t.dim() == 4 else m_t.view_as(truth)
        rep[nm] = dict(engine_vs_truth=rel_l2(m_t, truth), oracle_vs_truth=rel_l2(cap.bwd[key], truth),
                       engine_vs_oracle=rel_l2(m_t, cap.bwd[key]))
    out['relation_level%d' % i] = rep
print(json.dumps(out, indent=1))
json.dump(out, open('gpurun_out/diag_tf.json', 'w'), indent=1)
