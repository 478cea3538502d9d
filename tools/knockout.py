"""In-situ cost of each kernel class: time the captured FarSeg-R50 8x512^2 step with one class of C-ABI calls replaced by
no-ops (results are garbage, only the timing is meaningful).  Isolated micro-benchmarks and cold-cache ncu durations
over-state kernels whose inputs are L2-resident in the real step and kernels that overlap the weight-gradient branch; the
drop in step time when a class is removed is what optimising it can buy at most.  PROFILING TOOL ONLY."""
import json
import sys

sys.path.insert(0, '.')
import torch  # noqa: E402
from bench import CONFIGS, model_config, synthetic, units_per_gpu  # noqa: E402

CFG = CONFIGS['c2']
from ever_b200._lib import lib  # noqa: E402
from ever_b200.module import FarSegB200  # noqa: E402

CLASSES = {
    'none': [],
    'bn_apply(fwd)': ['evb_bn_apply'],
    'bn_bwd': ['evb_bn_bwd'],
    'bn_finalize(fwd)': ['evb_bn_finalize', 'evb_bn_stats'],
    'bilinear_up': ['evb_bilinear_up'],
    'bilinear_bwd': ['evb_bilinear_up_bwd_sep'],
    'loss': ['evb_loss_stats', 'evb_loss_grad'],
    'stem_im2col': ['evb_im2col_nchw'],
    'maxpool': ['evb_maxpool3x3s2_fwd', 'evb_maxpool3x3s2_bwd'],
    'relation': ['evb_relation_fwd', 'evb_relation_bwd'],
    'pack_weights': ['evb_pack_weights_tiled'],
    'wgrad': ['evb_conv2d_wgrad'],
    'conv_fwd': ['evb_conv2d_fwd', 'evb_conv2d_fwd_stats', 'evb_conv2d_fwd_bias_stats'],
    'dgrad': ['evb_conv2d_dgrad'],
    'merge/scale_add/sumpool': ['evb_merge4', 'evb_scale_add', 'evb_sumpool2'],
    'gap+linear': ['evb_gap_fwd', 'evb_gap_bwd', 'evb_linear_fwd', 'evb_linear_bwd'],
    'bias_grad+copy2d': ['evb_bias_grad', 'evb_copy2d_f32'],
    'sgd': ['evb_grad_norm', 'evb_sgd_step'],
}


def run(names, iters=20):
    L = lib()
    saved = {}
    for n in names:
        saved[n] = getattr(L, n)
        if n in ('evb_conv2d_fwd_stats', 'evb_conv2d_fwd_bias_stats'):
            def fake(*a, _i=12 if n == 'evb_conv2d_fwd_stats' else 13):
                a[_i]._obj.value = 1
                return 0
            setattr(L, n, fake)
        else:
            setattr(L, n, lambda *a: 0)
    try:
        torch.manual_seed(0)
        m = FarSegB200(model_config(CFG)).cuda().train()
        eng = m._engine()
        x, y = synthetic(CFG, units_per_gpu(CFG, 1))
        replay, out = eng.capture_step(x.cuda(), y['cls'].cuda())
        for _ in range(3):
            replay(); eng.sgd_step(0.007)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            replay(); eng.sgd_step(0.007)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        del m, eng, replay
        torch.cuda.empty_cache()
    finally:
        for n, f in saved.items():
            setattr(L, n, f)
    return ms


if __name__ == '__main__':
    only = sys.argv[1:]
    base = None
    rows = []
    for name, fns in CLASSES.items():
        if only and name != 'none' and name not in only:
            continue
        ms = run(fns)
        if name == 'none':
            base = ms
        rows.append(dict(removed=name, ms=round(ms, 3), saves_ms=round(base - ms, 3)))
        print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open('gpurun_out/knockout.json', 'w'), indent=1)
