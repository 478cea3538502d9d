"""Shared helpers of the GPU tests (tests/ is on sys.path under pytest's default import mode)."""
import torch


def eval_r18_model(k=5, calibrate_hw=(128, 128)):
    """eval-mode R18 FarSegB200 with BN running statistics calibrated on a random batch (well-scaled outputs)"""
    from ever_b200.module import FarSegB200
    from oracle.farseg_oracle import FarSegOracle, deterministic_fill, synthetic_batch
    ora = deterministic_fill(FarSegOracle('resnet18', k, 128), 0).cuda().train()
    for m_ in ora.modules():
        if isinstance(m_, torch.nn.BatchNorm2d):
            m_.momentum = 1.0
    x, _ = synthetic_batch(2, calibrate_hw[0], calibrate_hw[1], k)
    with torch.no_grad():
        ora.logits(x.cuda())
    mine = FarSegB200(dict(encoder=dict(resnet_type='resnet18'),
                           head=dict(fpn_decoder=dict(out_channels=128, classifier_config=dict(num_classes=k)))))
    mine.load_state_dict(ora.state_dict(), strict=True)
    return mine.cuda().eval()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


class RefCapture:
    """Records, for one forward + backward of a reference-structured model (the oracle or the real ``ever`` modules),
    the output and output-gradient of every Conv2d / ReLU / MaxPool2d / Identity / up-sampling wrapper / down-sample BN
    (key = module path; a block's shared ``relu`` gets '#k' per call) and the input and input-gradient of every Conv2d
    (key = path + ':in').  Values are cloned at hook time (later in-place ReLUs / adds do not alter them)."""

    def __init__(self, model):
        import torch.nn as nn
        self.fwd, self.bwd, self.handles, self._calls = {}, {}, [], {}
        for name, mod in model.named_modules():
            leafish = isinstance(mod, (nn.Conv2d, nn.ReLU, nn.MaxPool2d, nn.Identity, nn.Dropout2d, nn.Dropout)) or \
                type(mod).__name__ in ('_Fp32Around', 'Bf16compatible', 'SEBlock', 'SEBlockOracle') or \
                (isinstance(mod, nn.modules.batchnorm._BatchNorm) and '.downsample.' in name)
            if not leafish:
                continue
            self.handles.append(mod.register_forward_hook(self._out_hook(name)))
            if isinstance(mod, nn.Conv2d):
                self.handles.append(mod.register_forward_pre_hook(self._in_hook(name + ':in')))

    def _key(self, name):
        if name.endswith('.relu'):
            k = self._calls.get(name, 0)
            self._calls[name] = k + 1
            return '%s#%d' % (name, k)
        return name

    def _record(self, key, t):
        self.fwd[key] = t.detach().clone()
        if t.requires_grad:
            t.register_hook(lambda g, key=key: self.bwd.__setitem__(key, g.detach().clone()))

    def _out_hook(self, name):
        def hook(mod, inp, out):
            self._record(self._key(name), out)
        return hook

    def _in_hook(self, key):
        def hook(mod, inp):
            self._record(key, inp[0])
        return hook

    def remove(self):
        for h in self.handles:
            h.remove()


class TeacherForcing:
    """engine.tf callable: compares each engine tensor (NHWC, channel-padded) with the reference tensor of the same
    module path (NCHW) and -- when ``force`` -- overwrites the engine's with the reference's, so the next op runs on the
    reference's own inputs."""

    def __init__(self, cap, force=True):
        self.cap, self.force = cap, force
        self.err = dict(fwd={}, bwd={})
        self.missing = []

    def __call__(self, kind, name, t):
        ref = (self.cap.fwd if kind == 'fwd' else self.cap.bwd).get(name)
        if ref is None:
            self.missing.append((kind, name))
            return
        if ref.dim() == 4 and t.dim() == 4:
            r = ref.permute(0, 2, 3, 1)
            view = t[..., :r.shape[-1]]
        else:
            r = ref.reshape(t.shape)
            view = t
        assert view.shape == r.shape, (kind, name, tuple(view.shape), tuple(r.shape))
        self.err[kind][name] = rel_l2(view.float(), r.float())
        if self.force:
            view.copy_(r)
