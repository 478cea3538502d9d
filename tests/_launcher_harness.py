"""Harness for driving a model through the REAL reference training loop (``ever.core.launcher.Launcher.train_iters``,
ever/core/launcher.py:248-367) -- test infrastructure.

The unmodified reference is installed to ``baseline/_ref`` (git-ignored, travels to the GPU box) by
``pip install --no-index --no-deps --target baseline/_ref /root/reference``; ``prettytable`` /
``albumentations`` are satisfied by the two import stubs under tests/golden/_stubs.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.ref_glue import import_reference, make_reference_farseg, reference_available  # noqa: E402,F401


class _ListLoader:
    """minimal DataLoader stand-in: what Iterator (ever/core/iterator.py:42-75) touches is iter(), len() and .sampler"""
    sampler = None
    batch_sampler = None

    def __init__(self, items):
        self.items = list(items)

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


def run_launcher(model, batches, num_iters, model_dir, forward_times=1, mixed_precision='bf16', base_lr=0.01,
                 wrap=None):
    """Launcher.train_iters over `batches` (list of (x, dict(cls=y)) host tensors) with the reference's own optimizer /
    LR factories (SGD momentum 0.9 wd 1e-4, grad_clip max_norm 35; poly LR).  `wrap`: callable model -> wrapped model
    (e.g. DistributedDataParallel) applied AFTER the optimizer saw the parameters, as THDDPTrainer does
    (ever/trainer/th_ddp_trainer.py:25-43 builds the optimizer from model.module.custom_param_groups()).
    Returns (per-forward loss dicts as python floats, the Launcher's last logged dict)."""
    import_reference()
    from ever.core.builder import make_learningrate, make_optimizer
    from ever.core.launcher import Launcher
    inner = model
    opt = make_optimizer(dict(type='sgd', params=dict(momentum=0.9, weight_decay=1e-4, lr=base_lr),
                              grad_clip=dict(max_norm=35, norm_type=2)), params=inner.custom_param_groups())
    lr = make_learningrate(dict(type='poly', params=dict(base_lr=base_lr, power=0.9, max_iters=num_iters)))
    if wrap is not None:
        model = wrap(model)
    seen = []
    h = inner.register_forward_hook(lambda m, i, o: seen.append({k: v.detach() for k, v in o.items()})
                                    if isinstance(o, dict) else None)
    tl = Launcher(model_dir=model_dir, model=model, optimizer=opt, lr_schedule=lr, mixed_precision=mixed_precision)
    last = tl.train_iters(_ListLoader(batches), num_iters=num_iters, forward_times=forward_times, log_interval_step=1,
                          save_ckpt_interval_epoch=10 ** 6)
    h.remove()
    return [{k: float(v) for k, v in d.items()} for d in seen], last
