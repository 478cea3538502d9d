"""The step loop for B200 plugin models: the native counterpart of ``Launcher.train_iters``
(ever/core/launcher.py:248-367) + ``compute_loss_gradient`` (:193-200) + ``ERModule.apply_gradients`` /
``clip_grad`` (ever/interface/module.py:83-108) + ``update_training_status`` / ``_update_lr`` (launcher.py:224-237).

Per iteration the reference does: fetch a batch, ``to_device``, autocast forward, sum the ``*loss`` keys, backward (DDP
all-reduce), clip + ``optimizer.step()`` + ``zero_grad``, reduce + ``.item()`` every loss (a device->host sync per
step), set the LR for the *next* iteration.  Here: the batch is copied into the model's static buffers, one cached CUDA
graph replays forward + loss + backward, one NCCL all-reduce averages the flat gradient arena, two kernels do
clip + SGD + zero_grad, and the losses are read back only every ``log_interval_step`` iterations.

LR timing is the reference's: iteration 1 runs at ``base_lr``; after iteration k the LR is set from
``schedule(k - 1)`` (the not-yet-incremented global step), i.e. iteration k >= 2 runs at ``schedule(k - 2)``.
Any ``ever.opt`` LearningRate object (``.step(global_step, optimizer)`` setting ``param_groups[*]['lr']``,
ever/opt/learning_rate.py) or a plain ``callable(step) -> lr`` can be the schedule.
"""
import time

import torch

from . import checkpoint as ckpt


class _LRHolder:
    """optimizer stand-in for reference LearningRate objects (they only touch param_groups[*]['lr'])"""

    def __init__(self, lr):
        self.param_groups = [dict(lr=lr)]


def poly_lr(base_lr, power, max_iters):
    """PolyLearningRate without warm-up (ever/opt/learning_rate.py:109-120)"""
    return lambda step: base_lr * (1 - step / max_iters) ** power


class StepLoop:
    def __init__(self, model, lr_schedule, base_lr, momentum=0.9, weight_decay=1e-4, max_norm=35.0, rank=0, world=1,
                 log_interval_step=50, log_fn=None):
        self.model, self.schedule, self.base_lr = model, lr_schedule, float(base_lr)
        self.momentum, self.weight_decay, self.max_norm = momentum, weight_decay, max_norm
        self.rank, self.world = rank, world
        self.log_interval_step, self.log_fn = log_interval_step, log_fn
        self.global_step = 0
        self._holder = _LRHolder(self.base_lr)
        # static-shape graph replay is what this loop is for; the previous setting is put back by close()
        self._prev_cuda_graph = bool(model.config.cuda_graph)
        model.config.cuda_graph = True
        model._engine().set_distributed(rank, world)

    def close(self):
        """restore the model's ``config.cuda_graph`` as it was before this loop took the model over"""
        self.model.config.cuda_graph = self._prev_cuda_graph

    @property
    def lr(self):
        return self._holder.param_groups[0]['lr']

    def _update_lr(self):
        """launcher.py:228-237: called with the NOT yet incremented global step"""
        if hasattr(self.schedule, 'step'):
            self.schedule.step(self.global_step, self._holder)
        else:
            self._holder.param_groups[0]['lr'] = float(self.schedule(self.global_step))

    def train_iters(self, batches, num_iters):
        """batches: iterator of (x, y) with x a float NCHW / uint8 NHWC tensor (host-pinned or device) and y the label
        tensor or dict.  Returns the last logged loss dict (python floats), like Launcher.train_iters."""
        model = self.model
        eng = model._engine()
        model.train()
        last, t0, lr_used = {}, time.time(), []
        while self.global_step < num_iters:
            x, y = next(batches)
            out = model(x, y)                                   # graph replay of forward + loss + backward
            model.backward(out, None, None)                     # gradient all-reduce (world > 1)
            grad_norm = eng.sgd_step(self.lr, self.momentum, self.weight_decay, self.max_norm)
            lr_used.append(self.lr)
            self._update_lr()
            self.global_step += 1
            if self.global_step % self.log_interval_step == 0 or self.global_step == num_iters:
                vec = torch.stack([v.detach().float().reshape(()) for v in out.values()] + [grad_norm[0]])
                if self.world > 1:   # reduce_loss_dict (ever/core/launcher.py:202-209): the logged losses are rank means
                    import torch.distributed as dist
                    nl = len(out)
                    red = vec[:nl].clone()
                    dist.all_reduce(red)
                    vec = torch.cat([red / self.world, vec[nl:]])
                vals = vec.cpu().tolist()
                last = dict(zip(list(out.keys()) + ['grad_norm'], vals))
                last['total_loss'] = sum(v for k, v in last.items() if k.endswith('loss'))
                if self.log_fn is not None and self.rank == 0:
                    self.log_fn(self.global_step, last, self.lr, time.time() - t0)
        self.lr_used = lr_used
        return last

    # ------------------------------------------------------------------ checkpoints in the reference's format
    def save_checkpoint(self, model_dir, filename=None):
        """CheckPoint.save (ever/core/checkpoint.py:51-73): model state_dict, global step and the optimizer state in
        torch.optim.SGD's own layout (the flat momentum arena is split back into per-parameter momentum buffers), so the
        reference Launcher can resume from it."""
        if self.rank != 0:
            return None
        eng = self.model._engine()
        params = list(self.model.parameters())
        mom = getattr(eng, '_mom', None)
        opt_state = ckpt.sgd_state_from_flat(params, mom, self.lr, self.momentum, self.weight_decay,
                                             has_momentum=mom is not None and not getattr(eng, '_sgd_first', True))
        state = {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}
        return ckpt.write_checkpoint(model_dir, state, opt_state, self.global_step, filename)

    def try_resume(self, model_dir):
        """CheckPoint.try_resume (checkpoint.py:79-107): load the checkpoint named under 'last' in checkpoint_info.json --
        written by this class or by the reference Launcher with torch.optim.SGD -- into the model, the momentum arena and
        the step counter.  Returns True if a checkpoint was loaded."""
        c = ckpt.read_last_checkpoint(model_dir)
        if c is None:
            return False
        self.model.load_state_dict(c[ckpt.MODEL])
        eng = self.model._engine()
        eng.ensure_optimizer_state()
        has, group = ckpt.flat_from_sgd_state(list(self.model.parameters()), c[ckpt.OPTIMIZER], eng._mom)
        eng._sgd_first = not has
        self.global_step = int(c[ckpt.GLOBALSTEP])
        self._holder.param_groups[0]['lr'] = float(group['lr'])
        return True
