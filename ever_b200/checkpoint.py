"""Checkpoint bridge: the fused clip+SGD step keeps its momentum in one flat fp32 arena; this module
converts it to / from the ``torch.optim.SGD.state_dict()`` layout and reads / writes checkpoints in the reference's own
format, so a run can move between ``ever_b200.trainer.StepLoop`` and the reference ``Launcher`` in either direction.

Reference format (ever/core/checkpoint.py:51-117): ``torch.save(OrderedDict(model=state_dict, global_step=int,
opt=optimizer.state_dict()), model_dir/'checkpoint-<step>.pth')`` plus ``checkpoint_info.json`` =
``{'last': {'step': s, 'name': file}, '<step>': file, ...}``; ``try_resume`` loads the file named under 'last'.
"""
import json
import os
from collections import OrderedDict

import torch

MODEL, OPTIMIZER, GLOBALSTEP, LAST, INFO_NAME = 'model', 'opt', 'global_step', 'last', 'checkpoint_info.json'


def checkpoint_name(global_step):
    return 'checkpoint-{}.pth'.format(global_step)   # CheckPoint.get_checkpoint_name, checkpoint.py:139-141


def param_slots(params):
    """(offset, numel) of every parameter in the flat arenas: slots are 16-byte aligned (engine._flatten_params)"""
    out, off = [], 0
    for p in params:
        out.append((off, p.numel()))
        off += (p.numel() + 3) // 4 * 4
    return out, off


def sgd_state_from_flat(params, mom_flat, lr, momentum, weight_decay, has_momentum=True):
    """flat momentum arena -> torch.optim.SGD.state_dict() (one param group over ``params`` in order)"""
    params = list(params)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.empty(0)) for _ in params], lr=float(lr), momentum=float(momentum),
                          weight_decay=float(weight_decay))
    sd = opt.state_dict()      # the exact key set of this torch version (foreach, fused, maximize, ...)
    slots, _ = param_slots(params)
    state = {}
    if has_momentum and mom_flat is not None:
        for i, (p, (off, n)) in enumerate(zip(params, slots)):
            if p.requires_grad:
                state[i] = dict(momentum_buffer=mom_flat[off:off + n].detach().reshape(p.shape).clone().cpu())
    sd['state'] = state
    return sd


def flat_from_sgd_state(params, opt_state, mom_flat):
    """torch.optim.SGD.state_dict() -> flat momentum arena (in place); returns (has_momentum, param_group_0)"""
    params = list(params)
    group = opt_state['param_groups'][0]
    if len(opt_state['param_groups']) != 1:
        raise ValueError('the fused SGD step keeps one parameter group (the reference default, '
                         'ERModule.custom_param_groups)')
    order = list(group['params'])
    trainable = [i for i, p in enumerate(params) if p.requires_grad]
    if len(order) not in (len(params), len(trainable)):
        raise ValueError('optimizer state covers %d parameters, the model has %d (%d trainable)'
                         % (len(order), len(params), len(trainable)))
    index_of = list(range(len(params))) if len(order) == len(params) else trainable
    slots, _ = param_slots(params)
    mom_flat.zero_()
    any_buf = False
    for key, pi in zip(order, index_of):
        st = opt_state['state'].get(key)
        if st is None or st.get('momentum_buffer') is None:
            continue
        off, n = slots[pi]
        mom_flat[off:off + n].copy_(st['momentum_buffer'].reshape(-1).to(mom_flat.device, mom_flat.dtype))
        any_buf = True
    return any_buf, group


def write_checkpoint(model_dir, model_state, opt_state, global_step, filename=None):
    """CheckPoint.save (checkpoint.py:51-73): the .pth file and checkpoint_info.json"""
    os.makedirs(model_dir, exist_ok=True)
    filename = filename or checkpoint_name(global_step)
    ckpt = OrderedDict([(MODEL, model_state), (GLOBALSTEP, int(global_step)), (OPTIMIZER, opt_state)])
    torch.save(ckpt, os.path.join(model_dir, filename))
    info_path = os.path.join(model_dir, INFO_NAME)
    info = {LAST: dict(step=0, name='')}
    if os.path.exists(info_path):
        with open(info_path) as f:
            info = json.load(f)
    info[str(global_step)] = filename
    if global_step > info[LAST]['step']:
        info[LAST] = dict(step=int(global_step), name=filename)
    with open(info_path, 'w') as f:
        json.dump(info, f)
    return os.path.join(model_dir, filename)


def read_last_checkpoint(model_dir):
    """CheckPoint.try_resume steps 1-3 (checkpoint.py:86-99): json -> path -> checkpoint dict (None if there is none)"""
    info_path = os.path.join(model_dir, INFO_NAME)
    if not os.path.exists(info_path):
        return None
    with open(info_path) as f:
        info = json.load(f)
    name = info[LAST]['name']
    if not name:
        return None
    return torch.load(os.path.join(model_dir, name), map_location='cpu', weights_only=False)
