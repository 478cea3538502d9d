"""Checkpoint bridge (ever_b200/checkpoint.py): flat momentum arena <-> torch.optim.SGD.state_dict(), and the reference's
checkpoint files in both directions (the real ever.core.checkpoint.CheckPoint when /root/reference is present)."""
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

from ever_b200 import checkpoint as ckpt

HAVE_REF = os.path.isdir('/root/reference/ever')


def _toy():
    torch.manual_seed(0)
    m = nn.Sequential(nn.Conv2d(3, 5, 3), nn.BatchNorm2d(5), nn.Conv2d(5, 7, 1))
    m[2].bias.requires_grad = False   # a frozen parameter: torch SGD keeps no state for it
    return m


def _step(m, opt, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 3, 8, 8, generator=g)
    opt.zero_grad()
    m(x).square().mean().backward()
    opt.step()


def _flat(params):
    slots, total = ckpt.param_slots(params)
    return torch.zeros(total)


def test_flat_arena_round_trip_continues_training_identically():
    a, b = _toy(), _toy()
    oa = torch.optim.SGD(a.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    for s in range(5):
        _step(a, oa, s)
    ob = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    for s in range(3):
        _step(b, ob, s)
    # optimizer state -> flat arena -> optimizer state of a fresh SGD
    params = list(b.parameters())
    flat = _flat(params)
    has, group = ckpt.flat_from_sgd_state(params, ob.state_dict(), flat)
    assert has and group['momentum'] == 0.9
    slots, _ = ckpt.param_slots(params)
    assert all(off % 4 == 0 for off, _ in slots)
    sd = ckpt.sgd_state_from_flat(params, flat, 0.05, 0.9, 1e-4)
    assert set(sd['param_groups'][0]) == set(ob.state_dict()['param_groups'][0])
    assert sorted(sd['state']) == sorted(ob.state_dict()['state'])   # no entry for the frozen bias
    for k, v in ob.state_dict()['state'].items():
        assert torch.equal(sd['state'][k]['momentum_buffer'], v['momentum_buffer'])
    oc = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    oc.load_state_dict(sd)
    for s in range(3, 5):
        _step(b, oc, s)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)


def test_checkpoint_files_follow_reference_layout(tmp_path):
    m = _toy()
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    _step(m, opt, 0)
    d = str(tmp_path)
    ckpt.write_checkpoint(d, m.state_dict(), opt.state_dict(), 7)
    ckpt.write_checkpoint(d, m.state_dict(), opt.state_dict(), 3)      # an older step never becomes 'last'
    import json
    info = json.load(open(os.path.join(d, 'checkpoint_info.json')))
    assert info['last'] == dict(step=7, name='checkpoint-7.pth') and info['3'] == 'checkpoint-3.pth'
    c = ckpt.read_last_checkpoint(d)
    assert list(c.keys()) == ['model', 'global_step', 'opt'] and c['global_step'] == 7
    assert ckpt.read_last_checkpoint(str(tmp_path / 'nothing')) is None


@pytest.mark.skipif(not HAVE_REF, reason='reference tree only exists in the build container')
def test_interchange_with_real_reference_checkpoint(tmp_path):
    for p_ in ('/root/reference', os.path.join(os.path.dirname(__file__), 'golden', '_stubs')):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    from ever.core.checkpoint import CheckPoint
    import logging
    d = str(tmp_path)
    m = _toy()
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    _step(m, opt, 0)
    launcher = types.SimpleNamespace(model_dir=d, unwrapped_model=m, optimizer=opt, logger=logging.getLogger('t'),
                                     checkpoint=None)
    # reference -> bridge
    cp = CheckPoint(launcher)
    cp.set_global_step(11)
    cp.save()
    c = ckpt.read_last_checkpoint(d)
    assert c['global_step'] == 11
    params = list(m.parameters())
    flat = _flat(params)
    has, _ = ckpt.flat_from_sgd_state(params, c['opt'], flat)
    assert has
    # bridge -> reference: a file written by the bridge resumes the real Launcher pieces
    sd = ckpt.sgd_state_from_flat(params, flat, 0.1, 0.9, 1e-4)
    ckpt.write_checkpoint(d, {k: v.clone() for k, v in m.state_dict().items()}, sd, 12)
    m2 = _toy()
    opt2 = torch.optim.SGD(m2.parameters(), lr=0.5, momentum=0.9, weight_decay=1e-4)
    launcher2 = types.SimpleNamespace(model_dir=d, unwrapped_model=m2, optimizer=opt2, logger=logging.getLogger('t'),
                                      checkpoint=None)
    cp2 = CheckPoint(launcher2)
    launcher2.checkpoint = cp2
    cp2.try_resume()
    assert cp2.global_step == 12
    for pa, pb in zip(m.parameters(), m2.parameters()):
        assert torch.equal(pa, pb)
    for k, v in opt.state_dict()['state'].items():
        assert torch.equal(opt2.state_dict()['state'][k]['momentum_buffer'], v['momentum_buffer'])
    assert opt2.param_groups[0]['lr'] == 0.1
