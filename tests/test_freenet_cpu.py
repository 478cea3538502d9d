"""CPU-side checks of the FreeNet plugin model: parameter layout against the restated oracle, the oracle's squeeze-excitation
block against the REAL ever.module.se_block.SEBlock (the one piece of FreeNet that is in the reference tree), registry."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('cin,k', [(103, 9), (200, 16)])
def test_freenet_state_dict_contract(cin, k):
    from ever_b200.freenet import FreeNetB200
    from oracle.freenet_oracle import FreeNetOracle
    m, o = FreeNetB200(dict(in_channels=cin, num_classes=k)), FreeNetOracle(cin, k)
    a = [(n, tuple(v.shape), v.dtype) for n, v in m.state_dict().items()]
    b = [(n, tuple(v.shape), v.dtype) for n, v in o.state_dict().items()]
    assert a == b and len(a) == 60
    m.load_state_dict(o.state_dict(), strict=True)
    # squeeze-excitation keys are those of ever.module.se_block.SEBlock (se_block.py:13-18)
    assert 'feature_ops.1.0.0.seq.0.weight' in m.state_dict() and 'feature_ops.1.0.0.seq.2.bias' in m.state_dict()


def test_freenet_registered_and_refuses_cpu():
    from ever_b200._ever_api import MODEL, ERModule
    from ever_b200.freenet import FreeNetB200
    assert MODEL['FreeNetB200'] is FreeNetB200 and MODEL['FreeNet'] is FreeNetB200 and issubclass(FreeNetB200, ERModule)
    m = FreeNetB200(dict(in_channels=32, num_classes=4)).train()
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(torch.zeros(1, 32, 16, 16), torch.ones(1, 16, 16, dtype=torch.long), torch.ones(1, 16, 16))


@pytest.mark.skipif(not (os.path.isdir('/root/reference/ever') or os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'ever'))),
                    reason='reference package not available')
def test_seblock_oracle_bit_exact_vs_reference():
    from oracle.freenet_oracle import FreeNetOracle, SEBlockOracle
    from oracle.ref_glue import import_reference
    import_reference()
    from ever.module.se_block import SEBlock
    torch.manual_seed(0)
    ref, ora = SEBlock(96, 16), SEBlockOracle(96, 16)
    ora.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 96, 20, 12, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    yr, yo = ref(x), ora(x2)
    assert torch.equal(yr, yo)
    yr.square().sum().backward()
    yo.square().sum().backward()
    assert torch.equal(x.grad, x2.grad)
    for (ka, pa), (kb, pb) in zip(ref.named_parameters(), ora.named_parameters()):
        assert ka == kb and torch.equal(pa.grad, pb.grad)
    # the whole restated network built on the REAL block gives the same loss and gradients as on the restated block
    a, b = FreeNetOracle(24, 5, se_cls=SEBlock), FreeNetOracle(24, 5)
    b.load_state_dict(a.state_dict(), strict=True)
    from oracle.freenet_oracle import synthetic_cube
    xx, yy, ww = synthetic_cube(1, 24, 32, 24, 5, labelled_frac=0.3)
    la, lb = a.train()(xx, yy, ww)['cls_loss'], b.train()(xx, yy, ww)['cls_loss']
    assert torch.equal(la, lb)
