"""evaluate_pixel_prediction: GPU confusion-matrix accumulation over a dataloader == the host-side accumulation the reference
does (ever/metric/confusion_matrix.py:11-25) on the same predictions; exact integer counts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_evaluate_loop_counts_equal_host_accumulation():
    from ever_b200.evaluate import evaluate_pixel_prediction, metric_summary
    from ever_b200.infer import SlidingWindowPredictor
    from _helpers import eval_r18_model as _model
    k = 5
    model = _model(k)
    g = torch.Generator().manual_seed(12)
    batches = []
    for i in range(3):
        x = torch.randn(2, 3, 96, 128, generator=g)
        y = torch.randint(0, k, (2, 96, 128), generator=g)
        y[torch.rand(2, 96, 128, generator=g) < 0.1] = 255
        batches.append((x, dict(cls=y) if i % 2 else y))
    dense, summary = evaluate_pixel_prediction(model, batches, k)
    want = np.zeros((k, k), dtype=np.int64)
    for x, y in batches:
        lab = (y['cls'] if isinstance(y, dict) else y).numpy().reshape(-1)
        _, mask = model._engine().forward_eval(x.cuda(), return_mask=True)
        pred = mask.cpu().numpy().reshape(-1)
        keep = lab < k
        np.add.at(want, (lab[keep], pred[keep]), 1)
    assert dense.dtype == np.int64 and np.array_equal(dense, want)
    assert int(dense.sum()) == sum(int(((y['cls'] if isinstance(y, dict) else y) < k).sum()) for _, y in batches)
    ref = metric_summary(want)
    assert all(np.array_equal(summary[key], ref[key]) for key in ref)
    assert not model.training
    # a custom predictor (sliding windows over each image of the batch) goes through the same accumulation
    sw = SlidingWindowPredictor(model, tile=64, stride=32, batch=4)
    dense2, _ = evaluate_pixel_prediction(model, batches[:1], k,
                                          predictor=lambda xb: torch.stack([sw(img)[1] for img in xb]))
    assert int(dense2.sum()) == int((batches[0][1] < k).sum())
