"""metric_summary (ever_b200/evaluate.py) against the real ever.metric.pixel.PixelMetric (build container) and against
hand-computed values."""
import os
import sys

import numpy as np
import pytest

from ever_b200.evaluate import metric_summary

HAVE_REF = os.path.isdir('/root/reference/ever')


def test_metric_summary_known_values():
    cm = np.array([[5, 1], [2, 4]])
    s = metric_summary(cm)
    assert s['oa'] == 0.75
    assert np.allclose(s['iou'], [5 / 8, 4 / 7], atol=1e-5)
    assert np.allclose(s['precision'], [5 / 7, 4 / 5], atol=1e-5) and np.allclose(s['recall'], [5 / 6, 4 / 6], atol=1e-5)
    assert abs(s['kappa'] - 0.5) < 1e-4      # po = .75, pe = .5


@pytest.mark.skipif(not HAVE_REF, reason='reference tree only exists in the build container')
def test_metric_summary_equals_reference_pixel_metric():
    for p_ in ('/root/reference', os.path.join(os.path.dirname(__file__), 'golden', '_stubs')):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    from ever.metric.pixel import PixelMetric
    rng = np.random.RandomState(0)
    for k in (2, 5, 15):
        cm = rng.randint(0, 100000, size=(k, k)).astype(np.int64)
        cm[np.arange(k), np.arange(k)] += 300000
        s = metric_summary(cm)
        f32 = cm.astype(np.float32)
        assert np.array_equal(s['iou'], np.round(PixelMetric.compute_iou_per_class(cm), 5))
        assert np.array_equal(s['f1'], np.round(PixelMetric.compute_F_measure_per_class(cm, beta=1.0), 5))
        assert np.array_equal(s['precision'], np.round(PixelMetric.compute_precision_per_class(cm), 5))
        assert np.array_equal(s['recall'], np.round(PixelMetric.compute_recall_per_class(cm), 5))
        assert s['oa'] == np.round(PixelMetric.compute_overall_accuracy(cm), 5)
        assert s['kappa'] == np.round(PixelMetric.cohen_kappa_score(f32), 5)
