"""Evaluation loop for the B200 plugin models: the counterpart of
``evaluate_pixel_prediction_task`` (ever/metric/evaluate_fn.py:22-59) + ``ConfusionMatrix.forward``
(ever/metric/confusion_matrix.py:11-25).

The reference moves every prediction to the host and accumulates a scipy COO matrix per batch.  Here the argmax masks stay on
the GPU, ``evb_confusion_matrix`` accumulates int64 counts there, and ONE [K, K] matrix crosses to the host at the end; it is
handed to the reference's own ``PixelMetric.summary_all(dense_cm=...)`` (ever/metric/pixel.py:163-205, left as-is) when
``ever`` is importable, otherwise to ``metric_summary`` below, which restates the same formulas.
"""
import numpy as np
import torch

EPS = 1e-7   # ever/metric/pixel.py:12


def metric_summary(dense_cm, dec=5):
    """per-class IoU / F1 / precision / recall, their means, overall accuracy and Cohen's kappa, rounded to `dec` places:
    the numbers PixelMetric.summary_all prints (ever/metric/pixel.py:68-125,163-180).  dense_cm: [K, K], row = ground truth."""
    cm = np.asarray(dense_cm)
    diag = np.diag(cm)
    pred_tot, true_tot = cm.sum(axis=0), cm.sum(axis=1)
    iou = np.round(diag / (pred_tot + true_tot - diag + EPS), dec)
    precision_raw = diag / (pred_tot + EPS)
    recall_raw = diag / (true_tot + EPS)
    f1 = np.round(2.0 * precision_raw * recall_raw / (precision_raw + recall_raw + EPS), dec)
    precision, recall = np.round(precision_raw, dec), np.round(recall_raw, dec)
    oa = np.round(diag.sum() / (cm.sum() + EPS), dec)
    c32 = cm.astype(np.float32)
    s0, s1 = c32.sum(axis=0), c32.sum(axis=1)
    expected = np.outer(s0, s1) / (np.sum(s0) + EPS)
    w = np.ones_like(expected, dtype=np.float64)
    np.fill_diagonal(w, 0)
    kappa = np.round(1.0 - np.sum(w * c32) / (np.sum(w * expected) + EPS), dec)
    return dict(iou=iou, f1=f1, precision=precision, recall=recall, miou=np.round(iou.mean(), dec),
                mf1=np.round(f1.mean(), dec), mprecision=np.round(precision.mean(), dec),
                mrecall=np.round(recall.mean(), dec), oa=oa, kappa=kappa)


@torch.no_grad()
def evaluate_pixel_prediction(model, dataloader, num_classes, predictor=None, pixel_metric=None):
    """Run `model` (eval mode) over `dataloader` items ``(x, y)`` (y: label tensor or dict with 'cls'; labels outside
    [0, K), e.g. 255, are ignored) and return ``(dense_cm int64 [K, K] on the host, summary)``.

    predictor: optional callable ``x -> uint8 mask`` (e.g. a SlidingWindowPredictor wrapper or a TTA closure); default is the
    engine's eval forward with the argmax kernel.  pixel_metric: a reference ``PixelMetric`` instance -> its own
    ``summary_all(dense_cm=...)`` table is returned as the summary; otherwise the ``metric_summary`` dict."""
    was_training = model.training
    model.eval()
    eng = model._engine()
    dev = next(model.parameters()).device
    cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=dev)
    for x, y in dataloader:
        labels = y['cls'] if isinstance(y, dict) else y
        x = x.to(dev, non_blocking=True)
        labels = labels.to(dev, non_blocking=True)
        if predictor is not None:
            mask = predictor(x)
        else:
            res = eng.forward_eval(x, return_mask=True)
            # FarSeg: (prob, mask); ChangeStar: dict with 'seg_mask' / 'change_mask' -- the segmentation mask of t1 is scored
            # against y['cls'] here; score the change head with predictor=lambda x: eng.forward_eval(x)['change_mask']
            mask = res['seg_mask'] if isinstance(res, dict) else res[1]
        eng.confusion_matrix(mask.contiguous(), labels, cm)
    dense = cm.cpu().numpy()      # the only device -> host transfer of the evaluation
    if was_training:
        model.train()
    if pixel_metric is not None:
        return dense, pixel_metric.summary_all(dense_cm=dense.astype(np.float32))
    return dense, metric_summary(dense)
