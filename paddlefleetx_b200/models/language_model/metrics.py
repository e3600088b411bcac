"""Classification / regression metrics for the GLUE fine-tune tasks (reference language_model/metrics.py:31-692):
``Accuracy``, ``AccuracyAndF1``, ``Mcc``, ``PearsonAndSpearman``, ``MultiLabelsMetric``.  Streaming interface:
``compute(pred, label)`` -> stats, ``update(stats)``, ``accumulate()``, ``reset()``."""
from __future__ import annotations

import math

import numpy as np
import torch


def _np(x):
    return x.detach().float().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class Metric:
    def reset(self): raise NotImplementedError
    def compute(self, pred, label, *a): return pred, label
    def update(self, *a): raise NotImplementedError
    def accumulate(self): raise NotImplementedError
    def name(self): return self.__class__.__name__.lower()


class Accuracy(Metric):
    def __init__(self, topk=(1,), name="acc"):
        self.topk, self._name = tuple(topk), name
        self.reset()

    def reset(self):
        self.correct = [0] * len(self.topk); self.total = 0

    def compute(self, pred, label, *a):
        pred, label = _np(pred), _np(label).reshape(-1)
        order = np.argsort(-pred, axis=-1)[:, :max(self.topk)]
        return (order == label[:, None])

    def update(self, correct):
        correct = _np(correct)
        self.total += correct.shape[0]
        for i, k in enumerate(self.topk):
            self.correct[i] += int(correct[:, :k].any(-1).sum())
        return self.accumulate()

    def accumulate(self):
        res = [c / max(self.total, 1) for c in self.correct]
        return res[0] if len(res) == 1 else res


class AccuracyAndF1(Metric):
    def __init__(self, topk=(1,), pos_label=1, name="acc_and_f1"):
        self.acc, self.pos_label = Accuracy(topk), pos_label
        self.reset()

    def reset(self):
        self.acc.reset(); self.tp = self.fp = self.fn = 0

    def compute(self, pred, label, *a):
        self._pred, self._label = _np(pred), _np(label).reshape(-1)
        return self.acc.compute(pred, label)

    def update(self, correct):
        self.acc.update(correct)
        p = self._pred.argmax(-1)
        pos = self.pos_label
        self.tp += int(((p == pos) & (self._label == pos)).sum())
        self.fp += int(((p == pos) & (self._label != pos)).sum())
        self.fn += int(((p != pos) & (self._label == pos)).sum())

    def accumulate(self):
        acc = self.acc.accumulate()
        prec = self.tp / (self.tp + self.fp) if self.tp + self.fp else 0.0
        rec = self.tp / (self.tp + self.fn) if self.tp + self.fn else 0.0
        f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
        return acc, prec, rec, f1, (acc + f1) / 2


class Mcc(Metric):
    def __init__(self, name="mcc"):
        self.reset()

    def reset(self):
        self.tp = self.fp = self.tn = self.fn = 0

    def compute(self, pred, label, *a):
        return _np(pred).argmax(-1), _np(label).reshape(-1)

    def update(self, pl):
        p, l = pl
        self.tp += int(((p == 1) & (l == 1)).sum()); self.fp += int(((p == 1) & (l == 0)).sum())
        self.tn += int(((p == 0) & (l == 0)).sum()); self.fn += int(((p == 0) & (l == 1)).sum())

    def accumulate(self):
        den = math.sqrt((self.tp + self.fp) * (self.tp + self.fn) * (self.tn + self.fp) * (self.tn + self.fn))
        return ((self.tp * self.tn - self.fp * self.fn) / den if den else 0.0,)


class PearsonAndSpearman(Metric):
    def __init__(self, name="pearson_and_spearman"):
        self.reset()

    def reset(self):
        self.preds, self.labels = [], []

    def compute(self, pred, label, *a):
        return _np(pred).reshape(-1), _np(label).reshape(-1)

    def update(self, pl):
        self.preds.append(pl[0]); self.labels.append(pl[1])

    @staticmethod
    def _pearson(a, b):
        a, b = a - a.mean(), b - b.mean()
        den = math.sqrt(float((a * a).sum() * (b * b).sum()))
        return float((a * b).sum()) / den if den else 0.0

    @staticmethod
    def _rank(x):
        order = np.argsort(x, kind="mergesort")
        ranks = np.empty(len(x), dtype=np.float64)
        sx = x[order]
        i = 0
        while i < len(x):
            j = i
            while j + 1 < len(x) and sx[j + 1] == sx[i]:
                j += 1
            ranks[order[i:j + 1]] = (i + j) / 2.0 + 1
            i = j + 1
        return ranks

    def accumulate(self):
        p, l = np.concatenate(self.preds), np.concatenate(self.labels)
        pe = self._pearson(p.astype(np.float64), l.astype(np.float64))
        sp = self._pearson(self._rank(p), self._rank(l))
        return pe, sp, (pe + sp) / 2


class MultiLabelsMetric(Metric):
    def __init__(self, num_labels: int, name="multi_labels_metric"):
        self.num_labels = num_labels
        self.reset()

    def reset(self):
        self.conf = np.zeros((self.num_labels, self.num_labels), dtype=np.int64)

    def compute(self, pred, label, *a):
        return _np(pred).argmax(-1).reshape(-1), _np(label).reshape(-1).astype(np.int64)

    def update(self, pl):
        p, l = pl
        np.add.at(self.conf, (l, p), 1)

    def accumulate(self, average="micro", pos_label=1):
        tp = np.diag(self.conf).astype(np.float64)
        fp, fn = self.conf.sum(0) - tp, self.conf.sum(1) - tp
        if average == "micro":
            prec, rec = tp.sum() / max(tp.sum() + fp.sum(), 1), tp.sum() / max(tp.sum() + fn.sum(), 1)
        elif average == "macro":
            prec, rec = np.mean(tp / np.maximum(tp + fp, 1)), np.mean(tp / np.maximum(tp + fn, 1))
        else:
            prec, rec = tp[pos_label] / max(tp[pos_label] + fp[pos_label], 1), tp[pos_label] / max(tp[pos_label] + fn[pos_label], 1)
        f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
        return float(prec), float(rec), float(f1)
