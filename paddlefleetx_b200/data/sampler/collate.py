"""Collate helpers: ``Stack`` / ``Pad`` / ``Tuple`` / ``Dict`` combinators and the task-level collate functions
(reference ppfleetx/data/sampler/collate.py:27-317): ``Stack`` / ``Pad`` / ``Tuple`` / ``Dict``."""
from __future__ import annotations

from typing import Callable, List
from typing import Dict as _Dict          # the combinator below is called ``Dict`` like the reference's

import numpy as np
import torch


class Stack:
    def __init__(self, axis: int = 0, dtype=None):
        self.axis, self.dtype = axis, dtype

    def __call__(self, data):
        out = np.stack([np.asarray(d) for d in data], axis=self.axis)
        return out.astype(self.dtype) if self.dtype else out


class Pad:
    def __init__(self, pad_val=0, axis: int = 0, ret_length=None, dtype=None, pad_right: bool = True):
        self.pad_val, self.axis, self.ret_length, self.dtype, self.pad_right = pad_val, axis, ret_length, dtype, pad_right

    def __call__(self, data):
        arrs = [np.asarray(d) for d in data]
        lens = [a.shape[self.axis] for a in arrs]
        mx = max(lens)
        shape = list(arrs[0].shape)
        shape[self.axis] = mx
        out = np.full([len(arrs)] + shape, self.pad_val, dtype=self.dtype or arrs[0].dtype)
        for i, a in enumerate(arrs):
            sl = [slice(None)] * a.ndim
            n = a.shape[self.axis]
            sl[self.axis] = slice(0, n) if self.pad_right else slice(mx - n, mx)
            out[(i, *sl)] = a
        if self.ret_length:
            return out, np.asarray(lens, dtype="int32" if self.ret_length is True else self.ret_length)
        return out


class Tuple:
    def __init__(self, fn, *args):
        self.fns = list(fn) if isinstance(fn, (list, tuple)) else [fn, *args]

    def __call__(self, data):
        assert len(data[0]) == len(self.fns), f"sample has {len(data[0])} fields, Tuple has {len(self.fns)} functions"
        out = []
        for i, f in enumerate(self.fns):
            r = f([sample[i] for sample in data])
            out.extend(r) if isinstance(r, tuple) else out.append(r)
        return tuple(out)


class Dict:
    def __init__(self, fn: _Dict[str, Callable]):
        self.fns = dict(fn)

    def __call__(self, data):
        out = []
        for k, f in self.fns.items():
            r = f([sample[k] for sample in data])
            out.extend(r) if isinstance(r, tuple) else out.append(r)
        return tuple(out)
