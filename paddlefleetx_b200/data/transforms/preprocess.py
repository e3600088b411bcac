"""Image transforms (reference data/transforms/preprocess.py:40-437): DecodeImage, ResizeImage, CenterCropImage,
RandCropImage, RandFlipImage, NormalizeImage, ToCHWImage, ColorJitter, GaussianBlur, Pixels, RandomErasing,
RandomGrayscale.  numpy HWC uint8/float in, numpy out; PIL is used when available, otherwise a pure-numpy bilinear
resize keeps the pipeline importable on minimal machines."""
from __future__ import annotations

import io
import math
import random
from typing import Optional, Sequence

import numpy as np

try:
    from PIL import Image, ImageFilter
    _HAS_PIL = True
except ImportError:  # pragma: no cover
    _HAS_PIL = False


def _number(expr) -> float:
    """YAML recipes write constants like ``1.0/255.0``; evaluate that arithmetic without ``eval``."""
    import ast
    import operator as op

    if not isinstance(expr, str):
        return float(expr)
    ops = {ast.Add: op.add, ast.Sub: op.sub, ast.Mult: op.mul, ast.Div: op.truediv, ast.USub: op.neg, ast.Pow: op.pow}

    def ev(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.BinOp) and type(node.op) in ops:
            return ops[type(node.op)](ev(node.left), ev(node.right))
        if isinstance(node, ast.UnaryOp) and type(node.op) in ops:
            return ops[type(node.op)](ev(node.operand))
        raise ValueError(f"unsupported numeric expression: {expr!r}")

    return float(ev(ast.parse(expr, mode="eval").body))


def _resize(img: np.ndarray, size, interpolation: str = "bilinear") -> np.ndarray:
    h, w = (int(size[0]), int(size[1])) if isinstance(size, (tuple, list)) else (int(size), int(size))      # size = (height, width)
    if _HAS_PIL:
        mode = {"nearest": Image.NEAREST, "bilinear": Image.BILINEAR, "bicubic": Image.BICUBIC, "lanczos": Image.LANCZOS}.get(interpolation, Image.BILINEAR)
        return np.asarray(Image.fromarray(img.astype(np.uint8) if img.dtype != np.uint8 else img).resize((w, h), mode))
    ys = (np.arange(h) + 0.5) * img.shape[0] / h - 0.5
    xs = (np.arange(w) + 0.5) * img.shape[1] / w - 0.5
    y0, x0 = np.clip(np.floor(ys).astype(int), 0, img.shape[0] - 1), np.clip(np.floor(xs).astype(int), 0, img.shape[1] - 1)
    y1, x1 = np.clip(y0 + 1, 0, img.shape[0] - 1), np.clip(x0 + 1, 0, img.shape[1] - 1)
    wy, wx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    f = img.astype(np.float32)
    out = f[y0][:, x0] * (1 - wy) * (1 - wx) + f[y0][:, x1] * (1 - wy) * wx + f[y1][:, x0] * wy * (1 - wx) + f[y1][:, x1] * wy * wx
    return out.astype(img.dtype)


class OperatorParamError(ValueError):
    """An image operator was configured with an impossible combination of parameters (reference preprocess.py:34-37)."""


class UnifiedResize:
    """``resize(src, (width, height))`` with the interpolation chosen by name, through OpenCV (``backend="cv2"``, when installed) or PIL
    (reference preprocess.py:63-104; note the OpenCV-style ``(width, height)`` size order of this call).  Unknown back ends fall back to the
    first one available; with neither library a NumPy bilinear kernel is used."""

    _CV2 = {"nearest": "INTER_NEAREST", "bilinear": "INTER_LINEAR", "area": "INTER_AREA", "bicubic": "INTER_CUBIC", "lanczos": "INTER_LANCZOS4"}

    def __init__(self, interpolation=None, backend="cv2"):
        self.interpolation = (interpolation or "bilinear").lower() if isinstance(interpolation, (str, type(None))) else interpolation
        self.backend = str(backend).lower()
        self._cv2 = None
        if self.backend == "cv2":
            try:
                import cv2

                self._cv2 = cv2
            except ImportError:
                self.backend = "pil"
        elif self.backend != "pil":
            self.backend = "pil"

    def __call__(self, src, size):
        w, h = int(size[0]), int(size[1])
        if self._cv2 is not None:
            interp = getattr(self._cv2, self._CV2.get(self.interpolation, "INTER_LINEAR")) if isinstance(self.interpolation, str) else self.interpolation
            return self._cv2.resize(src, (w, h), interpolation=interp)
        return _resize(src, (h, w), self.interpolation if isinstance(self.interpolation, str) else "bilinear")


class DecodeImage:
    def __init__(self, to_rgb: bool = True, channel_first: bool = False, **unused):
        self.to_rgb, self.channel_first = to_rgb, channel_first

    def __call__(self, img):
        if isinstance(img, (bytes, bytearray)):
            assert _HAS_PIL, "PIL is required to decode image bytes"
            img = np.asarray(Image.open(io.BytesIO(img)).convert("RGB"))
        elif isinstance(img, str):
            assert _HAS_PIL
            img = np.asarray(Image.open(img).convert("RGB"))
        img = np.asarray(img)
        if not self.to_rgb and img.ndim == 3:
            img = img[:, :, ::-1]
        return img.transpose(2, 0, 1) if self.channel_first else img


class ResizeImage:
    def __init__(self, size=None, resize_short=None, interpolation="bilinear", backend="pil", **unused):
        if not (resize_short and resize_short > 0) and size is None:
            raise OperatorParamError("invalid params for ResizeImage: both 'size' and 'resize_short' are None")
        self.size, self.resize_short, self.interp = size, resize_short, interpolation or "bilinear"

    def __call__(self, img):
        h, w = img.shape[:2]
        if self.resize_short:
            s = self.resize_short / min(h, w)
            return _resize(img, (int(round(h * s)), int(round(w * s))), self.interp)
        size = (self.size, self.size) if isinstance(self.size, int) else tuple(self.size)
        return _resize(img, size, self.interp)


class CenterCropImage:
    def __init__(self, size):
        self.size = (size, size) if isinstance(size, int) else tuple(size)

    def __call__(self, img):
        h, w = img.shape[:2]
        th, tw = self.size
        t, l = max((h - th) // 2, 0), max((w - tw) // 2, 0)
        return img[t:t + th, l:l + tw]


class RandCropImage:
    def __init__(self, size, scale=(0.08, 1.0), ratio=(3.0 / 4, 4.0 / 3), interpolation="bilinear", backend="pil", **unused):
        self.size = (size, size) if isinstance(size, int) else tuple(size)
        self.scale, self.ratio, self.interp = tuple(scale), tuple(ratio), interpolation

    def __call__(self, img):
        h, w = img.shape[:2]
        area = h * w
        for _ in range(10):
            ta = random.uniform(*self.scale) * area
            ar = math.exp(random.uniform(math.log(self.ratio[0]), math.log(self.ratio[1])))
            cw, ch = int(round(math.sqrt(ta * ar))), int(round(math.sqrt(ta / ar)))
            if 0 < cw <= w and 0 < ch <= h:
                t, l = random.randint(0, h - ch), random.randint(0, w - cw)
                return _resize(img[t:t + ch, l:l + cw], self.size, self.interp)
        return _resize(CenterCropImage(min(h, w))(img), self.size, self.interp)


class RandFlipImage:
    def __init__(self, flip_code: int = 1):
        assert flip_code in (-1, 0, 1)
        self.flip_code = flip_code

    def __call__(self, img):
        if random.random() < 0.5:
            if self.flip_code == 1:
                return img[:, ::-1]
            if self.flip_code == 0:
                return img[::-1]
            return img[::-1, ::-1]
        return img


class NormalizeImage:
    def __init__(self, scale=None, mean=None, std=None, order="chw", output_fp16=False, channel_num=3, **unused):
        self.scale = _number(scale if scale is not None else 1.0 / 255.0)
        shape = (3, 1, 1) if order == "chw" else (1, 1, 3)
        self.mean = np.asarray(mean if mean is not None else [0.485, 0.456, 0.406], np.float32).reshape(shape)
        self.std = np.asarray(std if std is not None else [0.229, 0.224, 0.225], np.float32).reshape(shape)
        self.fp16 = output_fp16

    def __call__(self, img):
        out = (np.asarray(img).astype(np.float32) * self.scale - self.mean) / self.std
        return out.astype(np.float16) if self.fp16 else out


class ToCHWImage:
    def __call__(self, img):
        return np.ascontiguousarray(np.asarray(img).transpose(2, 0, 1))


class ColorJitter:
    def __init__(self, brightness=0.0, contrast=0.0, saturation=0.0, hue=0.0, p=1.0, **unused):
        self.b, self.c, self.s, self.h, self.p = brightness, contrast, saturation, hue, p

    def __call__(self, img):
        if random.random() > self.p:
            return img
        f = img.astype(np.float32)
        ops = []
        if self.b:
            ops.append(lambda x: x * random.uniform(max(0, 1 - self.b), 1 + self.b))
        if self.c:
            ops.append(lambda x: (x - x.mean()) * random.uniform(max(0, 1 - self.c), 1 + self.c) + x.mean())
        if self.s:
            def sat(x):
                g = x.mean(-1, keepdims=True)
                return (x - g) * random.uniform(max(0, 1 - self.s), 1 + self.s) + g
            ops.append(sat)
        random.shuffle(ops)
        for op in ops:
            f = op(f)
        return np.clip(f, 0, 255).astype(img.dtype)


class GaussianBlur:
    def __init__(self, sigma=(0.1, 2.0), p=0.5, **unused):
        self.sigma, self.p = tuple(sigma), p

    def __call__(self, img):
        if random.random() > self.p:
            return img
        s = random.uniform(*self.sigma)
        if _HAS_PIL:
            return np.asarray(Image.fromarray(img.astype(np.uint8)).filter(ImageFilter.GaussianBlur(radius=s)))
        k = max(int(2 * round(3 * s) + 1), 3)
        ax = np.arange(k) - k // 2
        ker = np.exp(-0.5 * (ax / s) ** 2); ker /= ker.sum()
        f = img.astype(np.float32)
        f = np.apply_along_axis(lambda m: np.convolve(m, ker, mode="same"), 0, f)
        f = np.apply_along_axis(lambda m: np.convolve(m, ker, mode="same"), 1, f)
        return f.astype(img.dtype)


class RandomGrayscale:
    def __init__(self, p=0.2):
        self.p = p

    def __call__(self, img):
        if random.random() < self.p:
            g = (img[..., 0] * 0.299 + img[..., 1] * 0.587 + img[..., 2] * 0.114).astype(img.dtype)
            return np.stack([g, g, g], -1)
        return img


class Pixels:
    def __init__(self, mode="const", mean=(0.0, 0.0, 0.0)):
        self.mode, self.mean = mode, mean

    def __call__(self, h=224, w=224, c=3):
        if self.mode == "rand":
            return np.random.normal(size=(1, 1, 3))
        if self.mode == "pixel":
            return np.random.normal(size=(h, w, c))
        return np.asarray(self.mean, np.float32)


class RandomErasing:
    def __init__(self, EPSILON=0.5, sl=0.02, sh=0.4, r1=0.3, mean=(0.0, 0.0, 0.0), attempt=100, use_log_aspect=False, mode="const", **unused):
        self.p, self.sl, self.sh, self.r1, self.attempt, self.log = _number(EPSILON), sl, sh, r1, attempt, use_log_aspect
        self.get = Pixels(mode, mean)

    def __call__(self, img):
        if random.random() > self.p:
            return img
        h, w = img.shape[:2]
        for _ in range(self.attempt):
            ta = random.uniform(self.sl, self.sh) * h * w
            ar = math.exp(random.uniform(math.log(self.r1), math.log(1 / self.r1))) if self.log else random.uniform(self.r1, 1 / self.r1)
            eh, ew = int(round(math.sqrt(ta * ar))), int(round(math.sqrt(ta / ar)))
            if ew < w and eh < h:
                t, l = random.randint(0, h - eh), random.randint(0, w - ew)
                img = img.copy()
                img[t:t + eh, l:l + ew] = self.get(eh, ew, img.shape[2])
                return img
        return img


def build_transforms(ops: Optional[Sequence]):
    import sys

    me = sys.modules[__name__]
    out = []
    for op in ops or []:
        if isinstance(op, str):
            out.append(getattr(me, op)())
        else:
            (name, kw), = op.items()
            out.append(getattr(me, name)(**(kw or {})))
    return out


def transform(data, ops):
    for op in ops:
        data = op(data)
    return data
