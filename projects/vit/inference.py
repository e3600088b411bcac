"""Serve an exported ViT classifier (reference projects/vit/inference.py): decode + resize + centre-crop + normalise an image
(or a random tensor with --random), run the ``InferenceEngine`` and print the top-5 classes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import numpy as np  # noqa: E402

from paddlefleetx_b200.core.engine.inference_engine import InferenceEngine  # noqa: E402


def load_image(path, size):
    from paddlefleetx_b200.data.transforms import preprocess as T

    ops = [T.DecodeImage(to_rgb=True), T.ResizeImage(resize_short=int(size * 256 / 224)), T.CenterCropImage(size=size),
           T.NormalizeImage(scale=1.0 / 255.0, mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5], order=""), T.ToCHWImage()]
    with open(path, "rb") as f:
        img = f.read()
    for op in ops:
        img = op(img)
    return np.asarray(img, dtype=np.float32)[None]


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model_dir", default="./output")
    p.add_argument("--image", default=None)
    p.add_argument("--size", type=int, default=224)
    p.add_argument("--random", action="store_true")
    a = p.parse_args(argv)
    engine = InferenceEngine(a.model_dir, 1)
    x = np.random.RandomState(0).randn(1, 3, a.size, a.size).astype(np.float32) if (a.random or not a.image) else load_image(a.image, a.size)
    logits = next(iter(engine.predict([x]).values()))
    top5 = np.argsort(-logits[0])[:5]
    print("top-5 classes:", top5.tolist(), "scores:", logits[0][top5].round(3).tolist())
    return logits


if __name__ == "__main__":
    main()
