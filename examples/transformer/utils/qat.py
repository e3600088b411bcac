"""Quantisation-aware-training hook of the engine-less examples (reference examples/transformer/utils/qat.py:21-60 drives paddleslim)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from paddlefleetx_b200.utils import compression_helper  # noqa: E402


def compress_model(config, model, input_spec=None):
    """Apply the ``Compress`` section (pruning first, then fake-quant wrapping) to ``model``.

    Returns ``(model, quanter)``: ``quanter(model)`` folds the observers into int8 weights at export time (``None`` when
    quantisation is off) — the pair the reference's loop keeps around.  ``input_spec`` is accepted for signature parity; nothing is traced."""
    del input_spec
    compress = config.get("Compress", None)
    if not compress:
        return model, None
    model, quantised = compression_helper.compress_model(model, compress)
    return model, (compression_helper.convert_to_int8 if quantised else None)
