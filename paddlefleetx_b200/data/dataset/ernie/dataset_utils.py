"""ERNIE dataset helpers under the reference's module path (ppfleetx/data/dataset/ernie/dataset_utils.py:32-700): memory-mapped corpus
reader, sentence-pair sampling, span / whole-word masking and padding.  Implementations are in ``ernie_dataset.py``."""
from .ernie_dataset import (MMapIndexedDataset, create_masked_lm_predictions, create_tokens_and_tokentypes, get_a_and_b_segments,  # noqa: F401
                            get_samples_mapping, pad_and_convert_to_numpy, truncate_segments)


def make_indexed_dataset(data_prefix, data_impl=None, skip_warmup=False):
    """``<prefix>_ids.npy`` + ``<prefix>_idx.npz`` -> ``MMapIndexedDataset`` (``data_impl`` / ``skip_warmup`` kept for call-site parity)."""
    return MMapIndexedDataset(data_prefix)


def is_start_piece(piece: str) -> bool:
    """WordPiece continuation pieces start with ``##``; everything else starts a word (whole-word masking groups on this)."""
    return not piece.startswith("##")


def get_train_valid_test_split_(splits, size: int):
    """``"949,50,1"`` or ``[949, 50, 1]`` -> four cumulative document boundaries ``[0, a, b, size]``."""
    parts = [float(s) for s in (splits.split(",") if isinstance(splits, str) else splits)]
    parts = (parts + [0.0, 0.0, 0.0])[:3]
    total = sum(parts)
    assert total > 0
    bounds, acc = [0], 0.0
    for p in parts:
        acc += p / total
        bounds.append(int(round(acc * size)))
    bounds[-1] = size
    return bounds
