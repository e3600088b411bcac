"""ERNIE corpus preparation (reference ppfleetx/data/data_tools/ernie/preprocess/create_pretraining_data.py): sentence-split
tokenisation to ``<prefix>_ids.npy`` + ``<prefix>_idx.npz`` with per-sentence ``lens`` and ``docs`` boundaries, which ``ErnieDataset`` needs
for sentence-order prediction and span masking.  Thin front-end over the GPT tool with ``--split_sentences`` forced on."""
import sys

from ..gpt import preprocess_data
from ..gpt.preprocess_data import (Converter, IdentitySplitter, NewlineSplitter, chinese_segmentation_fn, get_whole_word_mask_tokens,  # noqa: F401
                                   jieba_segmentation_fn, lexical_analysis_fn)


def get_args(argv=None):
    """The GPT tool's arguments with ``--split_sentences`` forced on (reference create_pretraining_data.py:40-130)."""
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--split_sentences" not in argv:
        argv.append("--split_sentences")
    return preprocess_data.get_args(argv)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--split_sentences" not in argv:
        argv.append("--split_sentences")
    preprocess_data.main(argv)


if __name__ == "__main__":
    main()
