"""ERNIE corpus preparation (reference ppfleetx/data/data_tools/ernie/preprocess/create_pretraining_data.py): sentence-split
tokenisation to ``<prefix>_ids.npy`` + ``<prefix>_idx.npz`` with per-sentence ``lens`` and ``docs`` boundaries, which ``ErnieDataset`` needs
for sentence-order prediction and span masking.  Thin front-end over the GPT tool with ``--split_sentences`` forced on."""
import sys

from ..gpt import preprocess_data


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--split_sentences" not in argv:
        argv.append("--split_sentences")
    preprocess_data.main(argv)


if __name__ == "__main__":
    main()
