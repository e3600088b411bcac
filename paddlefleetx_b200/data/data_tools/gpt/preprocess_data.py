"""jsonl -> token-id corpus: ``<prefix>_ids.npy`` (all documents' ids, flat) + ``<prefix>_idx.npz`` (``lens`` per document; with
``--split_sentences`` ``lens`` per sentence + ``docs`` = document boundaries in sentences) — the on-disk format ``GPTDataset`` / ``ErnieDataset`` mmap.

CLI parity with the reference tool (ppfleetx/data/data_tools/gpt/preprocess_data.py:40-125): ``--model_name``,
``--tokenizer_name``, ``--input_path``, ``--output_prefix``, ``--data_format JSON``, ``--json_key``, ``--split_sentences``,
``--chinese``, ``--cn_whole_word_segment``, ``--cn_seg_func``, ``--cn_splited``, ``--cn_split_dimer``, ``--append_eos``,
``--log_interval``, ``--workers``.  Tokenisation runs in a process pool (one tokenizer per worker, chunked imap); ids are
appended to fixed-size numpy blocks, so peak memory is O(block) + the final concatenation.
"""
import argparse
import json
import multiprocessing as mp
import os
import re
import sys
import time

import numpy as np

_TOK = None
_ARGS = None


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model_name", default="gpt2")
    p.add_argument("--tokenizer_name", default="GPTTokenizer", choices=["GPTTokenizer", "GPTChineseTokenizer", "ErnieTokenizer", "ByteTokenizer"])
    g = p.add_argument_group("data input/output")
    g.add_argument("--input_path", required=True, help="jsonl file or folder of jsonl files")
    g.add_argument("--output_prefix", required=True)
    g.add_argument("--data_format", default="JSON", choices=["JSON"])
    g.add_argument("--json_key", default="text")
    g.add_argument("--split_sentences", action="store_true")
    g = p.add_argument_group("chinese words")
    g.add_argument("--chinese", action="store_true")
    g.add_argument("--cn_whole_word_segment", action="store_true")
    g.add_argument("--cn_seg_func", default="jieba", choices=["lac", "seg", "jieba"])
    g.add_argument("--cn_splited", action="store_true", help="corpus is already word-segmented")
    g.add_argument("--cn_split_dimer", default=" ")
    g = p.add_argument_group("common config")
    g.add_argument("--append_eos", action="store_true")
    g.add_argument("--log_interval", type=int, default=100)
    g.add_argument("--workers", type=int, default=1)
    return p.parse_args(argv)


class _CorpusWordPiece:
    """ERNIE tokenizer as the corpus tools need it: ids without the ``[CLS] ... [SEP]`` frame, documents closed by ``[SEP]``."""

    def __init__(self, tok):
        self.tok, self.eos_token_id = tok, tok.sep_token_id

    def encode(self, text):
        return self.tok.encode_plain(text)

    def __len__(self):
        return len(self.tok)


def build_tokenizer(args):
    """``--tokenizer_name`` x ``--model_name`` (a local vocabulary directory or a cached name) -> an object with ``encode(text) -> ids``
    *without* sequence-template tokens, ``eos_token_id`` and ``__len__``."""
    from paddlefleetx_b200.data.tokenizers import ErnieTokenizer, GPTChineseTokenizer, GPTTokenizer

    if args.tokenizer_name == "ByteTokenizer":
        return GPTTokenizer.byte_fallback()
    if args.tokenizer_name == "ErnieTokenizer":
        return _CorpusWordPiece(ErnieTokenizer.from_pretrained(args.model_name))
    if args.tokenizer_name == "GPTChineseTokenizer":
        return GPTChineseTokenizer.from_pretrained(args.model_name)
    try:
        return GPTTokenizer.from_pretrained(args.model_name)
    except FileNotFoundError as exc:   # offline box without the vocab files
        print(f"[preprocess] {exc}; falling back to the byte-level tokenizer", file=sys.stderr)
        return GPTTokenizer.byte_fallback()


_SENT_EN = re.compile(r"(?<=[.!?])\s+")
_SENT_CN = re.compile(r"(?<=[。！？；!?;])")


def split_sentences(text, chinese):
    parts = (_SENT_CN if chinese else _SENT_EN).split(text)
    return [s for s in (p.strip() for p in parts) if s]


def segment_chinese(text, args):
    if args.cn_splited:
        return text.split(args.cn_split_dimer)
    try:
        import jieba

        return list(jieba.cut(text))
    except ImportError:        # character-level fallback keeps the tool usable without the segmenter
        return list(text)


def _init(args):
    global _TOK, _ARGS
    _ARGS = args
    _TOK = build_tokenizer(args)


def encode_line(line):
    line = line.strip()
    if not line:
        return [], 0
    text = json.loads(line).get(_ARGS.json_key, "")
    if not text:
        return [], len(line)
    sents = split_sentences(text, _ARGS.chinese) if _ARGS.split_sentences else [text]
    out = []
    for s in sents:
        if _ARGS.chinese and _ARGS.cn_whole_word_segment:
            s = " ".join(segment_chinese(s, _ARGS))
        ids = _TOK.encode(s)
        if ids:
            out.append(ids)
    if out and _ARGS.append_eos:
        out[-1] = list(out[-1]) + [_TOK.eos_token_id]
    return out, len(line)


def main(argv=None):
    a = get_args(argv)
    files = sorted(os.path.join(a.input_path, f) for f in os.listdir(a.input_path)) if os.path.isdir(a.input_path) else [a.input_path]
    files = [f for f in files if f.endswith((".json", ".jsonl"))] or files
    vocab_probe = build_tokenizer(a)
    dtype = np.uint16 if len(vocab_probe) < 65500 else np.int32
    blocks, lens, sent_lens, doc_sent_counts = [], [], [], []
    t0, nbytes, ndocs = time.time(), 0, 0
    pool = mp.Pool(a.workers, initializer=_init, initargs=(a,)) if a.workers > 1 else None
    if pool is None:
        _init(a)
    for path in files:
        with open(path, "r", encoding="utf-8") as f:
            it = pool.imap(encode_line, f, 64) if pool else map(encode_line, f)
            for sents, nb in it:
                nbytes += nb
                if not sents:
                    continue
                ndocs += 1
                doc_len = 0
                for ids in sents:
                    blocks.append(np.asarray(ids, dtype=dtype))
                    sent_lens.append(len(ids))
                    doc_len += len(ids)
                lens.append(doc_len)
                doc_sent_counts.append(len(sents))
                if ndocs % a.log_interval == 0:
                    dt = time.time() - t0
                    print(f"processed {ndocs} documents ({ndocs / dt:.1f} docs/s, {nbytes / dt / 1e6:.2f} MB/s)", file=sys.stderr)
    if pool:
        pool.close()
    ids = np.concatenate(blocks) if blocks else np.zeros(0, dtype=dtype)
    os.makedirs(os.path.dirname(os.path.abspath(a.output_prefix)) or ".", exist_ok=True)
    np.save(a.output_prefix + "_ids.npy", ids)
    if a.split_sentences:
        # sentence-addressable corpus (ERNIE; reference create_pretraining_data.py:400-405): ``lens`` = tokens per SENTENCE, ``docs`` = document
        # boundaries counted in sentences — what MMapIndexedDataset / build_mapping read
        np.savez(a.output_prefix + "_idx.npz", lens=np.asarray(sent_lens, dtype=np.int32), docs=np.cumsum([0] + doc_sent_counts).astype(np.int64))
        print(f"{ndocs} documents, {len(sent_lens)} sentences, {ids.size} tokens -> {a.output_prefix}_ids.npy / _idx.npz")
    else:
        # document-addressable corpus (GPT): ``lens`` = tokens per document
        np.savez(a.output_prefix + "_idx.npz", lens=np.asarray(lens, dtype=np.int32))
        print(f"{ndocs} documents, {ids.size} tokens -> {a.output_prefix}_ids.npy / _idx.npz")


if __name__ == "__main__":
    main()
