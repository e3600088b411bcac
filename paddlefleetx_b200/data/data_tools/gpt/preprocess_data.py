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

    def tokenize(self, text):
        return self.tok.tokenize(text)

    def convert_tokens_to_ids(self, tokens):
        return self.tok.convert_tokens_to_ids(tokens)

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


# ---- Chinese word segmentation back ends (reference preprocess_data.py:136-165): each factory returns ``text -> list of words``
def lexical_analysis_fn():
    """Baidu LAC, lexical-analysis mode (words with part-of-speech tags; the tags are dropped)."""
    from LAC import LAC

    lac = LAC(mode="lac")
    return lambda line: lac.run(line)[0]


def chinese_segmentation_fn():
    """Baidu LAC, segmentation-only mode."""
    from LAC import LAC

    lac = LAC(mode="seg")
    return lambda line: lac.run(line)


def jieba_segmentation_fn():
    import jieba

    return lambda line: list(jieba.cut(line))


CHINESE_SEG_FUNC = {"lac": lexical_analysis_fn, "seg": chinese_segmentation_fn, "jieba": jieba_segmentation_fn}
_CJK = re.compile("[\u4E00-\u9FA5]")


def get_whole_word_mask_tokens(tokens, words, max_word_length=4):
    """Mark Chinese whole words on a WordPiece sequence: the pieces are single Chinese characters, ``words`` the segmentation of the same text;
    every character that continues a segmented word gets the ``##`` prefix (``通 过 利 用`` + {通过, 利用} -> ``通 ##过 利 ##用``), so that the
    masking code, which groups on ``##``, masks whole words.  Longest match first, at most ``max_word_length`` characters; pieces without
    Chinese characters pass through (they already carry WordPiece's own ``##``).  Reference preprocess_data.py:168-227."""
    vocabulary = set(words)
    out, i, n = [], 0, len(tokens)
    while i < n:
        piece = tokens[i]
        if _CJK.search(piece) is None:
            out.append(piece)
            i += 1
            continue
        span = next((k for k in range(min(max_word_length, n - i), 0, -1) if "".join(tokens[i:i + k]) in vocabulary), 1)
        out.append(piece)
        out.extend("##" + t for t in tokens[i + 1:i + span])
        i += span
    return out


def _join_words(words) -> str:
    """Re-assemble segmented words into the text the tokenizer sees: nothing between Chinese neighbours, a space between two non-Chinese words
    (a pre-split corpus has lost the original spaces; gluing ``hello`` and ``world`` together would change their word pieces)."""
    out = []
    for w in words:
        if out and not _CJK.search(out[-1][-1:]) and not _CJK.search(w[:1]) and not out[-1][-1:].isspace() and not w[:1].isspace():
            out.append(" ")
        out.append(w)
    return "".join(out)


class IdentitySplitter:
    """No sentence splitting: the document is one "sentence"."""

    def tokenize(self, *text):
        return text


class NewlineSplitter:
    """One sentence per line (the convention of the pre-split Chinese corpora)."""

    def tokenize(self, text):
        return text.split("\n")


class _RegexSplitter:
    """Punctuation-based sentence splitter used when NLTK's punkt model is not on the machine."""

    def __init__(self, chinese):
        self.chinese = chinese

    def tokenize(self, text):
        return split_sentences(text, self.chinese)


def segment_chinese(text, args):
    if args.cn_splited:
        return text.split(args.cn_split_dimer)
    try:
        return CHINESE_SEG_FUNC[getattr(args, "cn_seg_func", "jieba")]()(text)
    except ImportError:        # character-level fallback keeps the tool usable without the segmenter
        return list(text)


class Converter:
    """jsonl line -> list of sentences, each a list of token ids (reference preprocess_data.py:240-294).  ``initializer()`` builds the
    tokenizer, the sentence splitter and the Chinese segmenter once per worker process; ``encode(line)`` returns ``(doc_ids, n_bytes)``."""

    def __init__(self, args):
        self.args = args

    def initializer(self):
        a = self.args
        Converter.tokenizer = build_tokenizer(a)
        if a.split_sentences:
            Converter.splitter = _ChineseSplitter() if a.chinese else _english_splitter()
        else:
            Converter.splitter = IdentitySplitter()
        tok = Converter.tokenizer
        if a.cn_whole_word_segment and hasattr(tok, "tokenize") and hasattr(tok, "convert_tokens_to_ids"):
            if a.cn_splited:
                segment = lambda text: text.split(a.cn_split_dimer)      # noqa: E731
            else:
                try:
                    segment = CHINESE_SEG_FUNC[a.cn_seg_func]()
                except ImportError:
                    segment = list                                       # no segmenter installed: every character is its own word
            Converter.segment_func = staticmethod(segment)

            def process(text):
                words = [w for w in segment(text) if w.strip()]
                pieces = get_whole_word_mask_tokens(tok.tokenize(_join_words(words)), words)
                return tok.convert_tokens_to_ids(pieces)
        else:
            Converter.segment_func = staticmethod(lambda x: x)

            def process(text):
                return tok.encode(text)
        Converter.process = staticmethod(process)

    def encode(self, json_line):
        text = json.loads(json_line).get(self.args.json_key, "")
        doc_ids = []
        for sentence in Converter.splitter.tokenize(text):
            ids = Converter.process(sentence.strip()) if sentence.strip() else []
            if len(ids) > 0:
                doc_ids.append(list(ids))
        if doc_ids and self.args.append_eos:
            doc_ids[-1].append(Converter.tokenizer.eos_token_id)
        return doc_ids, len(text.encode("utf-8"))


class _ChineseSplitter:
    """Newline-separated sentences when the document has them (pre-split corpora), punctuation otherwise."""

    def tokenize(self, text):
        return [s for s in text.split("\n") if s.strip()] if "\n" in text.strip() else split_sentences(text, True)


def _english_splitter():
    try:
        import nltk

        return nltk.load("tokenizers/punkt/english.pickle")
    except Exception:            # noqa: BLE001 - nltk or its punkt data missing (offline): regex splitter
        return _RegexSplitter(False)


_CONVERTER = None


def _init(args):
    global _TOK, _ARGS, _CONVERTER
    _ARGS = args
    _CONVERTER = Converter(args)
    _CONVERTER.initializer()
    _TOK = Converter.tokenizer


def encode_line(line):
    line = line.strip()
    if not line:
        return [], 0
    sents, _ = _CONVERTER.encode(line)
    return sents, len(line)


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
