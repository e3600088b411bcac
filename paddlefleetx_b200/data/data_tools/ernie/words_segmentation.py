"""Chinese sentence splitting + word segmentation ahead of whole-word masking (reference .../ernie/preprocess/words_segmentation.py).

Two flows share the same segmentation code:

* the reference's: ``--input_path <file or folder> --output_path <folder> --data_format jsonl|wudao --cn_seg_func jieba|lac|seg`` reads the
  ``text`` (jsonl) / ``content`` (WuDao json) field of every document and writes one TEXT file per input: one sentence per line, words separated
  by a space, a blank line between documents — the input of ``trans_to_json``;
* jsonl in, jsonl out (``--output_path something.jsonl``): every document's ``--json_key`` is rewritten with its words separated by
  ``--cn_split_dimer``; downstream use ``create_pretraining_data --cn_splited``.
"""
import argparse
import json
import multiprocessing as mp
import os
import re
import sys
import time
from functools import partial

from ..gpt.preprocess_data import CHINESE_SEG_FUNC, chinese_segmentation_fn, jieba_segmentation_fn, lexical_analysis_fn  # noqa: F401

special_chars = ["\n", "。", "?", "？", " ", ";", "；", "！", "!"]
split_chars = ["。", "?", "？", ";", "；", "!", "！"]


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--input_path", required=True, help="raw file, or a folder of them")
    p.add_argument("--workers", type=int, default=1)
    p.add_argument("--output_path", default="./tmp", help="folder for the segmented text files, or a .jsonl file for the jsonl -> jsonl flow")
    p.add_argument("--data_format", default="jsonl", choices=["jsonl", "wudao"])
    p.add_argument("--cn_seg_func", default="jieba", choices=["lac", "seg", "jieba"])
    p.add_argument("--log_interval", type=int, default=1)
    p.add_argument("--json_key", default="text")
    p.add_argument("--cn_split_dimer", default=" ")
    return p.parse_args(argv)


def _segmenter(name: str):
    """``text -> words`` for the chosen back end; characters when the library is not installed (the tool stays usable offline)."""
    try:
        return CHINESE_SEG_FUNC[name]()
    except ImportError:
        return list


def read_wudao(path):
    """Documents of one WuDao shard: a json list of ``{"content": ...}``."""
    print("Loading %s" % path)
    with open(path, "r", encoding="utf-8") as f:
        try:
            contents = json.load(f)
        except Exception:                    # noqa: BLE001 - a broken shard is skipped, not fatal
            print("Failed to load %s" % path)
            return
    for js in contents:
        yield js["content"]


def read_jsonl(path, key: str = "text"):
    """Documents of a jsonl file: the ``text`` field of every line."""
    print("Loading %s" % path)
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                yield json.loads(line)[key]


READFILE_FUNC = {"jsonl": read_jsonl, "wudao": read_wudao}


def split_and_segment(text: str, seg) -> str:
    """One document -> lines of space-separated words, one sentence per line: runs of a separator (and blanks after it) collapse to one, every
    sentence-final mark ends a line (reference words_segmentation.py:156-170)."""
    for ch in special_chars:
        text = re.sub("[" + re.escape(ch) + "]+[ ]*", lambda _m, c=ch: c, text)
    for ch in split_chars:
        text = text.replace(ch, ch + "\n")
    return "".join(" ".join(w for w in seg(line) if w.strip()) + "\n" for line in text.split("\n") if line)


def text_to_text(path, output_path, read_func, seg_func):
    """Segment one input file into ``output_path/<tail of the input name>``; returns ``(bytes read, None)`` (skips outputs that already exist)."""
    out_name = os.path.join(output_path, os.path.basename(path)[-20:])
    print("Write into %s" % out_name)
    if os.path.exists(out_name):
        print("File exists %s" % out_name)
        return 0, None
    seg = _segmenter(seg_func) if isinstance(seg_func, str) else seg_func
    reader = READFILE_FUNC[read_func] if isinstance(read_func, str) else read_func
    nbytes = 0
    with open(out_name, "w", encoding="utf-8") as f:
        for text in reader(path):
            nbytes += len(text.encode("utf-8"))
            f.write(split_and_segment(text, seg) + "\n")
    return nbytes, None


# ---- jsonl -> jsonl flow
_SEG = None


def _init_worker(seg_name):
    global _SEG
    _SEG = _segmenter(seg_name)


def _segment_line(job):
    line, key, dimer = job
    d = json.loads(line)
    d[key] = dimer.join(w for w in _SEG(d.get(key, "")) if w.strip())
    return json.dumps(d, ensure_ascii=False)


def _jsonl_to_jsonl(a):
    with open(a.input_path, encoding="utf-8") as f, open(a.output_path, "w", encoding="utf-8") as out:
        jobs = ((line, a.json_key, a.cn_split_dimer) for line in f if line.strip())
        if a.workers > 1:
            with mp.Pool(a.workers, initializer=_init_worker, initargs=(a.cn_seg_func,)) as pool:
                for r in pool.imap(_segment_line, jobs, 64):
                    out.write(r + "\n")
        else:
            _init_worker(a.cn_seg_func)
            for j in jobs:
                out.write(_segment_line(j) + "\n")


def main(argv=None):
    a = get_args(argv)
    if a.output_path.endswith((".jsonl", ".json")) and os.path.isfile(a.input_path):
        return _jsonl_to_jsonl(a)
    files = [a.input_path] if os.path.isfile(a.input_path) else sorted(os.path.join(r, f) for r, _, fs in os.walk(a.input_path) for f in fs)
    os.makedirs(a.output_path, exist_ok=True)
    job = partial(text_to_text, output_path=a.output_path, seg_func=a.cn_seg_func, read_func=a.data_format)
    t0, total = time.time(), 0
    pool = mp.Pool(a.workers) if a.workers > 1 else None
    for i, (nbytes, _) in enumerate(pool.imap(job, files, 1) if pool else map(job, files), 1):
        total += nbytes
        if i % a.log_interval == 0:
            dt = max(time.time() - t0, 1e-9)
            print(f"Processed {i} files ({i / dt:.2f} files/s, {total / dt / 1024 / 1024:.2f} MB/s).", file=sys.stderr)
    if pool:
        pool.close()


if __name__ == "__main__":
    main()
