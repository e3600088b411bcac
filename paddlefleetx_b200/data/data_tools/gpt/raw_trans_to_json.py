"""Raw text -> jsonl (one ``{"text": ...}`` document per line).

Same command line as the reference tool (ppfleetx/data/data_tools/gpt/raw_trans_to_json.py:28-75): ``--input_path`` (file or
folder), ``--output_path``, ``--json_key``, ``--doc_spliter`` (empty = blank line separates documents), ``--min_doc_length``,
``--workers``, ``--log_interval``, ``--no-merge``, ``--no-shuffle``.  Files are processed in a process pool, each worker streams
its file and emits documents as soon as the splitter is seen, so memory stays flat for multi-GB inputs.
"""
import argparse
import json
import multiprocessing as mp
import os
import random
import shutil
import sys
import time


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--input_path", required=True, help="raw text file or a folder of them")
    p.add_argument("--output_path", required=True, help="output prefix; <output_path>.jsonl is written")
    p.add_argument("--json_key", default="text")
    p.add_argument("--doc_spliter", default="", help="line that separates documents (default: blank line)")
    p.add_argument("--min_doc_length", type=int, default=10)
    p.add_argument("--workers", type=int, default=1)
    p.add_argument("--log_interval", type=int, default=1)
    p.add_argument("--no-merge", action="store_true", help="keep one jsonl per input file")
    p.add_argument("--no-shuffle", action="store_true", help="keep the input file order")
    return p.parse_args(argv)


def iter_documents(path, spliter, min_len):
    buf = []
    with open(path, "r", encoding="utf-8", errors="ignore") as f:
        for line in f:
            s = line.strip()
            if s == spliter:
                doc = "\n".join(buf).strip()
                buf = []
                if len(doc) >= min_len:
                    yield doc
            elif s:
                buf.append(s)
    doc = "\n".join(buf).strip()
    if len(doc) >= min_len:
        yield doc


def convert_file(job):
    path, out_path, key, spliter, min_len = job
    n = 0
    with open(out_path, "w", encoding="utf-8") as out:
        for doc in iter_documents(path, spliter, min_len):
            out.write(json.dumps({key: doc}, ensure_ascii=False) + "\n")
            n += 1
    return path, out_path, n


def raw_text_to_json(path, doc_spliter="", json_key="text", min_doc_length=10):
    """One raw text file -> ``<path>.jsonl``; returns ``(bytes read, output path)`` (``(0, None)`` for a missing file) — the reference's
    per-file entry point (raw_trans_to_json.py:75-105)."""
    path = os.path.abspath(path)
    if not os.path.exists(path):
        print("No found file %s" % path)
        return 0, None
    out_path = path + ".jsonl"
    convert_file((path, out_path, json_key, doc_spliter, min_doc_length))
    return os.path.getsize(path), out_path


def merge_file(file_paths, output_path):
    """Concatenate the per-file jsonl parts into ``output_path`` (``.jsonl`` appended when missing) and delete the parts."""
    if not output_path.endswith(".jsonl"):
        output_path += ".jsonl"
    print("Merging files into %s" % output_path)
    with open(output_path, "wb") as out:
        for part in file_paths:
            if part is not None and os.path.exists(part):
                with open(part, "rb") as f:
                    shutil.copyfileobj(f, out)
                os.remove(part)
    print("File save in %s" % output_path)
    return output_path


def shuffle_file(output_path, seed: int = 1234):
    """Shuffle the lines of a jsonl file in place (the reference shells out to ``shuf``; here: line offsets shuffled in Python, the file
    re-written through a temporary — no dependency on coreutils, deterministic under ``seed``)."""
    if not os.path.exists(output_path):
        raise ValueError("File not found: %s" % output_path)
    print("Shuffling the jsonl file...")
    offsets = []
    with open(output_path, "rb") as f:
        pos = 0
        for line in f:
            offsets.append((pos, len(line)))
            pos += len(line)
    random.Random(seed).shuffle(offsets)
    tmp = output_path + ".shuf"
    with open(output_path, "rb") as f, open(tmp, "wb") as out:
        for pos, n in offsets:
            f.seek(pos)
            line = f.read(n)
            out.write(line if line.endswith(b"\n") else line + b"\n")
    os.replace(tmp, output_path)
    print("File shuffled!!!")


def main(argv=None):
    a = get_args(argv)
    if os.path.isdir(a.input_path):
        files = sorted(os.path.join(r, f) for r, _, fs in os.walk(a.input_path) for f in fs)
    else:
        files = [a.input_path]
    if not a.no_shuffle:
        random.Random(1234).shuffle(files)
    os.makedirs(os.path.dirname(os.path.abspath(a.output_path)) or ".", exist_ok=True)
    jobs = [(f, f"{a.output_path}.part{i:05d}.jsonl", a.json_key, a.doc_spliter, a.min_doc_length) for i, f in enumerate(files)]
    t0, total, parts = time.time(), 0, []
    pool = mp.Pool(a.workers) if a.workers > 1 else None
    it = pool.imap(convert_file, jobs) if pool else map(convert_file, jobs)
    for i, (src, part, n) in enumerate(it, 1):
        total += n
        parts.append(part)
        if i % a.log_interval == 0:
            print(f"[{i}/{len(jobs)}] {src}: {n} docs ({total} total, {time.time() - t0:.1f}s)", file=sys.stderr)
    if pool:
        pool.close()
    if not a.no_merge:
        with open(a.output_path + ".jsonl", "wb") as out:
            for part in parts:
                with open(part, "rb") as f:
                    shutil.copyfileobj(f, out)
                os.remove(part)
    print(f"{total} documents -> {a.output_path}{'.jsonl' if not a.no_merge else '.part*.jsonl'}")


if __name__ == "__main__":
    main()
