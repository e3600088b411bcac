"""Chinese word segmentation pass for whole-word masking (reference .../ernie/preprocess/words_segmentation.py): rewrites a
jsonl corpus so that words are separated by ``--cn_split_dimer``; downstream use ``preprocess_data --cn_splited``."""
import argparse
import json
import multiprocessing as mp


def _segment(line, key, dimer):
    d = json.loads(line)
    text = d.get(key, "")
    try:
        import jieba

        words = list(jieba.cut(text))
    except ImportError:
        words = list(text)
    d[key] = dimer.join(w for w in words if w.strip())
    return json.dumps(d, ensure_ascii=False)


def _job(args):
    return _segment(*args)


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--input_path", required=True)
    p.add_argument("--output_path", required=True)
    p.add_argument("--json_key", default="text")
    p.add_argument("--cn_split_dimer", default=" ")
    p.add_argument("--workers", type=int, default=1)
    a = p.parse_args(argv)
    with open(a.input_path, encoding="utf-8") as f, open(a.output_path, "w", encoding="utf-8") as out:
        jobs = ((line, a.json_key, a.cn_split_dimer) for line in f if line.strip())
        if a.workers > 1:
            with mp.Pool(a.workers) as pool:
                for r in pool.imap(_job, jobs, 64):
                    out.write(r + "\n")
        else:
            for j in jobs:
                out.write(_job(j) + "\n")


if __name__ == "__main__":
    main()
