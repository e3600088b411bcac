"""Summarise a ``bench.py --profile`` kernel timeline: per-kernel device time, per-stream busy time, idle gaps on the compute stream, and how
much communication-stream time is hidden under compute.

    python tools/analyze_trace.py gpurun_out/bench_trace_n2_rank0.json.gz [--top 30]
"""
import gzip
import json
import sys
from collections import defaultdict


def merge(iv):
    iv = sorted(iv)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def total(iv):
    return sum(b - a for a, b in iv)


def intersect(x, y):
    i = j = 0
    s = 0.0
    while i < len(x) and j < len(y):
        a, b = max(x[i][0], y[j][0]), min(x[i][1], y[j][1])
        if a < b:
            s += b - a
        if x[i][1] < y[j][1]:
            i += 1
        else:
            j += 1
    return s


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    d = json.load(gzip.open(path, "rt"))
    ks = d["kernels"]
    steps = d["steps"]
    t0, t1 = ks[0][2], max(k[2] + k[3] for k in ks)
    span = (t1 - t0) / steps
    by_name, by_stream = defaultdict(lambda: [0.0, 0]), defaultdict(list)
    for name, stream, ts, dur in ks:
        name = name.replace("(anonymous namespace)::", "")
        key = name.split("<")[0].split("(")[0][:70] if not name.startswith("void pfx::gemm") else name[:60]
        by_name[key][0] += dur
        by_name[key][1] += 1
        by_stream[stream].append((ts, ts + dur))
    print(f"{path}: {len(ks)} kernels over {steps} steps, span {span / 1e3:.2f} ms/step")
    merged = {s: merge(v) for s, v in by_stream.items()}
    main_stream = max(merged, key=lambda s: total(merged[s]))
    print("streams (busy ms/step, kernels): " + ", ".join(f"{s}: {total(m) / steps / 1e3:.2f} ms ({len(by_stream[s])})" for s, m in sorted(merged.items(), key=lambda kv: -total(kv[1]))))
    busy_main = total(merged[main_stream]) / steps
    print(f"compute stream {main_stream}: busy {busy_main / 1e3:.2f} ms/step, idle {(span - busy_main) / 1e3:.2f} ms/step")
    for s, m in merged.items():
        if s != main_stream:
            ov = intersect(m, merged[main_stream])
            print(f"  stream {s}: {total(m) / steps / 1e3:.2f} ms/step busy, {ov / steps / 1e3:.2f} ms/step of it under compute-stream kernels")
    tot = sum(v[0] for v in by_name.values())
    print(f"per-kernel device time (sum {tot / steps / 1e3:.2f} ms/step):")
    for name, (dur, cnt) in sorted(by_name.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {dur / steps / 1e3:8.3f} ms/step  {100 * dur / tot:5.1f}%  x{cnt / steps:7.1f}  avg {dur / cnt:8.1f} us  {name}")


if __name__ == "__main__":
    main()
