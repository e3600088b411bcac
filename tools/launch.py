"""Process launcher: one worker per GPU, per-rank logs, group restart on failure.

Counterpart of ``python -m paddle.distributed.launch --log_dir … --devices "0,…,7" [--master ip:port --nnodes N --rank R]
[--max_restart K]`` used by every reference script (docs/quick_start.md:128-172).  Exports the ``torch.distributed`` env contract
(RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT), writes ``<log_dir>/workerlog.<local rank>`` (rank 0 is tee'd to
the console), and when a worker dies terminates its siblings by PID and — up to ``--max_restart`` times — starts the whole
group again (training resumes from the last checkpoint through ``Engine.save_load.ckpt_dir`` / auto-resume).

    python tools/launch.py --devices 0,1,2,3,4,5,6,7 --log_dir log tools/train.py -c cfg.yaml -o ...
"""
import argparse
import os
import signal
import subprocess
import sys
import threading
import time


def parse():
    p = argparse.ArgumentParser(allow_abbrev=False)
    p.add_argument("--devices", "--gpus", default=None, help='comma list of device ids, default: all visible (or "cpu:N" for N CPU workers)')
    p.add_argument("--log_dir", default="log")
    p.add_argument("--nnodes", type=int, default=1)
    p.add_argument("--rank", "--node_rank", type=int, default=0, dest="node_rank")
    p.add_argument("--master", default="127.0.0.1:29500", help="ip:port of node 0")
    p.add_argument("--max_restart", type=int, default=0)
    p.add_argument("--job_id", default="default")
    p.add_argument("script")
    p.add_argument("script_args", nargs=argparse.REMAINDER)
    return p.parse_args()


def _device_list(spec):
    if spec and spec.startswith("cpu:"):
        return [None] * int(spec.split(":")[1])
    if spec:
        return [d for d in spec.split(",") if d != ""]
    try:
        import torch

        n = torch.cuda.device_count()
    except (ImportError, RuntimeError):
        n = 0
    return [str(i) for i in range(n)] or [None]


def _tee(stream, path, echo):
    with open(path, "ab", buffering=0) as f:
        for line in iter(stream.readline, b""):
            f.write(line)
            if echo:
                sys.stdout.buffer.write(line)
                sys.stdout.buffer.flush()


def run_group(a, devices, attempt):
    addr, port = a.master.rsplit(":", 1)
    nproc = len(devices)
    world = nproc * a.nnodes
    procs, threads = [], []
    os.makedirs(a.log_dir, exist_ok=True)
    for local, dev in enumerate(devices):
        env = dict(os.environ, RANK=str(a.node_rank * nproc + local), LOCAL_RANK=str(local), WORLD_SIZE=str(world),
                   LOCAL_WORLD_SIZE=str(nproc), MASTER_ADDR=addr, MASTER_PORT=str(int(port) + attempt), PFX_JOB_ID=a.job_id,
                   PFX_RESTART_COUNT=str(attempt))
        if dev is not None:
            env["CUDA_VISIBLE_DEVICES"] = ",".join(d for d in devices)      # LOCAL_RANK indexes into this list
        p = subprocess.Popen([sys.executable, "-u", a.script] + a.script_args, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        t = threading.Thread(target=_tee, args=(p.stdout, os.path.join(a.log_dir, f"workerlog.{local}"), local == 0), daemon=True)
        t.start()
        procs.append(p)
        threads.append(t)
    failed = None
    while failed is None and any(p.poll() is None for p in procs):
        for p in procs:
            rc = p.poll()
            if rc not in (None, 0):
                failed = rc
                break
        time.sleep(0.2)
    if failed is None:
        failed = next((p.returncode for p in procs if p.returncode), 0)
    if failed:
        for p in procs:                      # stop the survivors: exact PIDs only
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)
        deadline = time.time() + 20
        for p in procs:
            try:
                p.wait(timeout=max(0.1, deadline - time.time()))
            except subprocess.TimeoutExpired:
                p.kill()
    for t in threads:
        t.join(timeout=2)
    return failed


def main():
    a = parse()
    devices = _device_list(a.devices)
    for attempt in range(a.max_restart + 1):
        rc = run_group(a, devices, attempt)
        if rc == 0:
            return 0
        print(f"[launch] worker exited with {rc} (attempt {attempt + 1}/{a.max_restart + 1})", file=sys.stderr)
    return rc


if __name__ == "__main__":
    sys.exit(main())
