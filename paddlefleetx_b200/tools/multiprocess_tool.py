"""Run a list of shell commands on a pool of worker processes (reference ppfleetx/tools/multiprocess_tool.py): bulk download / unpack /
pre-process jobs described one command per line.

    python -m paddlefleetx_b200.tools.multiprocess_tool --num_proc 10 --shell_cmd_list_filename batch_cmd.txt

Commands are pulled from a shared queue (a slow command does not hold back a pre-assigned slice), each runs through ``subprocess`` with
its exit status checked, and the tool ends with a summary and a non-zero exit code if anything failed; ``--retries`` re-runs failures.
"""
import argparse
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor, as_completed


def read_commands(path):
    with open(path, encoding="utf-8") as f:
        return [line.strip() for line in f if line.strip() and not line.lstrip().startswith("#")]


read_command = read_commands          # the reference's spelling (tools/multiprocess_tool.py:60)


def process_fn(cmd_list, retries=0, timeout=None):
    """Run a slice of commands one after another; failures are printed and do not stop the slice (reference multiprocess_tool.py:50-57).
    Returns the list of ``(cmd, returncode, seconds, error_tail)`` results."""
    results = []
    for cmd in cmd_list:
        res = run_one(cmd, retries, timeout)
        if res[1] != 0:
            print(f"execute command: {cmd} failed.")
        results.append(res)
    return results


def parallel_process(cmd_list, nproc=20, retries=0, timeout=None):
    """Run ``cmd_list`` with ``nproc`` commands in flight and return the per-command results in completion order
    (reference multiprocess_tool.py:69-86; here a shared queue instead of fixed slices)."""
    import os
    import warnings

    if nproc > (os.cpu_count() or 1):
        warnings.warn("The set number of processes exceeds the number of cpu cores, please confirm whether it is reasonable.")
    results = []
    if not cmd_list:
        return results
    # threads are enough: each worker blocks in subprocess.run, the work happens in the child processes
    with ThreadPoolExecutor(max_workers=max(1, min(nproc, len(cmd_list)))) as pool:
        futures = [pool.submit(run_one, c, retries, timeout) for c in cmd_list]
        for done, fut in enumerate(as_completed(futures), 1):
            res = fut.result()
            print(f"[{done}/{len(cmd_list)}] rc={res[1]} {res[2]:.1f}s  {res[0]}", flush=True)
            results.append(res)
    return results


def run_one(cmd, retries=0, timeout=None):
    err = ""
    for attempt in range(retries + 1):
        t0 = time.time()
        try:
            r = subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=timeout)
            if r.returncode == 0:
                return cmd, 0, time.time() - t0, ""
            err = (r.stderr or r.stdout).strip()[-400:]
            code = r.returncode
        except subprocess.TimeoutExpired:
            err, code = f"timed out after {timeout}s", 124
    return cmd, code, time.time() - t0, err


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--num_proc", type=int, default=4, help="number of commands in flight")
    p.add_argument("--shell_cmd_list_filename", required=True, help="text file, one shell command per line (# comments allowed)")
    p.add_argument("--retries", type=int, default=0)
    p.add_argument("--timeout", type=float, default=None, help="seconds per command")
    a = p.parse_args(argv)
    cmds = read_commands(a.shell_cmd_list_filename)
    if not cmds:
        print("no commands to run")
        return 0
    t0 = time.time()
    failed = [(cmd, code, err) for cmd, code, _, err in parallel_process(cmds, a.num_proc, a.retries, a.timeout) if code]
    print(f"{len(cmds) - len(failed)} succeeded, {len(failed)} failed in {time.time() - t0:.1f}s")
    for cmd, code, err in failed:
        print(f"FAILED rc={code}: {cmd}\n    {err}")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
