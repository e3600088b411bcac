"""Spawn N gloo ranks on CPU and run a top-level function in each (multi-process tests without a GPU)."""
import os
import socket
import sys
import traceback

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, fn, args, errq):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        if isinstance(fn, str):          # "module:function" resolved inside the child (spawn cannot pickle test-module functions)
            import importlib

            mod, name = fn.split(":")
            fn = getattr(importlib.import_module(mod), name)
        import torch
        import torch.distributed as dist

        torch.set_num_threads(1)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world: int, *args, timeout: int = 240):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.terminate()
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    assert not alive, f"{len(alive)} ranks timed out"
    assert not errs, "\n".join(f"[rank {r}]\n{tb}" for r, tb in errs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
