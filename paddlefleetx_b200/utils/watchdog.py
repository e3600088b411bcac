"""Failure detection and emergency checkpointing (the reference relies on the launcher's elastic heartbeat and
has no in-framework hang detection; on a long B200 job a hung NVLink collective otherwise shows up only as a silent stall).

* ``Heartbeat`` — each rank touches ``<dir>/rank_<r>.hb`` with (step, wall time); a daemon thread on every rank checks all the
  peers' files and calls ``on_stall`` (default: dump all Python stacks with ``faulthandler`` + abort the process group so the
  launcher can restart from the last checkpoint) when any rank has not advanced for ``timeout_s``.
* ``install_signal_checkpoint`` — SIGTERM/SIGUSR1 set a flag that the engine polls at step boundaries to write an emergency
  checkpoint before exiting (pre-emption on shared clusters).
* ``FaultInjector`` — test hook: ``PFX_FAULT=rank:step:kind`` (kind in hang|raise|nan) lets the test-suite exercise the above.
"""
from __future__ import annotations

import faulthandler
import json
import os
import signal
import sys
import threading
import time

from .log import logger


class Heartbeat:
    def __init__(self, directory: str, rank: int, world: int, timeout_s: float = 600.0, interval_s: float = 5.0, on_stall=None):
        self.dir, self.rank, self.world = directory, rank, world
        self.timeout_s, self.interval_s = timeout_s, interval_s
        self.on_stall = on_stall or self._default_on_stall
        os.makedirs(directory, exist_ok=True)
        self._step = -1
        self._stop = threading.Event()
        self._thread = None
        self.stalled = None

    def _path(self, r):
        return os.path.join(self.dir, f"rank_{r}.hb")

    def beat(self, step: int):
        self._step = step
        tmp = self._path(self.rank) + ".tmp"
        with open(tmp, "w") as f:
            json.dump({"step": step, "time": time.time()}, f)
        os.replace(tmp, self._path(self.rank))

    def _read(self, r):
        try:
            with open(self._path(r)) as f:
                return json.load(f)
        except (OSError, ValueError):
            return None

    def check(self, now=None):
        """Return the list of ranks whose last heartbeat is older than ``timeout_s``."""
        now = time.time() if now is None else now
        late = []
        for r in range(self.world):
            hb = self._read(r)
            if hb is not None and now - hb["time"] > self.timeout_s:
                late.append(r)
        return late

    def _default_on_stall(self, late):
        logger.error(f"[watchdog] ranks {late} made no progress for {self.timeout_s}s at step {self._step}; dumping stacks and aborting")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        os._exit(17)

    def _loop(self):
        while not self._stop.wait(self.interval_s):
            late = self.check()
            if late:
                self.stalled = late
                self.on_stall(late)
                return

    def start(self):
        self.beat(self._step)
        self._thread = threading.Thread(target=self._loop, name="pfx-heartbeat", daemon=True)
        self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2 * self.interval_s)


_EMERGENCY = threading.Event()


def install_signal_checkpoint(signals=(signal.SIGTERM, signal.SIGUSR1)):
    def handler(signum, frame):
        logger.warning(f"[watchdog] signal {signum}: emergency checkpoint requested")
        _EMERGENCY.set()

    for s in signals:
        try:
            signal.signal(s, handler)
        except ValueError:  # not in main thread
            pass


def emergency_requested() -> bool:
    return _EMERGENCY.is_set()


def clear_emergency():
    _EMERGENCY.clear()


class FaultInjector:
    """``PFX_FAULT="rank:step:kind"`` — kind: ``hang`` (sleep forever), ``raise`` (RuntimeError), ``nan`` (poison the loss)."""

    def __init__(self, spec=None, rank=0):
        spec = spec if spec is not None else os.environ.get("PFX_FAULT", "")
        self.active = False
        if spec:
            r, s, k = spec.split(":")
            self.active = int(r) == rank
            self.step, self.kind = int(s), k

    def maybe_fire(self, step, loss=None):
        if not self.active or step != self.step:
            return loss
        if self.kind == "hang":
            while True:
                time.sleep(3600)
        if self.kind == "raise":
            raise RuntimeError(f"injected fault at step {step}")
        if self.kind == "nan" and loss is not None:
            return loss * float("nan")
        return loss
