"""Multi-GPU evidence reachable from ``pytest -m gpu``: on any lease with >= 2 GPUs the peer-memory / NVLS kernels (symmetric allocator,
multimem reduce-scatter / all-gather / AdamW-broadcast, fused GEMM+collective kernels, flat-optimizer equivalence, MoE dispatch/combine) run
against NCCL + cuBLAS through ``tools/gpu_multi_selftest.py`` (one rank per GPU, torchrun)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs >= 2 GPUs on one node")
def test_peer_memory_and_nvls_kernels_match_nccl():
    n = 2 if _gpu_count() < 8 else 8
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "gpu_multi_selftest.py")]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    results = [json.loads(l[7:]) for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    bad = [r for r in results if not r.get("ok")]
    assert p.returncode == 0 and results and not bad, (p.returncode, bad, p.stdout[-2000:], p.stderr[-2000:])
