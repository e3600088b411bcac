"""Tiny driver for profiling the flash-attention kernels (ncu target): forward + backward at the GPT-6.7B layer shape.

    ncu --set full --import-source on --clock-control none -k regex:attention_ -c 4 -o gpurun_out/attn python tools/attn_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from paddlefleetx_b200.ops import _native  # noqa: E402


def main():
    B, S, H, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (8, 1024, 32, 128)))
    p = float(sys.argv[5]) if len(sys.argv) >= 6 else 0.1
    iters = int(sys.argv[6]) if len(sys.argv) >= 7 else 2
    lib = _native.require()
    torch.manual_seed(0)
    mix = torch.randn(B, S, H, 3, D, device="cuda").bfloat16()
    q, k, v = mix.unbind(3)
    dmix = torch.empty_like(mix)
    dq, dk, dv = dmix.unbind(3)
    go = torch.randn(B, S, H, D, device="cuda").bfloat16()
    scale = D ** -0.5
    for _ in range(iters):
        out, lse = lib.attention_fwd_v2(q, k, v, True, scale, p, 7)
        lib.attention_bwd(q, k, v, out, go, lse, dq, dk, dv, True, scale, p, 7)
    torch.cuda.synchronize()
    print("attn_bench done", float(out.float().abs().mean()), float(dmix.float().abs().mean()))


if __name__ == "__main__":
    main()
