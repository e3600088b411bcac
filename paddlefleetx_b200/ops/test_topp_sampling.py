"""``python -m paddlefleetx_b200.ops.test_topp_sampling`` — stand-alone check of the sort-free top-p sampling kernel on the current GPU
(reference ppfleetx/ops/test_topp_sampling.py): the empirical distribution of many draws must match the renormalised nucleus of the input
distribution, and no token outside the nucleus may ever be drawn."""
import sys

import torch


def main(batch: int = 4, vocab: int = 50304, top_p: float = 0.75, draws: int = 2000) -> int:
    if not torch.cuda.is_available():
        print("no CUDA device: skipped")
        return 0
    from . import functional as OF

    torch.manual_seed(0)
    logits = torch.randn(batch, vocab, device="cuda") * 3
    probs = torch.softmax(logits, -1)
    sorted_p, order = probs.sort(-1, descending=True)
    csum = sorted_p.cumsum(-1)
    # u ~ U(0, top_p) lands on the first token whose cumulative mass reaches it: the boundary token only gets the part of its mass below top_p
    mass = (csum.clamp(max=top_p) - (csum - sorted_p)).clamp(min=0) / top_p
    nucleus = torch.zeros_like(probs).scatter(1, order, mass)
    counts = torch.zeros_like(probs)
    ps = torch.full((batch,), top_p, device="cuda")
    for i in range(draws):
        _, ids = OF.topp_sampling(probs, ps, seed=1234 + i)
        counts.scatter_add_(1, ids.view(batch, 1).long(), torch.ones(batch, 1, device="cuda"))
    outside = float((counts * (nucleus == 0)).sum())
    tv = float(0.5 * (counts / draws - nucleus).abs().sum(-1).max())
    # yardstick: the same number of exact multinomial draws from the nucleus distribution (finite-sample noise dominates on a large support)
    ref = torch.zeros_like(probs).scatter_add_(1, torch.multinomial(nucleus, draws, replacement=True), torch.ones(batch, draws, device="cuda"))
    tv_ref = float(0.5 * (ref / draws - nucleus).abs().sum(-1).max())
    print(f"draws outside the nucleus: {outside:.0f}; max total-variation distance to the nucleus distribution: {tv:.3f} "
          f"(exact sampler with the same number of draws: {tv_ref:.3f})")
    return 0 if outside == 0 and tv <= 1.5 * tv_ref + 0.02 else 1


if __name__ == "__main__":
    sys.exit(main())
