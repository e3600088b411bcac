"""GPU debug: the exact body of tests/test_gpu_models.py::test_vision_multimodal_and_text_towers_on_gpu (ViT part), with probes."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

from paddlefleetx_b200.models.vision_model.factory import build  # noqa: E402
from paddlefleetx_b200.optims import FusedAdamW  # noqa: E402

dev = "cuda"
for variant in ("all_in_autocast", "step_outside"):
    torch.manual_seed(0)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16)
    ctx.__enter__()
    vit = build(dict(name="ViT_tiny_patch16_224", img_size=64, patch_size=8, depth=2, class_num=10)).to(dev)
    opt = FusedAdamW(1e-3, named_parameters=list(vit.named_parameters()))
    x, y = torch.randn(8, 3, 64, 64, device=dev), torch.randint(0, 10, (8,), device=dev)
    for step in range(3):
        loss = build(dict(name="CELoss", epsilon=0.1))(vit(x), y)
        if variant == "step_outside":
            ctx.__exit__(None, None, None)
        loss.backward()
        mg = {n: (str(p.main_grad.dtype), float(p.main_grad.float().norm())) for n, p in vit.named_parameters() if n.startswith("head")}
        w0 = float(vit.head.weight.float().abs().sum())
        opt.step()
        w1 = float(vit.head.weight.float().abs().sum())
        print(variant, step, "loss", float(loss), "gnorm", float(opt._gnorm), "found_inf", float(opt._found_inf), "head main_grad", mg, "head|w| before/after", w0, w1,
              "lr", opt.get_lr(), "groups", [(g.key, g.param_buf.dtype, g.grad_buf.dtype, g.meta["has_master"]) for g in opt.groups], flush=True)
        opt.clear_grad()
        if variant == "step_outside":
            ctx.__enter__()
    ctx.__exit__(None, None, None)
