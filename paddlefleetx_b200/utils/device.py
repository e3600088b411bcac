"""Device helpers (reference ppfleetx/utils/device.py:19-61) — one accelerator backend: CUDA."""
import torch


def get_device(configs=None) -> str:
    want = str(configs.Global.get("device", "gpu")).lower() if configs is not None else "gpu"
    return "gpu" if want == "gpu" and torch.cuda.is_available() else "cpu"


def torch_device(configs=None) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if get_device(configs) == "gpu" else torch.device("cpu")


def synchronize() -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()
