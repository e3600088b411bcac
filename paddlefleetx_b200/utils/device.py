"""Device helpers (reference ppfleetx/utils/device.py:19-61) — one accelerator backend: CUDA."""
import torch


def get_device_and_mapping():
    """``(device, {name: available})`` — which accelerator back ends this build can drive (reference utils/device.py:19-33).  This
    framework is written for one: CUDA on sm_100a; everything else reports ``False`` and ``cpu`` is the fallback for plumbing tests."""
    supported = {"gpu": torch.cuda.is_available(), "xpu": False, "rocm": False, "npu": False, "mlu": False, "cpu": True}
    return next(d for d, ok in supported.items() if ok), supported


def get_device(configs=None) -> str:
    want = str(configs.Global.get("device", "gpu")).lower() if configs is not None else "gpu"
    return "gpu" if want == "gpu" and torch.cuda.is_available() else "cpu"


def torch_device(configs=None) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if get_device(configs) == "gpu" else torch.device("cpu")


def synchronize() -> bool:
    """Wait for the device; ``True`` when something was synchronised (reference utils/device.py:44-61)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        return True
    return False
