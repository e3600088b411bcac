"""Config post-processing of the multimodal (Imagen) recipes — reference ppfleetx/models/multimodal_model/utils.py:31-137.  Each step fills
derived values into one section of the config; sections a recipe does not have are left alone (the reference indexes them unconditionally and
its recipes always carry ``Fused`` / ``Inference``)."""
from __future__ import annotations

from ...distributed.apis import env
from ...utils.log import logger


def process_global_configs(config) -> None:
    """Reconcile global / local batch size over the data replicas (dp x sharding); one of the two may be ``None``."""
    dp = config["Distributed"]["dp_degree"]
    sd = config["Distributed"]["sharding"]["sharding_degree"]
    g = config["Global"]
    gbs, lbs = g.get("global_batch_size"), g.get("local_batch_size")
    if gbs is None and lbs is None:
        raise ValueError("global_batch_size or local_batch_size should be set.")
    if gbs is not None and lbs is not None:
        assert gbs // lbs == dp * sd, (f"global_batch_size[{gbs}] should be divided by local_batch_size[{lbs}] when dp_degree is [{dp}] and "
                                       f"sharding_degree is [{sd}]")
    elif gbs is not None:
        assert gbs % (dp * sd) == 0, f"global_batch_size[{gbs}] should be divided by dp_degree[{dp}] times sharding_degree[{sd}]"
        g["local_batch_size"] = gbs // (dp * sd)
    else:
        g["global_batch_size"] = lbs * dp * sd
    assert g["local_batch_size"] % g["micro_batch_size"] == 0


def is_fused_matmul_bias_supported() -> bool:
    """Bias (+ activation) in the GEMM epilogue: always there when the native sm_100a library is loaded."""
    from ...ops import _native

    return _native.available()


def process_fused_configs(config) -> None:
    fused = config.get("Fused")
    if fused and fused.get("tensor_fusion"):
        assert env.world_size() == config["Distributed"]["dp_degree"], "tensor_fusion only support single card train or data parallel train"


def process_inference_configs(config) -> None:
    inf = config.get("Inference")
    if inf is None:
        return
    if inf.get("model_dir") is None:
        inf["model_dir"] = config["Engine"]["save_load"]["output_dir"]
    if inf.get("mp_degree") is None:
        inf["mp_degree"] = config["Distributed"]["mp_degree"]


def process_model_configs(config) -> None:
    m = config["Model"]
    if m.get("use_recompute") and not m.get("recompute_granularity"):
        m["recompute_granularity"] = "full"
    if m.get("fused_linear") and not is_fused_matmul_bias_supported():
        m["fused_linear"] = False
        logger.warning("The flag fused_linear needs the native GEMM library (bias in the epilogue); it is switched off for this run.")


def process_optim_configs(config) -> None:
    config["Optimizer"]["multi_precision"] = bool((config["Engine"].get("mix_precision") or {}).get("enable", False))


def process_engine_configs(config) -> None:
    e = config["Engine"]
    if e.get("test_iters") is None:
        e["test_iters"] = e["eval_iters"] * 10
    e["accumulate_steps"] = config["Global"]["local_batch_size"] // config["Global"]["micro_batch_size"]


def process_configs(config):
    process_fused_configs(config)
    process_model_configs(config)
    process_optim_configs(config)
    process_inference_configs(config)
    return config
