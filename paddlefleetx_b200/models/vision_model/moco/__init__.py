from .moco import MoCo, MoCoClassifier, MoCoV2Projector, concat_all_gather  # noqa: F401
