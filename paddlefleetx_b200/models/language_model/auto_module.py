"""``GPTModuleAuto`` / ``GPTGenerationModuleAuto`` (reference gpt/auto/auto_module.py:26-145 builds a static-graph model with
``auto.shard_tensor`` annotations).  The eager hybrid model already carries its sharding in the TP layers, so the auto names map
onto the same implementations and only validate the mesh."""
from .generation_module import GPTGenerationModule
from .language_module import GPTModule


class GPTModuleAuto(GPTModule):
    pass


class GPTGenerationModuleAuto(GPTGenerationModule):
    pass
