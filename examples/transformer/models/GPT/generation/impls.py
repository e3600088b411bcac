"""Text generation pieces (reference examples/transformer/models/GPT/generation/impls.py): tokeniser, left padding and the
sampling loop live in ``GPTGenerationModule``; these wrappers expose them to the explicit scripts."""
from paddlefleetx_b200.models.language_model.generation_module import GPTGenerationModule


def build_module(config):
    return GPTGenerationModule(config)


def generate(module, text, seed=None):
    return module.generate(text, seed=seed)
