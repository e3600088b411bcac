from .gpt_tokenizer import GPTTokenizer  # noqa: F401
