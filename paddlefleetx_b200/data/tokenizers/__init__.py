"""Tokenizers (reference ppfleetx/data/tokenizers/__init__.py:15-18).  None of them downloads: ``from_pretrained`` takes local files."""
from .debertav2_tokenizer import DebertaV2Tokenizer, SPMTokenizer, debertav2_tokenize, get_debertav2_tokenizer  # noqa: F401
from .ernie_tokenizer import BasicTokenizer, ErnieTokenizer, WordpieceTokenizer, get_ernie_tokenizer  # noqa: F401
from .gpt_tokenizer import GPTChineseTokenizer, GPTTokenizer  # noqa: F401
from .t5_tokenizer import T5Tokenizer, get_t5_tokenizer, t5_tokenize  # noqa: F401
from .tokenization_utils_base import (BatchEncoding, PaddingStrategy, PreTrainedTokenizer, PreTrainedTokenizerBase, SpecialTokensMixin,  # noqa: F401
                                      TruncationStrategy)


def get_text_tokenizer(name, strict: bool = False):
    """Tokenizer of an Imagen text tower by encoder name (``t5-*`` / ``t5/t5-11b`` -> T5, ``*deberta*`` -> DeBERTa-v2; reference
    imagen/modeling.py:228-238).  Returns None when the vocabulary files are not on this machine, unless ``strict``."""
    if not name:
        return None
    low = str(name).lower()
    try:
        if "deberta" in low:
            return get_debertav2_tokenizer(name)
        if "t5" in low:
            return get_t5_tokenizer(name)
        raise ValueError(f"no tokenizer registered for text encoder {name!r} (expected a t5-* or *deberta* name)")
    except FileNotFoundError:
        if strict:
            raise
        return None
