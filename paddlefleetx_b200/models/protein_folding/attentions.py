"""Attention blocks of the folding trunk under the reference's module name (ppfleetx/models/protein_folding/attentions.py:35-729); the
implementations live next to the Evoformer block in ``evoformer.py``."""
from .evoformer import GatedAttention as Attention  # noqa: F401
from .evoformer import (GatedAttention, GlobalAttention, MSAColumnAttention, MSAColumnGlobalAttention, MSARowAttentionWithPairBias,  # noqa: F401
                        TriangleAttention, TriangleMultiplication)
