"""Protein-folding trunk: Evoformer with DAP / BP parallelism, template embedding, extra-MSA stack and the geometry / chemistry
libraries its features need (reference ppfleetx/models/protein_folding/)."""
from . import all_atom, common, quat_affine, r3, residue_constants  # noqa: F401
from .evoformer import (DistEmbeddingsAndEvoformer, EmbeddingsAndEvoformer, EvoformerIteration, GatedAttention, GlobalAttention,  # noqa: F401
                        MSAColumnAttention, MSAColumnGlobalAttention, MSARowAttentionWithPairBias, OuterProductMean, Transition,
                        TriangleAttention, TriangleMultiplication)
from .template import SingleTemplateEmbedding, TemplateEmbedding, TemplatePair  # noqa: F401
