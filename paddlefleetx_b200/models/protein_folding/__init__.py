from .evoformer import (DistEmbeddingsAndEvoformer, EmbeddingsAndEvoformer, EvoformerIteration, GatedAttention, MSAColumnAttention,  # noqa: F401
                        MSARowAttentionWithPairBias, OuterProductMean, Transition, TriangleAttention, TriangleMultiplication)
