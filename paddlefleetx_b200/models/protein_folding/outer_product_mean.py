"""``OuterProductMean`` under the reference's module name (ppfleetx/models/protein_folding/outer_product_mean.py:23-150)."""
from .evoformer import OuterProductMean  # noqa: F401
