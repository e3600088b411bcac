"""Collate functions under the reference's module path (ppfleetx/data/utils/batch_collate_fn.py:31-190); they live in
``data/sampler/collate.py`` next to the samplers here."""
from ..sampler.collate import (DataCollatorWithPadding, Dict, ErnieCollateData, Pad, Stack, Tuple, default_collate_fn, gpt_collate_fn,  # noqa: F401
                               gpt_eval_collate_fn, imagen_collate_fn)

collate_fn = default_collate_fn
