"""Pipeline-parallel ERNIE (reference ernie/dygraph/hybrid_model.py:713-872): ``EmbeddingsPipe`` -> N x
``TransformerEncoderLayerPipe`` -> ``ErniePoolerPipe``; the heads and the loss live in the criterion owned by the last
stage (``ErniePretrainingCriterionPipe`` holds parameters).  Stage boundaries carry the hidden states; the attention mask
is rebuilt on every stage from the pad-id convention carried alongside (tuple activations)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ....parallel.pipeline import LayerDesc, PipelineLayer
from . import model as E


class EmbeddingsPipe(E.ErnieEmbeddings):
    def forward(self, input_ids, token_type_ids=None, attention_mask=None):
        return super().forward(input_ids, token_type_ids)


class TransformerEncoderLayerPipe(E.TransformerEncoderLayer):
    def forward(self, x, attn_mask=None):
        return super().forward(x, attn_mask)


class ErniePretrainingCriterionPipe(nn.Module):
    """Heads + loss on the last stage; ``forward(hidden, masked_positions, masked_lm_labels, next_sentence_labels)``."""

    def __init__(self, hidden, vocab_size, hidden_act, mp_group, init_std, binary_head, dtype, device):
        super().__init__()
        self.pooler = E.ErniePooler(hidden, init_std, dtype, device)
        self.cls = E.ErniePretrainingHeads(hidden, vocab_size, hidden_act, None, mp_group, init_std, binary_head, dtype, device)
        self.criterion = E.ErniePretrainingCriterion(binary_head, mp_group)

    def forward(self, hidden, masked_positions, masked_lm_labels, next_sentence_labels=None):
        scores, rel = self.cls(hidden, self.pooler(hidden), masked_positions)
        out = self.criterion(scores, rel, masked_lm_labels, next_sentence_labels)
        return out[0] if isinstance(out, tuple) else out


class ErnieForPretrainingPipe(PipelineLayer):
    def __init__(self, hcg, mp_group=None, vocab_size=40000, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, ffn_hidden_size=None,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02, use_recompute=False, binary_head=True, virtual_pp_degree=1, dtype=None, device=None, num_layers=None,
                 **unused):
        n = num_layers or num_hidden_layers
        ffn = ffn_hidden_size or 4 * hidden_size
        descs = [LayerDesc(EmbeddingsPipe, vocab_size, hidden_size, hidden_dropout_prob, max_position_embeddings, type_vocab_size, 3, 0, False,
                           initializer_range, mp_group, dtype, device)]
        for _ in range(n):
            descs.append(LayerDesc(TransformerEncoderLayerPipe, hidden_size, num_attention_heads, ffn, hidden_dropout_prob, hidden_act,
                                   attention_probs_dropout_prob, 0, False, mp_group, initializer_range, dtype, device))
        loss = ErniePretrainingCriterionPipe(hidden_size, vocab_size, hidden_act, mp_group, initializer_range, binary_head, dtype, device)
        super().__init__(layers=descs, loss_fn=loss, hcg=hcg, seg_method="layer:TransformerEncoderLayer", recompute_interval=1 if use_recompute else 0,
                         num_virtual_pipeline_stages=virtual_pp_degree if virtual_pp_degree and virtual_pp_degree > 1 else 1)
