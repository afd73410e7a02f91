"""Condition-embedding specification (reference: medical_diffusion/models/embedders/cond_embedders.py:5-23)."""


class LabelEmbedder:
    """nn.Embedding(num_classes, emb_dim) lookup added to the time embedding (conv_blocks.py:16-18)."""

    def __init__(self, emb_dim=32, num_classes=2, act_name=("SWISH", {})):
        self.emb_dim = emb_dim
        self.num_classes = num_classes
