"""Small module helpers kept for API parity (reference: pretorched/models/utils.py:81-87)."""
import torch.nn as nn


class Identity(nn.Module):
    """Drop-in replacement for ``last_linear`` when a caller wants pooled features (README.md:520-547)."""

    def forward(self, x):
        return x
