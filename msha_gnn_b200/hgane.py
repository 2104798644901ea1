"""Name-compatible view of HGANE.py: ``from msha_gnn_b200.hgane import *`` mirrors ``from HGANE import *``
(train.py:11), where the layer class is called ``GraphAttentionLayer`` (HGANE.py:11)."""
from .layers import HGANELayer as GraphAttentionLayer   # noqa: F401

__all__ = ["GraphAttentionLayer"]
