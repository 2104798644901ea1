"""Name-compatible view of LLP.py's model classes: ``from msha_gnn_b200.llp import *`` gives ``MLP``,
``LinkPredictor``, ``GraphAttentionLayer``, ``GAT`` (the ``forward(input, adj)`` teacher without a ``features``
parameter, LLP.py:148-168), ``Teacher_LinkPredictor`` and ``KD_cosine`` (LLP.py:34-198), plus ``llp_distill_loss`` --
the loss of one training step (LLP.py:230-237) with the gathers fused into the kernels."""
from .layers import MLP, LinkPredictor, GraphAttentionLayer, Teacher_LinkPredictor, KD_cosine, llp_distill_loss   # noqa: F401
from .layers import LLPGAT as GAT   # noqa: F401

__all__ = ["MLP", "LinkPredictor", "GraphAttentionLayer", "GAT", "Teacher_LinkPredictor", "KD_cosine", "llp_distill_loss"]
