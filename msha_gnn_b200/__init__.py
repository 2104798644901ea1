"""msha_gnn_b200 -- B200-native (sm_100a) graph-attention message passing + link scoring, a drop-in for the
torch nn.Module surface of Sienna12321/MSHA--GNN's GAT / MSHA ("Ours") / HGANE layers and LinkPredictor.

PyTorch provides device memory, streams and torch.distributed; all arithmetic of the hot path runs in the
hand-written CUDA kernels of ``csrc/`` behind the C-ABI declared in ``include/msha_b200.h``.
There is no CPU fallback: ops raise if the library is not built or a tensor is not on a CUDA device.
"""
from . import _lib                                   # noqa: F401
from .graph import Graph, as_graph                   # noqa: F401
from .intra import GroupLists                        # noqa: F401
from .layers import (GAT, GATConv, GATLinkModel, GraphAttentionLayer, GraphConvolution, HGANELayer, LinkPredictor,   # noqa: F401
                     Ours, OursLayer, OursLayer2, OursLayer3, Teacher_LinkPredictor, ablation1, ablation2,
                     ablation3, dense_attention, last_attention, MLP, KD_cosine, llp_distill_loss, GCN, GraphSAGE,
                     export_attention)
from .data import HigherDataset, read_flow_files, normalize_adjacency_matrix   # noqa: F401
from . import functional                             # noqa: F401
from .graphs import CapturedStep                     # noqa: F401

__all__ = ["Graph", "as_graph", "GroupLists", "GAT", "GATConv", "GATLinkModel", "GraphAttentionLayer",
           "GraphConvolution", "HGANELayer", "LinkPredictor", "Teacher_LinkPredictor", "Ours", "OursLayer", "OursLayer2",
           "OursLayer3", "ablation1", "ablation2", "ablation3", "functional", "MLP", "KD_cosine", "llp_distill_loss", "GCN",
           "GraphSAGE", "export_attention", "CapturedStep", "HigherDataset", "read_flow_files", "normalize_adjacency_matrix"]


def build(force: bool = False):
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    return _lib.build(force=force)
