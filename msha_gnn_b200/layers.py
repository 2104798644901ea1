"""Drop-in ``torch.nn.Module`` surface of the reference layers, backed by the sm_100a kernels.

Same constructor arguments, parameter names (``state_dict`` keys), initialisation order (a seeded
run initialises identically) and ``forward`` outputs as the reference classes; the adjacency
argument may be the reference's dense float ``(N, M)`` tensor, an int64 ``(2, E)`` ``edge_index``
(with ``n_rows``), or a prebuilt :class:`msha_gnn_b200.Graph`.  CUDA only -- a CPU tensor raises.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as Fn
from .graph import Graph, as_graph
from .ops import ACT_ELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SIGMOID_RELU


def _gdp_column(gdp):
    return torch.tensor(list(gdp.values())).view(-1, 1)


def _split_a(a: torch.Tensor, d: int):
    """(2d', 1) attention vector -> ([1,d'] neighbour half a[:d'], [1,d'] self half a[d':])  Ours.py:64."""
    return a[:d, 0].reshape(1, d), a[d:, 0].reshape(1, d)


# =================================================================================================
# a-1  GraphAttentionLayer                                     GAT.py:6-35 (copies: Ours.py:112-141 ...)
# =================================================================================================
class GraphAttentionLayer(nn.Module):
    def __init__(self, in_features, out_features, dropout):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.dropout = dropout
        self.W = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)

    def forward(self, input, adj):
        graph = as_graph(adj, n_rows=input.shape[0], n_cols=self.out_features)
        if graph.n_rows != input.shape[0] or graph.n_cols != self.out_features:
            # the reference's torch.where(adj > 0, e, zero_vec) needs adj.shape == (N, out_features) (GAT.py:30)
            raise RuntimeError(f"adjacency shape {(graph.n_rows, graph.n_cols)} must equal "
                               f"(N, out_features) = {(input.shape[0], self.out_features)}")
        h = Fn.linear(input, self.W)                                            # GAT.py:21
        return Fn.gal(h, self.a, graph, 1, self.dropout, self.training)         # GAT.py:24-35


def _gal_heads(x, layers, graph, training):
    """Head-batched evaluation of several GraphAttentionLayers sharing input and adjacency (GAT.py:55)."""
    H = len(layers)
    if H == 1:
        return layers[0](x, graph)
    W = torch.cat([l.W for l in layers], dim=1)                                 # (F, H*M): one GEMM for all heads
    h = Fn.linear(x, W)
    a = torch.cat([l.a for l in layers], dim=0)                                 # only to route the (zero) gradient
    return Fn.gal(h, a, graph, H, layers[0].dropout, training)


# =================================================================================================
# a-2  GAT                                                      GAT.py:38-58, LLP.py:148-168
# =================================================================================================
class GAT(nn.Module):
    owns_features = True      # GAT.py:41-42 registers `features`; LLP.py:151-152 has that line commented out

    def __init__(self, n_features, n_classes, n_heads, dropout, gdp, N):
        super().__init__()
        if self.owns_features:
            gdp_values = _gdp_column(gdp)
            self.features = nn.Parameter(torch.cat((torch.rand([N, n_features])[:, :-1], gdp_values), dim=1))
        self.n_classes = n_classes
        self.n_heads = n_heads
        self.dropout = dropout
        self.attentions = [GraphAttentionLayer(n_features, n_classes, dropout=dropout) for _ in range(n_heads)]
        for i, attention in enumerate(self.attentions):
            self.add_module('attention_{}'.format(i), attention)
        self.out_att = GraphAttentionLayer(n_features * n_heads, n_classes, dropout=dropout)

    def forward(self, *args):
        """``forward(adj)`` (GAT.py:53) or ``forward(input, adj)`` (LLP.py:163)."""
        if len(args) == 1:
            if not self.owns_features:
                raise TypeError("LLP.GAT.forward takes (input, adj)")
            x_in, adj = self.features, args[0]
        elif len(args) == 2:
            x_in, adj = args
        else:
            raise TypeError("GAT.forward takes (adj) or (input, adj)")
        graph = as_graph(adj, n_rows=x_in.shape[0], n_cols=self.n_classes)
        x = Fn.dropout(x_in, self.dropout, self.training)                      # GAT.py:54
        x = _gal_heads(x, self.attentions, graph, self.training)               # GAT.py:55
        x = Fn.dropout(x, self.dropout, self.training)                         # GAT.py:56
        x = self.out_att(x, graph)                                             # inner ELU GAT.py:35
        return Fn.log_softmax(x, pre_elu=True)                                 # outer ELU + log_softmax GAT.py:57-58


class LLPGAT(GAT):
    """``LLP.GAT`` (LLP.py:148-168): same layers, no ``features`` parameter (no ``torch.rand`` draw at construction, no
    such ``state_dict`` key), ``forward(input, adj)`` only.  ``msha_gnn_b200.llp.GAT`` is this class."""
    owns_features = False


# =================================================================================================
# a-3/a-4  MSHA layers                                          Ours.py:29-109, Ablation.py:10-277
# =================================================================================================
class _OursBase(nn.Module):
    """Parameters and init order of OursLayer / OursLayer2 / OursLayer3 (identical __init__, Ours.py:30-52)."""
    variant = 1

    def __init__(self, in_features, out_features, dropout):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = 0.2
        self.dropout = dropout
        self.W1 = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        self.W2 = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.W1.data, gain=1.414)
        nn.init.xavier_uniform_(self.W2.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)
        self.a3 = nn.Parameter(torch.zeros(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a3.data, gain=1.414)
        self.a4 = nn.Parameter(torch.zeros(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a4.data, gain=1.414)
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.bn1 = nn.BatchNorm1d(out_features)
        self.bn2 = nn.BatchNorm1d(out_features)
        self.bn3 = nn.BatchNorm1d(out_features)      # registered but unused, as in the reference

    def forward(self, Sinput, Rinput, inter_adj, city_adj=None, province_adj=None, source_index=None, record=False,
                Coeff12=None, Coeff3=None, Coeff4=None):
        out = msha_heads_forward([self], Sinput, Rinput, inter_adj, city_adj, province_adj, source_index,
                                 self.training, record=record, Coeff3=Coeff3, Coeff4=Coeff4)
        return out


class OursLayer(_OursBase):
    variant = 1


class OursLayer2(_OursBase):
    variant = 2


class OursLayer3(_OursBase):
    variant = 3


def _bn_heads(x, bns, training):
    """BatchNorm1d + LeakyReLU over head-concatenated columns; running stats are split back per head."""
    if len(bns) == 1:
        bn = bns[0]
        y = Fn.bn_lrelu(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, bn.momentum, bn.eps)
        if training:
            bn.num_batches_tracked += 1
        return y
    gamma = torch.cat([b.weight for b in bns])
    beta = torch.cat([b.bias for b in bns])
    rm = torch.cat([b.running_mean for b in bns])
    rv = torch.cat([b.running_var for b in bns])
    y = Fn.bn_lrelu(x, gamma, beta, rm, rv, training, bns[0].momentum, bns[0].eps)
    if training:
        d = bns[0].num_features
        with torch.no_grad():
            for i, b in enumerate(bns):
                b.running_mean.copy_(rm[i * d:(i + 1) * d])
                b.running_var.copy_(rv[i * d:(i + 1) * d])
                b.num_batches_tracked += 1
    return y


last_attention = {}   # filled when record=True: the export of train.py:284-291 / Ours.py:92-96


def msha_heads_forward(layers, Sinput, Rinput, inter_adj, city_adj, province_adj, source_index, training,
                       record=False, Coeff3=None, Coeff4=None):
    """Head-batched OursLayer{,2,3}.forward: returns (N, H*M) -- the per-head (N, M) outputs concatenated
    (Ours.py:164).  One GEMM per side, one fused attention pass and one BN pass for all heads."""
    H = len(layers)
    variant = layers[0].variant
    d = layers[0].out_features
    p = layers[0].dropout
    graph = as_graph(inter_adj, n_rows=Sinput.shape[0], n_cols=Rinput.shape[0])
    N, M = graph.n_rows, graph.n_cols
    if (N, M) != (Sinput.shape[0], Rinput.shape[0]):
        raise RuntimeError(f"inter adjacency {(N, M)} does not match Sinput/Rinput rows {(Sinput.shape[0], Rinput.shape[0])}")
    if H == 1:
        W1, W2 = layers[0].W1, layers[0].W2
    else:
        W1 = torch.cat([l.W1 for l in layers], dim=1)
        W2 = torch.cat([l.W2 for l in layers], dim=1)
    h1 = Fn.linear(Rinput, W1)                                                  # (M, H*d')  Ours.py:57
    h2 = Fn.linear(Sinput, W2)                                                  # (N, H*d')  Ours.py:58
    a_nbr = torch.cat([_split_a(l.a, d)[0] for l in layers], dim=0)            # (H, d')
    a_self = torch.cat([_split_a(l.a, d)[1] for l in layers], dim=0)
    s_nbr = Fn.node_scores(h1, a_nbr, None, H, d)                               # a[:d'].h1[j]   Ours.py:64
    s_self = Fn.node_scores(h2, a_self, None, H, d)                             # a[d':].h2[i]
    u_in, v_in, alpha = Fn.attention_block(graph, s_nbr, s_self, h1, h2, heads=H, act=ACT_NONE, dropout_p=p,
                                           training=training, want_cols=True)   # Ours.py:65-69,98,100
    if variant != 3:
        from .intra import intra_scales
        if source_index is None or city_adj is None or province_adj is None:
            raise ValueError("OursLayer / OursLayer2 need city_adj, province_adj and source_index")
        intra_nc, coeffs = intra_scales(layers, graph, h2, alpha, city_adj, province_adj, source_index, training,
                                        joint=(variant == 1), want_coeffs=record)
        u_in = u_in + intra_nc                                                  # Ours.py:99,101
        if record:
            last_attention.update(coeffs)
            if Coeff3 is not None:
                Coeff3[source_index] = coeffs["Coeff3"]
            if Coeff4 is not None:
                Coeff4[source_index] = coeffs["Coeff4"]
    if record:
        # per-head layout (H, N, M); the reference's heads overwrite one global (train.Coeff12new, Ours.py:94), so what its
        # explainer sees is the LAST head: exported under the reference's name as well
        last_attention["Coeff12"] = dense_attention(graph, alpha, H)
        last_attention["Coeff12new"] = last_attention["Coeff12"][-1]
    v = _bn_heads(v_in, [l.bn1 for l in layers], training)                      # Ours.py:100
    u = _bn_heads(u_in, [l.bn2 for l in layers], training)                      # Ours.py:101
    return Fn.matmul_nt_act(u, v, H, ACT_ELU)                                   # Ours.py:108-109


def dense_attention(graph: Graph, alpha, heads=1):
    """Dense (H, N, M) view of the per-edge attention (export path only; Ours.py:92-96)."""
    rp, col = graph.attention_csr()
    deg = (rp[1:] - rp[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(graph.n_rows, device=alpha.device), deg)
    cols = torch.where(col < 0, ~col, col).long()
    out = torch.zeros((heads, graph.n_rows, graph.n_cols), dtype=alpha.dtype, device=alpha.device)
    out[:, rows, cols] = alpha.detach().t()
    return out


class _MshaModel(nn.Module):
    """Ours / ablation2 / ablation3 (Ours.py:144-167, Ablation.py:208-231,279-301)."""
    layer_cls = OursLayer

    def __init__(self, in_features, out_features, n_classes, n_heads, dropout, gdp, Scount, Rcount):
        super().__init__()
        gdp_values = _gdp_column(gdp)
        self.Sfeatures = nn.Parameter(torch.cat((torch.rand([Scount, in_features])[:, :-1], gdp_values), dim=1))
        self.Rfeatures = nn.Parameter(torch.rand([Rcount, in_features]))
        self.n_classes = n_classes
        self.n_heads = n_heads
        self.dropout = dropout
        self.attentions = [self.layer_cls(in_features, out_features, dropout=dropout) for _ in range(n_heads)]
        for i, attention in enumerate(self.attentions):
            self.add_module('attention_{}'.format(i), attention)
        self.out_att = GraphAttentionLayer(n_classes * n_heads, n_classes, dropout=dropout)

    def forward(self, inter_adj, city_adj=None, province_adj=None, source_index=None, record=False, Coeff12=None,
                Coeff3=None, Coeff4=None):
        graph = as_graph(inter_adj, n_rows=self.Sfeatures.shape[0], n_cols=self.Rfeatures.shape[0])
        s_input = Fn.dropout(self.Sfeatures, self.dropout, self.training)       # Ours.py:161
        r_input = Fn.dropout(self.Rfeatures, self.dropout, self.training)       # Ours.py:162
        x = msha_heads_forward(self.attentions, s_input, r_input, graph, city_adj, province_adj, source_index,
                               self.training, record=record, Coeff3=Coeff3, Coeff4=Coeff4)   # Ours.py:164
        x = Fn.dropout(x, self.dropout, self.training)                          # Ours.py:165
        x = self.out_att(x, graph)                                              # Ours.py:166 (inner ELU)
        return Fn.log_softmax(x, pre_elu=True)                                  # Ours.py:166-167


class Ours(_MshaModel):
    layer_cls = OursLayer


class ablation2(_MshaModel):
    layer_cls = OursLayer2


class ablation3(_MshaModel):
    layer_cls = OursLayer3


class ablation1(nn.Module):
    """Single OursLayer, no out_att (Ablation.py:118-136)."""

    def __init__(self, in_features, out_features, n_classes, n_heads, dropout, gdp, Scount, Rcount):
        super().__init__()
        gdp_values = _gdp_column(gdp)
        self.Sfeatures = nn.Parameter(torch.cat((torch.rand([Scount, in_features])[:, :-1], gdp_values), dim=1))
        self.Rfeatures = nn.Parameter(torch.rand([Rcount, in_features]))
        self.n_classes = n_classes
        self.n_heads = n_heads
        self.dropout = dropout
        self.attention = OursLayer(in_features, out_features, dropout=dropout)

    def forward(self, inter_adj, city_adj, province_adj, source_index):
        s_input = Fn.dropout(self.Sfeatures, self.dropout, self.training)
        r_input = Fn.dropout(self.Rfeatures, self.dropout, self.training)
        x = self.attention(s_input, r_input, inter_adj, city_adj, province_adj, source_index)
        x = Fn.dropout(x, self.dropout, self.training)
        return Fn.log_softmax(x, pre_elu=True)                                  # Ablation.py:135-136


# =================================================================================================
# a-6  HGANE.GraphAttentionLayer                                                HGANE.py:11-76
# =================================================================================================
class HGANELayer(nn.Module):
    """Batch-subgraph GAT of HGANE.py (the class is called ``GraphAttentionLayer`` there; see ``msha_gnn_b200.hgane``).

    Inter attention (B x M) is a plain row softmax; the intra attention (B x B, genuine pairwise logits) is
    normalised by the *joint* sum ``sum exp(intra) + sum exp(inter)`` (HGANE.py:61-64), i.e.
    ``att_intra = softmax_intra * sigmoid(lse_intra - lse_inter)`` -- both softmaxes run in the fused attention kernel,
    which also returns the rows' log-sum-exp."""

    def __init__(self, in_features, out_features, Scount, Rcount, gdp, dropout=0.5):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.dropout = dropout
        self.alpha = 0.2
        gdp_values = _gdp_column(gdp)
        self.features = nn.Parameter(torch.cat((torch.rand([Scount, self.in_features])[:, :-1], gdp_values), dim=1))
        self.source_embedding = nn.Parameter(torch.rand([Scount, self.in_features]))
        self.recipient_embedding = nn.Parameter(torch.rand([Rcount, self.in_features]))
        self.W1 = nn.Linear(in_features, out_features, bias=False)
        self.W2 = nn.Linear(in_features, out_features, bias=False)
        self.a12 = nn.Linear(2 * out_features, 1, bias=False)
        self.a3 = nn.Linear(2 * out_features, 1, bias=False)
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.bn1 = nn.BatchNorm1d(out_features)
        self.bn2 = nn.BatchNorm1d(out_features)
        nn.init.xavier_uniform_(self.W1.weight)
        nn.init.xavier_uniform_(self.W2.weight)
        nn.init.xavier_uniform_(self.a12.weight)
        nn.init.xavier_uniform_(self.a3.weight)

    def forward(self, adj_inter, adj_intra, source_index):
        if not adj_inter.is_cuda:
            raise RuntimeError("msha_b200 is CUDA-only: move the adjacency to the GPU (no CPU fallback)")
        src = source_index
        d = self.out_features
        g_inter = Graph.from_dense(adj_inter[src])                              # (B, M)   HGANE.py:39
        g_intra = Graph.from_dense(adj_intra[src[:, None], src])                # (B, B)   HGANE.py:38
        g_inter.isolated = g_intra.isolated = "zero"          # no -9e15 masking here: exp() of the mask is 0 (HGANE.py:61)
        E_r = self.recipient_embedding
        E_s = self.source_embedding.index_select(0, src)
        h1 = Fn.linear_bias_act(E_r, self.W1.weight, None, ACT_NONE)            # HGANE.py:40
        h2 = Fn.linear_bias_act(E_s, self.W2.weight, None, ACT_NONE)            # HGANE.py:41
        a12, a3 = self.a12.weight, self.a3.weight
        s12_nbr = Fn.node_scores(h1, a12[:, :d], None, 1, d)                    # a12[:d'].h1[j]   HGANE.py:46-47
        s12_self = Fn.node_scores(h2, a12[:, d:], None, 1, d)                   # a12[d':].h2[i]
        s3_self = Fn.node_scores(h2, a3[:, :d], None, 1, d)                     # a3[:d'].h2[i]    HGANE.py:49-52
        s3_nbr = Fn.node_scores(h2, a3[:, d:], None, 1, d)                      # a3[d':].h2[j]
        agg_inter, vin_raw, _, lse12 = Fn.attention_block(g_inter, s12_nbr, s12_self, E_r, E_s, heads=1,
                                                          dropout_p=self.dropout, training=self.training,
                                                          want_cols=True, want_lse=True)   # HGANE.py:66-68
        agg_intra, _, lse3 = Fn.attention_block(g_intra, s3_nbr, s3_self, E_s, heads=1, dropout_p=self.dropout,
                                                training=self.training, want_lse=True)
        rho = torch.sigmoid(lse3 - lse12)                                       # S_intra / (S_intra + S_inter)  HGANE.py:61-64
        u_in = Fn.linear_bias_act(agg_inter, self.W1.weight, None, ACT_NONE) \
            + Fn.linear_bias_act(rho * agg_intra, self.W2.weight, None, ACT_NONE)               # HGANE.py:71-72
        v_in = Fn.linear_bias_act(vin_raw, self.W1.weight, None, ACT_NONE)                      # HGANE.py:73
        u = _bn_heads(u_in, [self.bn1], self.training)
        v = _bn_heads(v_in, [self.bn2], self.training)
        out = Fn.matmul_nt_act(u, v, 1, ACT_ELU)                                                # HGANE.py:75-76
        # a batch row without inter neighbour is 0/0 in the reference (HGANE.py:67-68) and poisons v, hence everything
        poison = torch.where((g_inter.degrees == 0).any(), float("nan"), 1.0)
        return out * poison


# =================================================================================================
# a-7  LinkPredictor / Teacher_LinkPredictor                                   LLP.py:86-115,170-198
# =================================================================================================
class LinkPredictor(nn.Module):
    def __init__(self, predictor, in_channels, hidden_channels, out_channels, num_layers, dropout):
        super().__init__()
        self.predictor = predictor
        self.lins = nn.ModuleList()
        self.lins.append(nn.Linear(in_channels, hidden_channels))
        for _ in range(num_layers - 2):
            self.lins.append(nn.Linear(hidden_channels, hidden_channels))
        self.lins.append(nn.Linear(hidden_channels, out_channels))   # allocated, never applied (LLP.py:111)
        self.dropout = dropout
        self.fused = True        # use the fused tensor-core scorer when the shape allows (False -> unfused kernels)

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()

    def forward(self, x_i, x_j):
        """Reference signature: rows already gathered by the caller (``h[source_index]``, LLP.py:233)."""
        return self._score(x_i, x_j, None, None)

    def forward_pairs(self, h_i, h_j, src, dst):
        """Fused pair-gather entry: scores pairs (src[p], dst[p]) of the embedding tables h_i / h_j."""
        return self._score(h_i, h_j, src, dst)

    def nll_loss_pairs(self, h_i, h_j, src, dst, target):
        """``F.nll_loss(self.forward_pairs(h_i, h_j, src, dst), target)`` (the read-out of LLP.py:233-235) with the loss
        folded into the fused scorer: d scores is generated inside the backward kernel instead of being streamed through
        HBM.  Falls back to the two separate ops whenever the fused scorer does not apply."""
        if self.predictor == 'mlp':
            hidden = list(self.lins)[:-1]
            drop_on = self.training and self.dropout > 0
            if (self.fused and len(hidden) == 1 and not drop_on
                    and Fn.score_mlp_nll_supported(h_i, h_j, hidden[0].weight)):
                loss, _ = Fn.score_mlp_nll(h_i, h_j, src, dst, hidden[0].weight, hidden[0].bias, target, ACT_SIGMOID_RELU)
                return loss
        return Fn.nll_loss(self._score(h_i, h_j, src, dst), target)

    def _score(self, hi, hj, src, dst):
        if self.predictor == 'mlp':
            hidden = list(self.lins)[:-1]
            drop_on = self.training and self.dropout > 0
            if (self.fused and len(hidden) == 1 and not drop_on
                    and Fn.score_mlp_supported(hi, hj, hidden[0].weight)):
                # one tensor-core kernel: gather, Hadamard, Linear, ReLU, sigmoid (LLP.py:105-115)
                return Fn.score_mlp(hi, hj, src, dst, hidden[0].weight, hidden[0].bias, ACT_SIGMOID_RELU)
            x = Fn.pair_mul(hi, hj, src, dst)                                   # LLP.py:105
            for k, lin in enumerate(hidden):
                last = k == len(hidden) - 1
                drop = self.training and self.dropout > 0
                if last and not drop:
                    x = Fn.linear_bias_act(x, lin.weight, lin.bias, ACT_SIGMOID_RELU)   # relu + sigmoid fused
                else:
                    x = Fn.linear_bias_act(x, lin.weight, lin.bias, ACT_RELU)           # LLP.py:108-109
                    x = Fn.dropout(x, self.dropout, self.training)                      # LLP.py:110
                    if last:
                        x = Fn._Act.apply(x, ACT_SIGMOID)                               # LLP.py:115
            if not hidden:
                x = Fn._Act.apply(x, ACT_SIGMOID)
            return x
        if self.predictor == 'inner':
            return Fn.pair_dot(hi, hj, src, dst, ACT_SIGMOID)                   # LLP.py:112-115
        # any other predictor string: the reference applies sigmoid to x_i * x_j
        return Fn._Act.apply(Fn.pair_mul(hi, hj, src, dst), ACT_SIGMOID)


class Teacher_LinkPredictor(LinkPredictor):
    pass


# =================================================================================================
# model.GraphConvolution / GCN                                                 model.py:11-64
# =================================================================================================
class GraphConvolution(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.rand([in_features, out_features]))
        if bias:
            self.bias = nn.Parameter(torch.tensor(out_features, dtype=torch.float32))   # 0-dim, model.py:23
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, input, adj, values=None, transposed=False):
        """``adj.T @ (input @ W) + bias`` (model.py:36-39).  ``adj`` dense/Graph; its stored values weight
        the edges (pass ``values`` to override, e.g. ``graph.normalized_values()``).  ``transposed=True`` evaluates the
        layer on ``adj.t()`` -- ``adj @ (input @ W) + bias``, GCN's second layer (model.py:61) -- over the same graph
        instead of building a second one from the transposed dense matrix."""
        graph = as_graph(adj)
        support = Fn.linear(input, self.weight)
        w = graph.val if values is None else values
        out = Fn.spmm(graph, w, support) if transposed else Fn.spmm_t(graph, w, support)
        return out + self.bias if self.bias is not None else out

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


# =================================================================================================
# generic multi-head GAT layer for the synthetic configs (SURVEY.md section 8a generalisation rule:
# a-3's inter-scale block with S = R, Ablation.py:262-271 + alpha @ h1 :274)
# =================================================================================================
class GATConv(nn.Module):
    def __init__(self, in_features, out_features, heads=1, concat=True, dropout=0.0, activation="elu"):
        super().__init__()
        self.in_features, self.out_features, self.heads, self.concat = in_features, out_features, heads, concat
        self.dropout = dropout
        self.activation = activation
        self.W = nn.Parameter(torch.zeros(size=(in_features, heads * out_features)))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.a_nbr = nn.Parameter(torch.zeros(size=(heads, out_features)))
        self.a_self = nn.Parameter(torch.zeros(size=(heads, out_features)))
        nn.init.xavier_uniform_(self.a_nbr.data, gain=1.414)
        nn.init.xavier_uniform_(self.a_self.data, gain=1.414)

    def forward(self, x, adj, return_alpha=False):
        graph = as_graph(adj, n_rows=x.shape[0], n_cols=x.shape[0])
        Wh = Fn.linear(x, self.W)
        s_nbr, s_self = Fn.node_scores(Wh, self.a_nbr, self.a_self, self.heads, self.out_features)
        fuse_elu = self.activation == "elu" and self.concat
        out, alpha = Fn.attention_block(graph, s_nbr, s_self, Wh, heads=self.heads,
                                        act=ACT_ELU if fuse_elu else ACT_NONE, dropout_p=self.dropout,
                                        training=self.training)
        if not self.concat:
            out = out.view(x.shape[0], self.heads, self.out_features).mean(dim=1)
            if self.activation == "elu":
                out = Fn.elu(out)
        return (out, alpha) if return_alpha else out


class GATLinkModel(nn.Module):
    """L-layer multi-head GAT encoder + LinkPredictor scorer (BASELINE.json configs[2]/[3])."""

    def __init__(self, in_features, hidden, heads, num_layers=2, predictor_hidden=None, dropout=0.0):
        super().__init__()
        assert hidden % heads == 0
        dims = [in_features] + [hidden] * num_layers
        self.convs = nn.ModuleList(GATConv(dims[i], hidden // heads, heads, concat=True, dropout=dropout)
                                   for i in range(num_layers))
        self.predictor = LinkPredictor('mlp', hidden, predictor_hidden or hidden, 1, 2, dropout)
        self.dropout = dropout

    def encode(self, x, graph):
        for conv in self.convs:
            x = conv(x, graph)
        return x

    def forward(self, x, graph, src, dst):
        h = self.encode(x, graph)
        return self.predictor.forward_pairs(h, h, src, dst)

    def loss(self, x, graph, src, dst, labels):
        """nll read-out of the pair scores (LLP.py:233-235) with the fused scorer + loss backward."""
        h = self.encode(x, graph)
        return self.predictor.nll_loss_pairs(h, h, src, dst, labels)


# =================================================================================================
# the LLP student (LLP.py:36-84) and its distillation step (LLP.py:217-248)
# =================================================================================================
class MLP(nn.Module):
    """``LLP.MLP`` (LLP.py:36-84): Linear (+ norm) + ReLU + dropout per hidden layer, plain Linear last.  Each
    ``relu(layer(h))`` is one tensor-core GEMM with the bias and ReLU in its epilogue."""

    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, dropout_ratio, norm_type="none"):
        super().__init__()
        self.num_layers = num_layers
        self.norm_type = norm_type
        self.dropout = nn.Dropout(dropout_ratio)
        self.layers = nn.ModuleList()
        self.norms = nn.ModuleList()
        if num_layers == 1:
            self.layers.append(nn.Linear(input_dim, output_dim))
        else:
            self.layers.append(nn.Linear(input_dim, hidden_dim))
            self._add_norm(hidden_dim)
            for _ in range(num_layers - 2):
                self.layers.append(nn.Linear(hidden_dim, hidden_dim))
                self._add_norm(hidden_dim)
            self.layers.append(nn.Linear(hidden_dim, output_dim))

    def _add_norm(self, dim):
        if self.norm_type == "batch":
            self.norms.append(nn.BatchNorm1d(dim))
        elif self.norm_type == "layer":
            self.norms.append(nn.LayerNorm(dim))

    def reset_parameters(self):
        for layer in self.layers:
            layer.reset_parameters()

    def forward(self, feats):
        h = feats
        p = self.dropout.p
        for l, layer in enumerate(self.layers):
            last = l == self.num_layers - 1
            if last:
                h = Fn.linear_bias_act(h, layer.weight, layer.bias, ACT_NONE)          # LLP.py:78
            elif self.norm_type == "none":
                h = Fn.linear_bias_act(h, layer.weight, layer.bias, ACT_RELU)          # LLP.py:78,82 fused
                h = Fn.dropout(h, p, self.training)                                     # LLP.py:83
            else:
                # norm_type 'batch' / 'layer' is never selected by the reference's call site (LLP.py:291 keeps the default
                # "none"); the normalisation itself stays a torch module, the Linear / ReLU / dropout are this library's
                h = Fn.linear_bias_act(h, layer.weight, layer.bias, ACT_NONE)
                h = self.norms[l](h)
                h = Fn._Act.apply(h, ACT_RELU)
                h = Fn.dropout(h, p, self.training)
        return h


def KD_cosine(s, t):
    """``LLP.KD_cosine`` (LLP.py:34-35) on already gathered rows."""
    return Fn.kd_cosine(s, t)


def llp_distill_loss(model, predictor, teacher_model, teacher_predictor, features, adj_inter, source_index,
                     recipient_index, True_label=10.0, KD_f=0.1, KD_p=100.0):
    """The loss of one LLP training step (LLP.py:230-237), every gather fused into its consumer:

        h = model(features); t_h = teacher_model(features, adj_inter)
        output = predictor(h[source_index], h[recipient_index])          -> fused pair-gather scorer
        label_loss = F.nll_loss(output, recipient_index)
        t_out = teacher_predictor(t_h[source_index], t_h[recipient_index]).detach()
        loss = True_label * label_loss + KD_f * KD_cosine(h[source_index], t_h[source_index]) + KD_p * mse(output, t_out)

    Returns ``(loss, parts)`` with the three un-weighted terms.  The teacher branch carries no gradient: LLP.py:297 hands
    only the student's and the predictor's parameters to the optimiser and detaches both teacher outputs."""
    h = model(features)
    with torch.no_grad():
        t_h = teacher_model(features, adj_inter)
        t_out = teacher_predictor.forward_pairs(t_h, t_h, source_index, recipient_index)
    output = predictor.forward_pairs(h, h, source_index, recipient_index)
    if output.dim() == 2 and output.shape[1] == 1:
        output = output.squeeze(1)                                              # .squeeze() LLP.py:233
    label_loss = Fn.nll_loss(output, recipient_index)                           # LLP.py:235
    kd_f = Fn.kd_cosine(h, t_h, source_index, source_index)                     # LLP.py:236
    kd_p = Fn.mse_loss(output, t_out.reshape(output.shape))                     # LLP.py:237
    loss = True_label * label_loss + KD_f * kd_f + KD_p * kd_p
    return loss, {"label_loss": label_loss, "KD_cosine": kd_f, "mse_loss": kd_p}


# =================================================================================================
# the remaining baselines -- GCN (model.py:48-64) and GraphSAGE (SGAE.py:41-56)
# =================================================================================================
class GCN(nn.Module):
    def __init__(self, nfeat, nhid, nclass, dropout, gdp, N):
        super().__init__()
        gdp_values = _gdp_column(gdp)
        self.features = nn.Parameter(torch.cat((torch.rand([N, nfeat])[:, :], gdp_values), dim=1))   # model.py:52
        self.gc1 = GraphConvolution(nfeat + 1, nhid)
        self.gc2 = GraphConvolution(nhid, nhid)
        self.gc3 = GraphConvolution(nhid, nclass)        # allocated, never applied (model.py:62-63)
        self.dropout = dropout

    def forward(self, adj):
        graph = as_graph(adj)
        x = Fn._Act.apply(self.gc1(self.features, graph), ACT_RELU)             # (M, nhid)  model.py:59
        x = Fn.dropout(x, self.dropout, self.training)                          # model.py:60
        x = Fn._Act.apply(self.gc2(x, graph, transposed=True), ACT_RELU)        # gc2(x, adj.t()) -> (N, nhid)  model.py:61
        return Fn.log_softmax(x, pre_elu=False)                                 # model.py:64


class GraphSAGE(nn.Module):
    """SGAE.py:41-56.  ``Scount`` is a module-level global there (SGAE.py:46); here it is a constructor argument."""

    def __init__(self, in_features, hidden_features, out_features, gdp, Scount=None):
        super().__init__()
        gdp_values = _gdp_column(gdp)
        Scount = gdp_values.shape[0] if Scount is None else Scount
        self.Sfeatures = nn.Parameter(torch.cat((torch.rand([Scount, in_features])[:, :-1], gdp_values), dim=1))
        self.linear1 = nn.Linear(in_features, hidden_features)
        self.linear2 = nn.Linear(hidden_features, out_features)

    def forward(self, source_index, adj, values=None):
        graph = as_graph(adj)
        x = self.Sfeatures.index_select(0, source_index)                                        # SGAE.py:51
        x = Fn.linear_bias_act(x, self.linear1.weight, self.linear1.bias, ACT_RELU)             # SGAE.py:52
        x = Fn.csr_rows_mul(graph, x, source_index, values)                                     # SGAE.py:53
        x = Fn.linear_bias_act(x, self.linear2.weight, self.linear2.bias, ACT_RELU)             # SGAE.py:54
        return Fn.log_softmax(x, pre_elu=False)                                                 # SGAE.py:55


# =================================================================================================
# attention export without the dense (N, M) / (N, N) dumps
# (train.py:284-321 fills Coeff12 / Coeff3 / Coeff4, Explainer.py:25-30 takes argwhere(row == max(row)))
# =================================================================================================
def export_attention(graph: Graph, alpha, heads=1, head=-1):
    """Sparse form of ``Coeff12`` and of what the explainer extracts from it.

    Returns a dict: ``edge_index`` (2, E) int64 and ``alpha`` (E,) -- the attention of ``head`` (mean over heads for
    ``head < 0``) in canonical row-major order; ``row_max`` / ``row_argmax`` / ``row_ties`` per source row (``interAttS``:
    the recipient a source attends to most, smallest column on ties, and how many columns tie); ``col_max`` /
    ``col_argmax`` / ``col_ties`` per recipient column (``interAttR``).  ``argmax`` is -1 for an empty row / column."""
    rp, col = graph.attention_csr()
    colptr, rowidx, perm = graph.attention_csc()
    a = alpha.detach()
    vmax_r, first_r, ties_r = Fn.segment_argmax(rp, a, heads, head)
    vmax_c, first_c, ties_c = Fn.segment_argmax(colptr, a, heads, head, perm=perm)
    cols = torch.where(col < 0, ~col, col).long()
    deg = (rp[1:] - rp[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(graph.n_rows, device=a.device), deg)
    vals = a.view(-1, heads)
    vals = vals[:, head] if head >= 0 else vals.mean(dim=1)
    neg_r = torch.full((graph.n_rows,), -1, dtype=torch.int64, device=a.device)
    neg_c = torch.full((graph.n_cols,), -1, dtype=torch.int64, device=a.device)
    if cols.numel():
        row_arg = torch.where(first_r >= 0, cols[first_r.clamp(min=0).long()], neg_r)
        col_arg = torch.where(first_c >= 0, rowidx.long()[first_c.clamp(min=0).long()], neg_c)
    else:
        row_arg, col_arg = neg_r, neg_c
    return {"edge_index": torch.stack([rows, cols]), "alpha": vals,
            "row_max": vmax_r, "row_argmax": row_arg, "row_ties": ties_r,
            "col_max": vmax_c, "col_argmax": col_arg, "col_ties": ties_c}
