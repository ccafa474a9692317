"""Superpoint transformer + coarse matching on the CUDA path -- mirrors of the reference modules
(same class names, constructor arguments and state_dict keys):

  geotransformer/modules/geotransformer/geotransformer.py   GeometricStructureEmbedding, GeometricTransformer
  geotransformer/modules/transformer/rpe_transformer.py      RPEMultiHeadAttention, RPEAttentionLayer, RPETransformerLayer
  geotransformer/modules/transformer/vanilla_transformer.py  MultiHeadAttention, AttentionLayer, TransformerLayer
  geotransformer/modules/transformer/output_layer.py         AttentionOutput
  geotransformer/modules/transformer/conditional_transformer.py  RPEConditionalTransformer
  geotransformer/modules/transformer/positional_embedding.py SinusoidalPositionalEmbedding
  geotransformer/modules/geotransformer/superpoint_matching.py  SuperPointMatching

Block lists of SE3ET-I / I2 ('self_eq', 'cross') and of SE3ET-E / E2 ('self_eq', 'cross_a_soft', 'cross_r_soft',
'self', invariant 'cross'; experiments/se3ete.3dmatch/config.py:194) run on the CUDA path.

Execution model: all clouds of all pairs are stored flat, ordered [ref_0..ref_{P-1}, src_0..src_{P-1}], equivariant
states as bf16 (T*A, C) rows in (point, anchor) order.  The (N, N, C) geometric embedding is built once per forward
(bf16) and `q . proj_p(embedding)` is evaluated as `embedding[n] @ (W_p^T q[n])`, so proj_p(embedding) and the
(N, N, 3, C) tensor of the reference are never materialised.
"""
import functools

import threading

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..ops import e2pn_ops as K
from ..ops import transformer_ops as T
from ..ops.gemm import linear_bf16
from .e2pn import _Bf16Cache, _act


# the tabulated embedding (se3et_geo_embed_lookup) replaces the in-kernel sinusoid GEMM unless switched off (tests)
_EMBED_TABLE = {'on': True}


# A/B switch (tests): the positional query projection folded into one Linear of the layer input
_QP_FOLDED = {'on': True}


class CloudContext:
    """Flat layout bookkeeping for a batch of pairs. sizes = [n_ref_0.., n_src_0..] (python ints)."""

    def __init__(self, ref_sizes, src_sizes, anchors, heads, device):
        self.sizes = [int(v) for v in ref_sizes] + [int(v) for v in src_sizes]
        self.A, self.H = anchors, heads
        nref = len(ref_sizes)
        paired = len(src_sizes) == nref
        self.P = nref if paired else 0
        sz = np.asarray(self.sizes, dtype=np.int64)
        cu = np.concatenate([[0], np.cumsum(sz)])
        eoff = np.concatenate([[0], np.cumsum(sz * sz)])
        self.T, self.R = int(cu[-1]), int(eoff[-1])
        self.Tr = int(cu[nref])
        self.max_n = int(sz.max()) if len(sz) else 0
        ah = anchors * heads
        b = len(sz)
        P = self.P
        self_pr = np.stack([cu[:-1], sz, cu[:-1], sz, eoff[:-1] * ah], 1)
        # cross attention: q rows / kv rows are relative to the side's slice of the flat arrays
        ref_pr = np.stack([cu[:P], sz[:P], cu[P:2 * P] - cu[P], sz[P:2 * P], np.zeros(P, np.int64)], 1)
        src_pr = np.stack([cu[P:2 * P] - cu[P], sz[P:2 * P], cu[:P], sz[:P], np.zeros(P, np.int64)], 1)
        host = np.concatenate([cu, eoff[:-1], self_pr.reshape(-1), ref_pr.reshape(-1), src_pr.reshape(-1)]).astype(np.int64)
        dev = torch.from_numpy(host).to(device, non_blocking=True)
        o = 0
        self.cu = dev[o:o + b + 1]; o += b + 1
        self.eoff = dev[o:o + b]; o += b
        self.self_problems = dev[o:o + 5 * b].view(b, 5); o += 5 * b
        self.ref_problems = dev[o:o + 5 * P].view(P, 5); o += 5 * P
        self.src_problems = dev[o:o + 5 * P].view(P, 5)
        self.max_ref = int(sz[:P].max()) if P else 0
        self.max_src = int(sz[P:2 * P].max()) if P else 0
        # point offsets of the clouds inside their own side (reference side / source side)
        self.cu_ref = self.cu[:P + 1]
        self.cu_src = (self.cu[P:2 * P + 1] - self.cu[P]) if P else self.cu[:1]
        self._args = (list(ref_sizes), list(src_sizes), heads, device)
        self._inv = None
        # grouped-GEMM table of the positional score term: one group per (cloud, query point)
        sizes_t = torch.from_numpy(sz).to(device, non_blocking=True)
        cloud = torch.repeat_interleave(torch.arange(b, device=device), sizes_t, output_size=self.T)
        nloc = torch.arange(self.T, device=device) - self.cu[cloud]
        nb = sizes_t[cloud]
        a_row0 = self.eoff[cloud] + nloc * nb
        self.rpe_groups = torch.stack([a_row0, torch.arange(self.T, device=device) * ah, nb, a_row0 * ah, nb,
                                       torch.zeros_like(nb)], 1).contiguous()


def _ctx_inv(ctx):
    """The same clouds with one anchor (invariant 'self' / 'cross' blocks)."""
    if ctx.A == 1:
        return ctx
    if ctx._inv is None:
        r, sz, h, dev = ctx._args
        ctx._inv = CloudContext(r, sz, 1, h, dev)
    return ctx._inv


class SinusoidalPositionalEmbedding(nn.Module):
    """positional_embedding.py:8-34 (kept for the div_term buffer; the kernel regenerates the same table)."""

    def __init__(self, d_model):
        super().__init__()
        if d_model % 2 != 0:
            raise ValueError(f'Sinusoidal positional encoding with odd d_model: {d_model}')
        self.d_model = d_model
        div_indices = torch.arange(0, d_model, 2).float()
        self.register_buffer('div_term', torch.exp(div_indices * (-np.log(10000.0) / d_model)))


class GeometricStructureEmbedding(nn.Module):
    """geotransformer.py:19-121.  n_level_equiv = 2 (SE3ET-E) adds the l <= 1 spherical-harmonics embedding, which
    the CUDA path never materialises: its score contribution is added by se3et_sh_bias_add inside the equivariant
    self attention.  e3nn is not part of this image; the convention is the one stated in DESIGN.md (l = 1 == (x, y, z),
    D^1(R) = R, 'integral' normalisation)."""

    def __init__(self, hidden_dim, sigma_d, sigma_a, angle_k, reduction_a='max', kanchor=1, n_level_equiv=0):
        super().__init__()
        if reduction_a != 'max' or n_level_equiv not in (0, 2):
            raise NotImplementedError("CUDA path: reduction_a='max', n_level_equiv in (0, 2)")
        self.sigma_d, self.sigma_a, self.angle_k = sigma_d, sigma_a, angle_k
        self.factor_a = 180.0 / (self.sigma_a * np.pi)
        self.embedding = SinusoidalPositionalEmbedding(hidden_dim)
        self.proj_d = nn.Linear(hidden_dim, hidden_dim)
        self.proj_a = nn.Linear(hidden_dim, hidden_dim)
        self.n_level_equiv, self.kanchor, self.reduction_a = n_level_equiv, kanchor, reduction_a
        if n_level_equiv > 0 and kanchor > 1:
            if kanchor != 6:
                raise NotImplementedError("CUDA path: kanchor = 6")
            from . import octahedral
            anchors = torch.tensor(octahedral.tables()["anchors"], dtype=torch.float32)
            # same names / shapes as the reference's ParameterList of Wigner-D matrices of anchors^T (l = 0, 1)
            self.anchors_wignerD = nn.ParameterList([
                nn.Parameter(torch.ones(kanchor, 1, 1), requires_grad=False),
                nn.Parameter(anchors.transpose(1, 2).contiguous(), requires_grad=False)])
        self._wd, self._wa = _Bf16Cache(), _Bf16Cache()
        self._table_lock, self._table_state, self._table_retired = threading.Lock(), None, []

    def anchors_matrix(self):
        """(A, 3, 3) fp32 anchors (the transpose of the stored l = 1 Wigner-D matrices)."""
        return self.anchors_wignerD[1].detach().transpose(1, 2).contiguous()

    TABLE_STEP = 1.0 / 512.0  # spacing of the tabulated projections (highest embedding frequency: 1 rad / unit)

    def _tables(self, u_max, device):
        """bf16 tables of W_d emb(u) + b_d (u in [0, u_max]) and W_a emb(a) + b_a (a in [0, 180 / sigma_a]): exact fp32
        sinusoids and an fp32 matmul, once per weight version and size bucket (parameter preprocessing like the bf16
        weight copies, not part of the per-pair path)."""
        nd = 1 << int(np.ceil(np.log2(max(u_max, 1.0) / self.TABLE_STEP + 2)))
        wkey = (self.proj_d.weight._version, self.proj_a.weight._version, self.proj_d.bias._version,
                self.proj_a.bias._version, self.proj_d.weight.data_ptr(), str(device))
        # shared by the host threads / streams of concurrent launch sequences: built under a lock, the LARGEST table is
        # kept (a smaller request reuses it: same step, more rows), other streams wait on the build's event once
        with self._table_lock:
            cur = self._table_state
            if cur is None or cur['wkey'] != wkey or cur['nd'] < nd:
                with torch.no_grad():
                    div = self.embedding.div_term.to(device).float()

                    def table(n, lin):
                        u = torch.arange(n, device=device, dtype=torch.float32) * self.TABLE_STEP
                        om = u[:, None] * div[None, :]
                        e = torch.stack([torch.sin(om), torch.cos(om)], dim=2).reshape(n, -1)
                        return torch.addmm(lin.bias.float(), e, lin.weight.float().t()).to(torch.bfloat16).contiguous()
                    na = int(180.0 / self.sigma_a / self.TABLE_STEP) + 2
                    st = torch.cuda.current_stream(device)
                    ev = torch.cuda.Event()
                    cur = {'wkey': wkey, 'nd': nd, 'd': table(nd, self.proj_d), 'a': table(na, self.proj_a), 'event': ev,
                           'synced': {st.cuda_stream}}
                    ev.record(st)
                if self._table_state is not None:
                    # another stream may still be reading the smaller tables: keep them alive (sizes only double)
                    self._table_retired.append(self._table_state)
                self._table_state = cur
            st = torch.cuda.current_stream(device)
            if st.cuda_stream not in cur['synced']:
                st.wait_event(cur['event'])
                cur['synced'].add(st.cuda_stream)
            return cur['d'], cur['a']

    def embed(self, points_flat, ctx):
        """-> bf16 (sum n_b^2, C): row eoff[b] + n*n_b + m is the embedding of the pair (n, m) of cloud b."""
        idx4 = T.geo_embed_indices(points_flat, ctx.cu, ctx.max_n, ctx.eoff, ctx.R, self.sigma_d, self.sigma_a,
                                   self.angle_k)
        c = self.proj_d.weight.shape[0]
        if _EMBED_TABLE['on'] and c in (64, 128, 256) and idx4.shape[0] > 0:
            u_max = float(idx4[:, 0].max())  # one scalar read-back per launch sequence
            td, ta = self._tables(u_max, idx4.device)
            return T.geo_embed_lookup(idx4, td, ta, self.TABLE_STEP)
        bias_sum = (self.proj_d.bias + self.proj_a.bias).detach().float().contiguous()
        return T.geo_embed_project(idx4, self._wd.get(self.proj_d.weight), self._wa.get(self.proj_a.weight), bias_sum)

    def forward(self, points):
        """points (B=1, N, 3) -> (1, N, N, C) as the reference (fp32)."""
        assert points.shape[0] == 1
        n = points.shape[1]
        ctx = CloudContext([n], [], 1, 1, points.device)
        e = self.embed(points[0].contiguous().float(), ctx)
        return e.view(1, n, n, -1).float()


class AttentionOutput(nn.Module):
    """output_layer.py:7-22: LayerNorm(x + squeeze(ReLU(expand(x))))."""

    def __init__(self, d_model, dropout=None, activation_fn='ReLU'):
        super().__init__()
        if activation_fn != 'ReLU' or dropout:
            raise NotImplementedError("CUDA path: ReLU, no dropout (inference)")
        self.expand = nn.Linear(d_model, d_model * 2)
        self.squeeze = nn.Linear(d_model * 2, d_model)
        self.norm = nn.LayerNorm(d_model)
        self._we, self._ws = _Bf16Cache(), _Bf16Cache()

    def fused(self, x, out=None):
        """x bf16 (rows, C) -> bf16 (rows, C)."""
        _, h = linear_bf16(x, self._we.get(self.expand.weight), self.expand.bias, relu=True, out_f32=False, out_bf16=True)
        ws = self._ws.get(self.squeeze.weight)
        if T.linear_add_layernorm_supported(h, ws):   # squeeze + residual + LayerNorm in one kernel
            return T.linear_add_layernorm(h, ws, self.squeeze.bias, x, 1, self.norm.weight, self.norm.bias, self.norm.eps)
        z, _ = linear_bf16(h, ws, self.squeeze.bias)
        _, y = T.add_layernorm(z, x, 1, self.norm.weight, self.norm.bias, self.norm.eps)
        return y

    def forward(self, input_states):
        shape = input_states.shape
        y = self.fused(_act(input_states).reshape(-1, shape[-1]).contiguous())
        return y.view(shape).to(input_states.dtype)


class RPEMultiHeadAttention(nn.Module):
    """rpe_transformer.py:18-131 (equivariant branch)."""

    def __init__(self, d_model, num_heads, dropout=None, equivariant=False, d_equiv_embed=0):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError('`d_model` ({}) must be a multiple of `num_heads` ({}).'.format(d_model, num_heads))
        if dropout or d_equiv_embed not in (0, 4):
            raise NotImplementedError("CUDA path: no dropout, d_equiv_embed in (0, 4)")
        self.d_model, self.num_heads = d_model, num_heads
        self.d_model_per_head = d_model // num_heads
        self.equivariant, self.d_equiv_embed = equivariant, d_equiv_embed
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        self.proj_p = nn.Linear(d_model, d_model)
        if self.equivariant and d_equiv_embed > 0:
            self.proj_eq = nn.Linear(d_equiv_embed, d_model)
        self._wqkv, self._wpt, self._wu, self._bqp = _Bf16Cache(), _Bf16Cache(), _Bf16Cache(), _Bf16Cache()

    def _u_weight(self):
        """(16, C) bf16, block diagonal over heads: row 3 h + d = proj_eq.weight[head h channels, 1 + d]."""
        h, hc = self.num_heads, self.d_model_per_head

        def tr(w):  # w: (C, 4)
            out = torch.zeros((16, self.d_model), dtype=w.dtype, device=w.device)
            for i in range(h):
                out[3 * i:3 * i + 3, i * hc:(i + 1) * hc] = w[i * hc:(i + 1) * hc, 1:4].t()
            return out
        return self._wu.get(self.proj_eq.weight, tr)

    def fused(self, x, emb, ctx, points=None, anchors_mat=None):
        """x bf16 (T*A, C) equivariant states, emb bf16 (R, C) -> attention output bf16 (T*A, C)."""
        c, h, a = self.d_model, self.num_heads, ctx.A
        hc = self.d_model_per_head
        w = self._wqkv.get(self.proj_q.weight, lambda q: torch.cat([q, self.proj_k.weight.detach(),
                                                                    self.proj_v.weight.detach()], 0))
        b = torch.cat([self.proj_q.bias, self.proj_k.bias, self.proj_v.bias]).detach()
        _, qkv = linear_bf16(x, w, b, out_f32=False, out_bf16=True)  # (T*A, 3C)
        # qp[(n,a), h, :] = W_p[h]^T q[(n,a), h]  (the proj_p bias only shifts every key equally: softmax-invariant).
        # With q = W_q x + b_q this is ONE Linear of x: qp[h] = (W_p[h]^T W_q[h]) x + W_p[h]^T b_q[h] -- the four per-head
        # GEMMs on the bf16 q (and q's rounding in between) become a single launch with 4 C output columns
        rows = x.shape[0]
        if _QP_FOLDED['on']:
            def fold(p):
                wq, wp = self.proj_q.weight.detach().float(), p.float()
                return torch.cat([wp[i * hc:(i + 1) * hc].t() @ wq[i * hc:(i + 1) * hc] for i in range(h)], 0)  # (h C, C)
            wqp = self._wpt.get(self.proj_p.weight, fold, extra=self.proj_q.weight)

            def fold_bias(p):
                bq, wp = self.proj_q.bias.detach().float(), p.float()
                return torch.cat([wp[i * hc:(i + 1) * hc].t() @ bq[i * hc:(i + 1) * hc] for i in range(h)])
            bqp = self._bqp.get(self.proj_p.weight, fold_bias, extra=self.proj_q.bias, dtype=torch.float32)
            _, qp = linear_bf16(x, wqp, bqp, out_f32=False, out_bf16=True)
        else:
            wpt = self._wpt.get(self.proj_p.weight, lambda p: p.t())
            qp = torch.empty((rows, h * c), dtype=torch.bfloat16, device=x.device)
            for i in range(h):
                linear_bf16(qkv[:, i * hc:(i + 1) * hc], wpt[:, i * hc:(i + 1) * hc], out_f32=False,
                            out_bf16=qp[:, i * c:(i + 1) * c])
        ah = a * h
        s_p = torch.empty((ctx.R * ah,), dtype=torch.float32, device=x.device)
        T.gemm_grouped_t(emb, ctx.R, qp.view(rows * h, c), rows * h, ctx.rpe_groups, ctx.max_n, (ah + 15) // 16 * 16, ah,
                         c, s_p)
        if self.equivariant and self.d_equiv_embed > 0 and a > 1:
            if points is None or anchors_mat is None:
                raise RuntimeError("equivariant embedding requires the superpoint coordinates")
            if h * 3 > 16:
                raise NotImplementedError("CUDA path: at most 5 heads with the equivariant embedding")
            u, _ = linear_bf16(qkv[:, :c], self._u_weight())  # fp32 (T*A, 16)
            T.sh_bias_add(points, ctx.self_problems, ctx.max_n, u, anchors_mat, h, float(np.sqrt(3.0 / (4.0 * np.pi))),
                          s_p)
        hidden = torch.empty((rows, c), dtype=torch.bfloat16, device=x.device)
        T.flash_attention(qkv, a * 3 * c, 3 * c, qkv[:, c:], a * 3 * c, 3 * c, qkv[:, 2 * c:], a * 3 * c, 3 * c, s_p,
                          ctx.self_problems, ctx.max_n, a, h, hc, hidden)
        return hidden

    def forward(self, input_q, input_k, input_v, embed_qk, key_weights=None, key_masks=None, attention_factors=None,
                embed_eq=None):
        """Self-attention form only (input_q is input_k is input_v): (1, A, N, C), embed (1, N, N, C)."""
        if key_weights is not None or key_masks is not None or attention_factors is not None or embed_eq is not None:
            raise NotImplementedError("CUDA path: plain equivariant self-attention")
        assert input_q.shape[0] == 1 and input_q.dim() == 4
        _, a, n, c = input_q.shape
        ctx = CloudContext([n], [], a, self.num_heads, input_q.device)
        x = _act(input_q[0]).transpose(0, 1).reshape(n * a, c).contiguous()
        hid = self.fused(x, _act(embed_qk).reshape(n * n, c).contiguous(), ctx)
        return hid.view(n, a, c).transpose(0, 1).unsqueeze(0).to(input_q.dtype), None


class RPEAttentionLayer(nn.Module):
    def __init__(self, d_model, num_heads, dropout=None, equivariant=False, d_equiv_embed=0):
        super().__init__()
        self.attention = RPEMultiHeadAttention(d_model, num_heads, dropout=dropout, equivariant=equivariant,
                                               d_equiv_embed=d_equiv_embed)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)
        self._wl = _Bf16Cache()

    def fused(self, x, emb, ctx, points=None, anchors_mat=None):
        hid = self.attention.fused(x, emb, ctx, points, anchors_mat)
        wl = self._wl.get(self.linear.weight)
        if T.linear_add_layernorm_supported(hid, wl):
            return T.linear_add_layernorm(hid, wl, self.linear.bias, x, 1, self.norm.weight, self.norm.bias, self.norm.eps)
        y, _ = linear_bf16(hid, wl, self.linear.bias)
        _, out = T.add_layernorm(y, x, 1, self.norm.weight, self.norm.bias, self.norm.eps)
        return out


class RPETransformerLayer(nn.Module):
    """rpe_transformer.py:168-194."""

    def __init__(self, d_model, num_heads, dropout=None, activation_fn='ReLU', equivariant=False, d_equiv_embed=0):
        super().__init__()
        self.attention = RPEAttentionLayer(d_model, num_heads, dropout=dropout, equivariant=equivariant,
                                           d_equiv_embed=d_equiv_embed)
        self.output = AttentionOutput(d_model, dropout=dropout, activation_fn=activation_fn)

    def fused(self, x, emb, ctx, points=None, anchors_mat=None):
        """equivariant: x (T*A, C) with ctx.A anchors; invariant ('self' blocks): x (T, C) with a one-anchor ctx."""
        return self.output.fused(self.attention.fused(x, emb, ctx, points, anchors_mat))

    def forward(self, input_states, memory_states, position_states, memory_weights=None, memory_masks=None,
                attention_factors=None, equiv_states=None):
        if memory_states is not input_states or memory_masks is not None:
            raise NotImplementedError("CUDA path: self-attention without masks")
        _, a, n, c = input_states.shape
        ctx = CloudContext([n], [], a, self.attention.attention.num_heads, input_states.device)
        x = _act(input_states[0]).transpose(0, 1).reshape(n * a, c).contiguous()
        out = self.fused(x, _act(position_states).reshape(n * n, c).contiguous(), ctx)
        return out.view(n, a, c).transpose(0, 1).unsqueeze(0).to(input_states.dtype), None


class MultiHeadAttention(nn.Module):
    """vanilla_transformer.py:23-85, invariant q/k with an equivariant (4-D) value."""

    def __init__(self, d_model, num_heads, dropout=None):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError('`d_model` ({}) must be a multiple of `num_heads` ({}).'.format(d_model, num_heads))
        if dropout:
            raise NotImplementedError("CUDA path: no dropout (inference)")
        self.d_model, self.num_heads = d_model, num_heads
        self.d_model_per_head = d_model // num_heads
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        self._wq, self._wk, self._wv = _Bf16Cache(), _Bf16Cache(), _Bf16Cache()

    def fused(self, q_inv, k_inv, v_eq, problems, max_q, anchors):
        """q_inv bf16 (Nq, C), k_inv bf16 (Nk, C), v_eq bf16 (Nk*A, C) -> bf16 (Nq*A, C)."""
        c, h, hc = self.d_model, self.num_heads, self.d_model_per_head
        _, q = linear_bf16(q_inv, self._wq.get(self.proj_q.weight), self.proj_q.bias, out_f32=False, out_bf16=True)
        _, k = linear_bf16(k_inv, self._wk.get(self.proj_k.weight), self.proj_k.bias, out_f32=False, out_bf16=True)
        _, v = linear_bf16(v_eq, self._wv.get(self.proj_v.weight), self.proj_v.bias, out_f32=False, out_bf16=True)
        hidden = torch.empty((q_inv.shape[0] * anchors, c), dtype=torch.bfloat16, device=q_inv.device)
        T.flash_attention(q, c, 0, k, c, 0, v, anchors * c, c, None, problems, max_q, anchors, h, hc, hidden)
        return hidden


class MultiHeadAttentionEQ(nn.Module):
    """vanilla_transformer.py:87-870 in the modes SE3ET-E uses: 'a_soft' (every query anchor attends to every key
    anchor, weighted by a global anchor-pair statistic) and 'r_soft' (the 24 rotations of the key anchors, weighted by
    a global rotation statistic).  Both reduce to  hidden[a] = sum_e W[a, e] softmax(q_a k_e^T / sqrt(c)) v_e  with a
    per-pair 6 x 6 matrix W (csrc/eq_attention.cu)."""

    def __init__(self, d_model, num_heads, dropout=None, attn_mode=None, alternative_impl=False, kanchor=4,
                 attn_r_positive='sq', attn_r_positive_rot_supervise='sigmoid'):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError('`d_model` ({}) must be a multiple of `num_heads` ({}).'.format(d_model, num_heads))
        if dropout or attn_mode not in ('a_soft', 'r_soft') or kanchor != 6 or attn_r_positive not in T.POSITIVE:
            raise NotImplementedError("CUDA path: attn_mode a_soft / r_soft, kanchor 6, attn_r_positive in %s"
                                      % sorted(T.POSITIVE))
        self.d_model, self.num_heads = d_model, num_heads
        self.d_model_per_head = d_model // num_heads
        self.attn_mode, self.kanchor = attn_mode, kanchor
        self.attn_r_positive = attn_r_positive
        self.attn_r_positive_rot_supervise = attn_r_positive_rot_supervise
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        perms, rots = _octahedral_rotations()
        # reference names (vanilla_transformer.py:177-184); a reference checkpoint overwrites them with its own order
        self.anchors = nn.Parameter(torch.tensor(rots, dtype=torch.float32), requires_grad=False)
        self.trace_idx_ori = nn.Parameter(torch.tensor(perms, dtype=torch.int64), requires_grad=False)
        self.trace_idx_rot = nn.Parameter(torch.tensor(np.argsort(perms, axis=1), dtype=torch.int64), requires_grad=False)
        self.nr, self.na = perms.shape
        self._wq, self._wkv = _Bf16Cache(), _Bf16Cache()

    def fused(self, x_q, x_k, problems, max_q, cloud_off, anchors):
        """x_q (Nq*A, C), x_k (Nk*A, C) bf16 equivariant states -> (hidden (Nq*A, C) bf16, W (P, A, A), attn_r)."""
        c, h, hc, a = self.d_model, self.num_heads, self.d_model_per_head, anchors
        _, q = linear_bf16(x_q, self._wq.get(self.proj_q.weight), self.proj_q.bias, out_f32=False, out_bf16=True)
        wkv = self._wkv.get(self.proj_k.weight, lambda k: torch.cat([k, self.proj_v.weight.detach()], 0))
        bkv = torch.cat([self.proj_k.bias, self.proj_v.bias]).detach()
        _, kv = linear_bf16(x_k, wkv, bkv, out_f32=False, out_bf16=True)  # (Nk*A, 2C): k | v
        g = T.anchor_pair_stats(q, a * c, c, kv, a * 2 * c, 2 * c, problems, max_q, a, c, h, self.attn_r_positive)
        perms = self.trace_idx_ori.to(torch.int32).contiguous()
        w, attn_r = T.anchor_mix_weights(g, problems, perms, self.attn_mode == 'r_soft')
        nq = x_q.shape[0] // a
        per_e = torch.empty((a, nq * a, c), dtype=torch.bfloat16, device=x_q.device)
        for e in range(a):  # key / value anchor e for every query anchor: rows e, e + A, ... of kv
            T.flash_attention(q, a * c, c, kv[e:, :c], a * 2 * c, 0, kv[e:, c:], a * 2 * c, 0, None, problems, max_q,
                              a, h, hc, per_e[e])
        hidden = T.anchor_mix(per_e, nq * a * c, a * c, c, w, cloud_off, a, c, nq)
        return hidden, w, attn_r


@functools.lru_cache(maxsize=None)
def _octahedral_rotations():
    """(perms (24, 6) int64, rotations (24, 3, 3)): the proper rotations of the octahedron and the permutation each
    induces on the vertices [+z,+x,+y,-x,-y,-z]: perms[r][a] = index of R_r v_a (fr.get_relativeV_index,
    utils_epn/rotation.py:581-601).  Row order: lexicographic in the permutation."""
    import itertools
    from . import octahedral
    vs = octahedral.VERTICES
    rows = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            R = np.zeros((3, 3))
            for i in range(3):
                R[i, perm[i]] = signs[i]
            if np.linalg.det(R) > 0:
                rows.append(([int(np.argmin(((vs - R @ v) ** 2).sum(1))) for v in vs], R))
    rows.sort(key=lambda t: t[0])
    return np.array([r[0] for r in rows], dtype=np.int64), np.stack([r[1] for r in rows])


class RotCompressOutput(nn.Module):
    """output_layer.py:24-47: LayerNorm(max_a x + squeeze(ReLU(expand(concat_a x))))."""

    def __init__(self, d_model, dropout=None, activation_fn='ReLU', na=12, dual_align=False):
        super().__init__()
        if dropout or activation_fn != 'ReLU' or dual_align:
            raise NotImplementedError("CUDA path: ReLU, no dropout, align_mode '0'")
        self.na = na
        self.expand = nn.Linear(d_model * na, d_model * 2)
        self.squeeze = nn.Linear(d_model * 2, d_model)
        self.norm = nn.LayerNorm(d_model)
        self._we, self._ws = _Bf16Cache(), _Bf16Cache()

    def fused(self, x_eq):
        """x_eq bf16 (T*A, C) -> bf16 (T, C)."""
        c = x_eq.shape[1]
        t = x_eq.shape[0] // self.na
        x_max = K.anchor_max(x_eq.view(t, self.na, c))
        _, hdn = linear_bf16(x_eq.view(t, self.na * c), self._we.get(self.expand.weight), self.expand.bias, relu=True,
                             out_f32=False, out_bf16=True)
        z, _ = linear_bf16(hdn, self._ws.get(self.squeeze.weight), self.squeeze.bias)
        _, y = T.add_layernorm(z, x_max, 1, self.norm.weight, self.norm.bias, self.norm.eps)
        return y


class AttentionLayer(nn.Module):
    def __init__(self, d_model, num_heads, dropout=None, equivariant=False, attn_mode=None, alternative_impl=False,
                 kanchor=4, attn_r_positive='sq', attn_r_positive_rot_supervise='sigmoid'):
        super().__init__()
        self.equivariant = equivariant
        if equivariant:
            self.attention = MultiHeadAttentionEQ(d_model, num_heads, dropout=dropout, attn_mode=attn_mode,
                                                  alternative_impl=alternative_impl, kanchor=kanchor,
                                                  attn_r_positive=attn_r_positive,
                                                  attn_r_positive_rot_supervise=attn_r_positive_rot_supervise)
        else:
            self.attention = MultiHeadAttention(d_model, num_heads, dropout=dropout)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)
        self._wl = _Bf16Cache()

    def fused(self, q_inv, k_inv, v_eq, problems, max_q, anchors):
        hid = self.attention.fused(q_inv, k_inv, v_eq, problems, max_q, anchors)
        wl = self._wl.get(self.linear.weight)
        # (N, C) residual broadcast onto (A, N, C) (vanilla_transformer.py:911)
        if T.linear_add_layernorm_supported(hid, wl):
            return T.linear_add_layernorm(hid, wl, self.linear.bias, q_inv, anchors, self.norm.weight, self.norm.bias,
                                          self.norm.eps)
        y, _ = linear_bf16(hid, wl, self.linear.bias)
        _, out = T.add_layernorm(y, q_inv, anchors, self.norm.weight, self.norm.bias, self.norm.eps)
        return out

    def fused_eq(self, x_q, x_k, problems, max_q, cloud_off, anchors):
        """equivariant layer (vanilla_transformer.py:886-915): x_q (Nq*A, C), x_k (Nk*A, C) -> ((Nq*A, C), W, attn_r)."""
        hid, w, attn_r = self.attention.fused(x_q, x_k, problems, max_q, cloud_off, anchors)
        wl = self._wl.get(self.linear.weight)
        if T.linear_add_layernorm_supported(hid, wl):
            return T.linear_add_layernorm(hid, wl, self.linear.bias, x_q, 1, self.norm.weight, self.norm.bias,
                                          self.norm.eps), w, attn_r
        y, _ = linear_bf16(hid, wl, self.linear.bias)
        _, out = T.add_layernorm(y, x_q, 1, self.norm.weight, self.norm.bias, self.norm.eps)
        return out, w, attn_r


class TransformerLayer(nn.Module):
    """vanilla_transformer.py:917-946."""

    def __init__(self, d_model, num_heads, dropout=None, activation_fn='ReLU', equivariant=False, attn_mode=None,
                 alternative_impl=False, kanchor=4, attn_r_positive='sq', attn_r_positive_rot_supervise='sigmoid'):
        super().__init__()
        self.equivariant = equivariant
        self.attention = AttentionLayer(d_model, num_heads, dropout=dropout, equivariant=equivariant,
                                        attn_mode=attn_mode, alternative_impl=alternative_impl, kanchor=kanchor,
                                        attn_r_positive=attn_r_positive,
                                        attn_r_positive_rot_supervise=attn_r_positive_rot_supervise)
        self.output = AttentionOutput(d_model, dropout=dropout, activation_fn=activation_fn)

    def fused(self, q_inv, k_inv, v_eq, problems, max_q, anchors):
        return self.output.fused(self.attention.fused(q_inv, k_inv, v_eq, problems, max_q, anchors))

    def fused_eq(self, x_q, x_k, problems, max_q, cloud_off, anchors):
        out, w, attn_r = self.attention.fused_eq(x_q, x_k, problems, max_q, cloud_off, anchors)
        return self.output.fused(out), w, attn_r

    def forward(self, input_states, memory_states, value_states=None, memory_weights=None, memory_masks=None,
                attention_factors=None, attention_masks=None, gt_indices=None, gt_overlap=None):
        """input (1, N, C), memory (1, M, C), value (1, A, M, C) -> (1, A, N, C)."""
        if value_states is None or value_states.dim() != 4 or memory_masks is not None:
            raise NotImplementedError("CUDA path: invariant q/k with an equivariant value, no masks")
        n, m = input_states.shape[1], memory_states.shape[1]
        a, c = value_states.shape[1], value_states.shape[-1]
        dev = input_states.device
        problems = torch.tensor([[0, n, 0, m, 0]], dtype=torch.int64, device=dev)
        v = _act(value_states[0]).transpose(0, 1).reshape(m * a, c).contiguous()
        out = self.fused(_act(input_states[0]).contiguous(), _act(memory_states[0]).contiguous(), v, problems, n, a)
        return out.view(n, a, c).transpose(0, 1).unsqueeze(0).to(input_states.dtype), None


def _check_block_type(block):
    if 'self' not in block and 'cross' not in block:
        raise ValueError('Unsupported block type "{}".'.format(block))


def _check_block_eq(block):
    return '_' in block  # 'self_eq', 'cross_a_soft', 'cross_r_soft' keep the anchor axis (conditional_transformer.py:86-91)


class RPEConditionalTransformer(nn.Module):
    """conditional_transformer.py:98-390 for the block lists of SE3ET-I / I2 ('self_eq', 'cross') and of SE3ET-E / E2
    ('self_eq', 'cross_a_soft', 'cross_r_soft', then invariant 'self' / 'cross'; align_mode '0')."""

    def __init__(self, blocks, d_model, num_heads, dropout=None, activation_fn='ReLU', return_attention_scores=False,
                 return_attention_weights=False, anchor_matching=False, parallel=False, na=4, attn_r_positive='sq',
                 attn_r_positive_rot_supervise='sigmoid', align_mode='0', alternative_impl=False, d_equiv_embed=0):
        super().__init__()
        if return_attention_scores or return_attention_weights or anchor_matching or parallel or align_mode != '0':
            raise NotImplementedError("CUDA path: inference configuration (align_mode '0', no attention outputs)")
        self.blocks, self.na, self.align_mode = list(blocks), na, align_mode
        self.d_equiv_embed = d_equiv_embed
        self.plain = all(b in ('self_eq', 'cross') for b in blocks)  # SE3ET-I / I2 schedule
        layers = []
        for i, block in enumerate(blocks):
            _check_block_type(block)
            if self.plain and block == 'self_eq' and (i + 1 >= len(blocks) or blocks[i + 1] != 'cross'):
                raise NotImplementedError("CUDA path: every 'self_eq' block is followed by a 'cross' block")
            if block in ('self_eq', 'self'):
                eq = block == 'self_eq'
                layers.append(RPETransformerLayer(d_model, num_heads, dropout=dropout, activation_fn=activation_fn,
                                                  equivariant=eq, d_equiv_embed=d_equiv_embed if eq else 0))
            elif block == 'cross':
                layers.append(TransformerLayer(d_model, num_heads, dropout=dropout, activation_fn=activation_fn,
                                               equivariant=False, kanchor=na))
            elif block in ('cross_a_soft', 'cross_r_soft'):
                layers.append(TransformerLayer(d_model, num_heads, dropout=dropout, activation_fn=activation_fn,
                                               equivariant=True, attn_mode=block[len('cross_'):], kanchor=na,
                                               alternative_impl=alternative_impl, attn_r_positive=attn_r_positive,
                                               attn_r_positive_rot_supervise=attn_r_positive_rot_supervise))
            else:
                raise NotImplementedError("CUDA path: block '%s' (a_best / r_best / cross_eq are not built)" % block)
        self.layers = nn.ModuleList(layers)
        if 'cross_r_soft' in blocks:
            self.rotcompress = RotCompressOutput(d_model, dropout=dropout, activation_fn=activation_fn, na=na)
        if not self.plain:
            # after the last equivariant block the features must have been pooled (r_soft + RotCompressOutput)
            last_eq = max(i for i, b in enumerate(blocks) if _check_block_eq(b))
            if blocks[last_eq] != 'cross_r_soft' or any(_check_block_eq(b) for b in blocks[last_eq + 1:]) or \
                    last_eq + 1 >= len(blocks):
                raise NotImplementedError("CUDA path: SE3ET-E schedule (equivariant blocks, 'cross_r_soft', then "
                                          "invariant 'self' / 'cross' blocks)")

    def run(self, x_eq, emb, ctx, points=None, anchors_mat=None):
        """x_eq bf16 (T*A, C), clouds ordered [refs | srcs] -> invariant bf16 (T, C)."""
        if not self.plain:
            return self._run_eq(x_eq, emb, ctx, points, anchors_mat)
        a = ctx.A
        c = x_eq.shape[1]
        tr = ctx.Tr
        x_inv = None
        for layer, block in zip(self.layers, self.blocks):
            if block == 'self_eq':
                x_eq = layer.fused(x_eq, emb, ctx, points, anchors_mat)
                x_inv = K.anchor_max(x_eq.view(-1, a, c))
            else:
                if x_inv is None:
                    x_inv = K.anchor_max(x_eq.view(-1, a, c))
                new_eq = torch.empty_like(x_eq)
                new_inv = torch.empty_like(x_inv)
                # the reference side is updated first; the source side then attends to the UPDATED reference
                # (conditional_transformer.py:297-302)
                r = layer.fused(x_inv[:tr], x_inv[tr:], x_eq[tr * a:], ctx.ref_problems, ctx.max_ref, a)
                new_eq[:tr * a] = r
                K.anchor_max(r.view(-1, a, c), out=new_inv[:tr])
                s = layer.fused(x_inv[tr:], new_inv[:tr], new_eq[:tr * a], ctx.src_problems, ctx.max_src, a)
                new_eq[tr * a:] = s
                K.anchor_max(s.view(-1, a, c), out=new_inv[tr:])
                x_eq, x_inv = new_eq, new_inv
        return x_inv

    def _run_eq(self, x, emb, ctx, points, anchors_mat):
        """SE3ET-E schedule (conditional_transformer.py:262-360 with feats*_eq = None throughout)."""
        a, tr = ctx.A, ctx.Tr
        c = x.shape[1]
        ctx1 = _ctx_inv(ctx)
        for i, (layer, block) in enumerate(zip(self.layers, self.blocks)):
            if block == 'self_eq':
                x = layer.fused(x, emb, ctx, points, anchors_mat)
            elif block in ('cross_a_soft', 'cross_r_soft'):
                # reference side first, the source side then attends to the UPDATED reference (:346-347)
                r, w0, _ = layer.fused_eq(x[:tr * a], x[tr * a:], ctx.ref_problems, ctx.max_ref, ctx.cu_ref, a)
                s, _, _ = layer.fused_eq(x[tr * a:], r, ctx.src_problems, ctx.max_src, ctx.cu_src, a)
                if block == 'cross_r_soft' and not _check_block_eq(self.blocks[i + 1]):
                    # eq2inv_soft, align_mode '0' (:209-249): source anchors mixed by the reference side's rotation
                    # weights, then RotCompressOutput on both sides
                    s = T.anchor_mix(s, c, a * c, 0, w0, ctx.cu_src, a, c, s.shape[0] // a)
                    x = torch.cat([self.rotcompress.fused(r), self.rotcompress.fused(s)], 0)
                else:
                    x = torch.cat([r, s], 0)
            elif block == 'self':
                x = layer.fused(x, emb, ctx1)
            else:  # invariant 'cross'
                r = layer.fused(x[:tr], x[tr:], x[tr:], ctx1.ref_problems, ctx1.max_ref, 1)
                s = layer.fused(x[tr:], r, r, ctx1.src_problems, ctx1.max_src, 1)
                x = torch.cat([r, s], 0)
        return x


class GeometricTransformer(nn.Module):
    """geotransformer.py:124-317.  forward() keeps the reference signature (one pair); forward_clouds() runs a
    batch of pairs in one launch sequence."""

    def __init__(self, input_dim, output_dim, hidden_dim, num_heads, blocks, sigma_d, sigma_a, angle_k, dropout=None,
                 activation_fn='ReLU', supervise_rotation=False, anchor_matching=False, reduction_a='max', na=None,
                 attn_r_positive='sq', attn_r_positive_rot_supervise='sigmoid', align_mode='0', alternative_impl=False,
                 n_level_equiv=0):
        super().__init__()
        if na is None or supervise_rotation or anchor_matching:
            raise NotImplementedError("CUDA path: equivariant features (na = kanchor), no rotation supervision")
        self.n_level_equiv, self.na = n_level_equiv, na
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.embedding = GeometricStructureEmbedding(hidden_dim, sigma_d, sigma_a, angle_k, reduction_a=reduction_a,
                                                     kanchor=na, n_level_equiv=n_level_equiv)
        self.in_proj = nn.Linear(input_dim, hidden_dim)
        self.d_equiv_embed = int((np.arange(n_level_equiv) * 2 + 1).sum())
        self.transformer = RPEConditionalTransformer(blocks, hidden_dim, num_heads, dropout=dropout,
                                                     activation_fn=activation_fn, na=na,
                                                     attn_r_positive=attn_r_positive,
                                                     attn_r_positive_rot_supervise=attn_r_positive_rot_supervise,
                                                     align_mode=align_mode, alternative_impl=alternative_impl,
                                                     d_equiv_embed=self.d_equiv_embed)
        self.out_proj = nn.Linear(hidden_dim, output_dim)
        self._wi, self._wo = _Bf16Cache(), _Bf16Cache()

    def forward_clouds(self, points, feats, ref_sizes, src_sizes):
        """points fp32 (T, 3), feats (T, A, Cin), clouds ordered [ref_0.., src_0..] with the given sizes.
        -> fp32 (T, output_dim) in the same order."""
        _lib.require_cuda(points, feats)
        ctx = CloudContext(ref_sizes, src_sizes, self.na, self.num_heads, points.device)
        assert ctx.T == points.shape[0] == feats.shape[0]
        emb = self.embedding.embed(points.contiguous().float(), ctx)
        x = _act(feats).reshape(ctx.T * self.na, -1).contiguous()
        _, x = linear_bf16(x, self._wi.get(self.in_proj.weight), self.in_proj.bias, out_f32=False, out_bf16=True)
        pts = points.contiguous().float()
        anchors_mat = self.embedding.anchors_matrix() if self.n_level_equiv > 0 else None
        x_inv = self.transformer.run(x, emb, ctx, pts, anchors_mat)
        out, _ = linear_bf16(x_inv, self._wo.get(self.out_proj.weight), self.out_proj.bias)
        return out

    def forward(self, ref_points, src_points, ref_feats, src_feats, ref_masks=None, src_masks=None, gt_indices=None,
                gt_overlap=None, ref_normal=None, src_normal=None):
        """(1, N, 3), (1, M, 3), (1, N, A, C), (1, M, A, C) -> (ref (1, N, Cout), src (1, M, Cout), None x 4)."""
        if ref_masks is not None or src_masks is not None or ref_normal is not None or src_normal is not None:
            raise NotImplementedError("CUDA path: no masks / normals (the reference's 3DMatch and KITTI configs)")
        assert ref_points.shape[0] == 1 and src_points.shape[0] == 1
        n, m = ref_points.shape[1], src_points.shape[1]
        points = torch.cat([ref_points[0], src_points[0]], 0)
        feats = torch.cat([_act(ref_feats[0]), _act(src_feats[0])], 0)
        out = self.forward_clouds(points, feats, [n], [m])
        return out[:n].unsqueeze(0), out[n:].unsqueeze(0), None, None, None, None


class SuperPointMatching(nn.Module):
    """superpoint_matching.py:7-55; ties in the top-k are ordered by flat index (the reference leaves them to
    torch.topk)."""

    def __init__(self, num_correspondences, dual_normalization=True):
        super().__init__()
        self.num_correspondences = num_correspondences
        self.dual_normalization = dual_normalization

    def forward_pairs(self, ref_feats, src_feats, ref_sizes, src_sizes, ref_masks=None, src_masks=None):
        """Batched: ref_feats (sum n_ref, C), src_feats (sum n_src, C) fp32 unit rows.
        -> ref_idx (P, k), src_idx (P, k) pair-local int64 (-1 padded), scores (P, k), counts (P,) int32."""
        rs, ss = np.asarray(ref_sizes, np.int64), np.asarray(src_sizes, np.int64)
        rcu = np.concatenate([[0], np.cumsum(rs)])
        scu = np.concatenate([[0], np.cumsum(ss)])
        eo = np.concatenate([[0], np.cumsum(rs * ss)])
        problems = torch.from_numpy(np.stack([rcu[:-1], rs, scu[:-1], ss, eo[:-1]], 1).astype(np.int64)).to(
            ref_feats.device, non_blocking=True)
        ri, si, sc, cnt, _ = T.superpoint_matching(ref_feats.float().contiguous(), src_feats.float().contiguous(),
                                                   ref_masks, src_masks, problems, int(rs.max()), int(ss.max()),
                                                   int(eo[-1]), self.num_correspondences, self.dual_normalization)
        return ri, si, sc, cnt

    def forward(self, ref_feats, src_feats, ref_masks=None, src_masks=None):
        n, m = ref_feats.shape[0], src_feats.shape[0]
        ri, si, sc, cnt = self.forward_pairs(ref_feats, src_feats, [n], [m], ref_masks, src_masks)
        k = int(cnt[0])  # min(num_correspondences, number of unmasked entries), as the reference
        return ri[0, :k], si[0, :k], sc[0, :k]
