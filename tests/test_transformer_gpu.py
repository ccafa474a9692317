"""GPU parity of the superpoint transformer and SuperPointMatching against the torch-fp32 oracle
(oracle/transformer.py) and the reference fixtures.  Tolerances: bf16 tensor-core operands with fp32 accumulation /
softmax => per-op rtol 2e-2 on unit-scale outputs, whole transformer cosine >= 0.999 per superpoint; matching scores
rtol 1e-4 (fp32), indices bit-exact given identical score matrices."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import transformer as ot
from se3et_b200.modules import transformer as MT
from se3et_b200.ops import transformer_ops as T
from test_oracle_transformer import coarse_inputs, transformer_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "model_small.npz"))


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def build_transformer(in_dim=256):
    S = helpers.SMALL_CFG
    net = MT.GeometricTransformer(in_dim, S["tr_output_dim"], S["hidden_dim"], S["num_heads"], S["blocks"], S["sigma_d"],
                                  S["sigma_a"], S["angle_k"], na=6)
    sd = transformer_state_dict(in_dim)
    own = net.state_dict()
    mine = {k[len("transformer."):]: v for k, v in sd.items()}
    assert set(mine) | {"embedding.embedding.div_term"} == set(own), set(own) ^ set(mine)
    net.load_state_dict(mine, strict=False)
    return net.to(DEV).eval(), sd


def test_geometric_embedding_matches_oracle_and_reference(gold):
    S = helpers.SMALL_CFG
    rp, sp, _, _ = coarse_inputs(gold)
    net, sd = build_transformer()
    with torch.no_grad():
        e = net.embedding(rp[None].to(DEV))[0].cpu()
    p = ot.Params(sd, "transformer.embedding.")
    want = ot.geometric_structure_embedding(p, rp, S["hidden_dim"], S["sigma_d"], S["sigma_a"], S["angle_k"])
    # bf16 output (8 mantissa bits) of an O(1..5)-magnitude embedding
    assert torch.allclose(e, want, rtol=2e-2, atol=3e-2), (e - want).abs().max().item()
    assert rel_err(e, want) < 6e-3
    assert rel_err(e, torch.from_numpy(gold["ref_embedding"].astype(np.float32))) < 6e-3


def test_embedding_indices_match_oracle():
    g = torch.Generator().manual_seed(0)
    pts = [torch.rand(n, 3, generator=g) * 2 for n in (37, 5, 130)]
    flat = torch.cat(pts).to(DEV)
    ctx = MT.CloudContext([37, 5], [130], 1, 1, DEV)
    idx4 = T.geo_embed_indices(flat, ctx.cu, ctx.max_n, ctx.eoff, ctx.R, 0.2, 15.0, 3).cpu()
    o = 0
    for p in pts:
        n = p.shape[0]
        d, a = ot.embedding_indices(p, 0.2, 15.0, 3)
        got = idx4[o:o + n * n].view(n, n, 4)
        # the reference's d^2 = x^2 - 2xy + y^2 cancels near the diagonal: compare d with an absolute tolerance
        assert torch.allclose(got[..., 0], d, rtol=1e-4, atol=5e-3)
        off = ~torch.eye(n, dtype=torch.bool)
        assert torch.allclose(got[..., 1:][off], a[off], rtol=1e-3, atol=2e-3)
        o += n * n


def test_flash_attention_matches_torch():
    g = torch.Generator().manual_seed(1)
    for d, h, a, sizes in ((16, 4, 6, [(70, 70), (33, 33)]), (64, 4, 6, [(130, 97)]), (32, 2, 1, [(64, 200), (5, 3)])):
        c = d * h
        nq, nk = sum(s[0] for s in sizes), sum(s[1] for s in sizes)
        q = torch.randn(nq, a, c, generator=g).bfloat16()
        k = torch.randn(nk, a, c, generator=g).bfloat16()
        v = torch.randn(nk, a, c, generator=g).bfloat16()
        bias_parts, problems, boff, qs, ks = [], [], 0, 0, 0
        for n, m in sizes:
            bias_parts.append(torch.randn(n, a, h, m, generator=g))
            problems.append([qs, n, ks, m, boff])
            boff += n * a * h * m
            qs += n
            ks += m
        bias = torch.cat([b.reshape(-1) for b in bias_parts])
        out = torch.zeros(nq * a, c, dtype=torch.bfloat16, device=DEV)
        T.flash_attention(q.to(DEV), a * c, c, k.to(DEV), a * c, c, v.to(DEV), a * c, c, bias.to(DEV),
                          torch.tensor(problems, dtype=torch.int64, device=DEV), max(s[0] for s in sizes), a, h, d, out)
        out = out.float().cpu().view(nq, a, h, d)
        qs = ks = 0
        for (n, m), b in zip(sizes, bias_parts):
            qq = q[qs:qs + n].float().view(n, a, h, d)
            kk = k[ks:ks + m].float().view(m, a, h, d)
            vv = v[ks:ks + m].float().view(m, a, h, d)
            s = (torch.einsum("nahc,mahc->nahm", qq, kk) + b) / d ** 0.5
            want = torch.einsum("nahm,mahc->nahc", torch.softmax(s, -1), vv)
            got = out[qs:qs + n]
            assert torch.allclose(got, want, rtol=2e-2, atol=2e-2), (got - want).abs().max().item()
            assert rel_err(got, want) < 1e-2
            qs += n
            ks += m


def test_add_layernorm_and_l2_normalize():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(301, 64, generator=g)
    r = torch.randn(51, 64, generator=g).bfloat16()
    gamma, beta = torch.randn(64, generator=g), torch.randn(64, generator=g)
    x6 = torch.randn(306, 64, generator=g)
    of, ob = T.add_layernorm(x6.to(DEV), r.to(DEV), 6, gamma.to(DEV), beta.to(DEV), out_f32=True)
    want = torch.nn.functional.layer_norm(x6 + r.float().repeat_interleave(6, 0), (64,), gamma, beta)
    assert torch.allclose(of.cpu(), want, rtol=1e-4, atol=1e-4)
    assert torch.allclose(ob.float().cpu(), want, rtol=1e-2, atol=2e-2)
    n = T.l2_normalize_rows(x.to(DEV)).cpu()
    assert torch.allclose(n, torch.nn.functional.normalize(x, p=2, dim=1), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rows,k,div", [(1000, 256, 1), (777, 512, 1), (6 * 301, 256, 6), (128, 64, 1), (5, 256, 1)])
def test_linear_add_layernorm_matches_torch(rows, k, div):
    """se3et_linear_add_layernorm (Linear + residual + LayerNorm in one tcgen05 kernel) against torch fp32 on the same
    bf16 operands: ragged last tile, both K of the model (256: `linear`, 512: `squeeze`), the (N, C) residual broadcast
    over anchors (resid_div = 6); and against the two-kernel form it replaces."""
    g = torch.Generator().manual_seed(rows + k)
    a = (torch.randn(rows, k, generator=g) * 0.7).bfloat16()
    w = (torch.randn(256, k, generator=g) / k ** 0.5).bfloat16()
    b = torch.randn(256, generator=g) * 0.1
    resid = (torch.randn((rows + div - 1) // div, 256, generator=g) * 1.5 + 0.3).bfloat16()
    gamma, beta = torch.randn(256, generator=g), torch.randn(256, generator=g)
    y = a.float() @ w.float().t() + b + resid.float().repeat_interleave(div, 0)[:rows]
    want = torch.nn.functional.layer_norm(y, (256,), gamma, beta, 1e-5)
    got = T.linear_add_layernorm(a.to(DEV), w.to(DEV), b.to(DEV), resid.to(DEV), div, gamma.to(DEV), beta.to(DEV), 1e-5)
    assert got.dtype == torch.bfloat16 and got.shape == (rows, 256)
    assert torch.allclose(got.float().cpu(), want, rtol=2e-2, atol=2e-2), (got.float().cpu() - want).abs().max()
    from se3et_b200.ops.gemm import linear_bf16
    z, _ = linear_bf16(a.to(DEV), w.to(DEV), b.to(DEV))
    _, two = T.add_layernorm(z, resid.to(DEV), div, gamma.to(DEV), beta.to(DEV), 1e-5)
    assert torch.allclose(got.float(), two.float(), rtol=2e-2, atol=2e-2)


def test_transformer_matches_oracle_and_reference(gold):
    S = helpers.SMALL_CFG
    rp, sp, rf, sf = coarse_inputs(gold)
    net, sd = build_transformer()
    with torch.no_grad():
        r, s, *_ = net(rp[None].to(DEV), sp[None].to(DEV), rf[None].to(DEV), sf[None].to(DEV))
    r, s = r[0].cpu(), s[0].cpu()
    wr, ws, _, _ = ot.geometric_transformer(sd, rp, sp, rf, sf, S["blocks"], S["hidden_dim"], S["num_heads"],
                                            S["sigma_d"], S["sigma_a"], S["angle_k"])
    for got, want, ref in ((r, wr, gold["ref_feats_c"]), (s, ws, gold["src_feats_c"])):
        cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
        assert cos.min().item() > 0.999, cos.min().item()
        assert rel_err(got, want) < 2e-2, rel_err(got, want)
        assert rel_err(got, torch.from_numpy(ref)) < 2e-2


def test_transformer_batched_pairs_equal_single_pairs(gold):
    rp, sp, rf, sf = coarse_inputs(gold)
    net, _ = build_transformer()
    g = torch.Generator().manual_seed(5)
    rp2, sp2 = torch.rand(40, 3, generator=g), torch.rand(23, 3, generator=g)
    rf2, sf2 = torch.randn(40, 6, 256, generator=g), torch.randn(23, 6, 256, generator=g)
    with torch.no_grad():
        a = net.forward_clouds(torch.cat([rp, sp]).to(DEV), torch.cat([rf, sf]).to(DEV), [len(rp)], [len(sp)])
        b = net.forward_clouds(torch.cat([rp2, sp2]).to(DEV), torch.cat([rf2, sf2]).to(DEV), [40], [23])
        both = net.forward_clouds(torch.cat([rp, rp2, sp, sp2]).to(DEV), torch.cat([rf, rf2, sf, sf2]).to(DEV),
                                  [len(rp), 40], [len(sp), 23])
    n, m = len(rp), len(sp)
    want = torch.cat([a[:n], b[:40], a[n:], b[40:]])
    assert rel_err(both, want) < 2e-3, rel_err(both, want)


def test_superpoint_matching_matches_oracle_and_reference(gold):
    rf, sf = torch.from_numpy(gold["spm_ref_feats"]), torch.from_numpy(gold["spm_src_feats"])
    rm, sm = torch.from_numpy(gold["spm_ref_masks"]), torch.from_numpy(gold["spm_src_masks"])
    spm = MT.SuperPointMatching(64, True)
    ri, si, sc = spm(rf.to(DEV), sf.to(DEV), rm.to(DEV), sm.to(DEV))
    ri, si, sc = ri.cpu(), si.cpu(), sc.cpu()
    wr, ws, wsc = ot.superpoint_matching(rf, sf, rm, sm, 64, True)
    assert torch.allclose(sc, wsc, rtol=1e-4, atol=1e-12)
    assert torch.allclose(sc, torch.from_numpy(gold["spm_scores"]), rtol=1e-4, atol=1e-12)
    # indices: identical wherever neighbouring scores are separated by more than the fp32 noise of the two pipelines
    gap = (wsc[:-1] - wsc[1:]) > 2e-4 * wsc[:-1]
    sep = torch.cat([torch.tensor([True]), gap]) & torch.cat([gap, torch.tensor([True])])
    assert torch.equal(ri[sep], wr[sep]) and torch.equal(si[sep], ws[sep])
    assert set(zip(ri.tolist(), si.tolist())) == set(zip(gold["spm_ref_idx"].tolist(), gold["spm_src_idx"].tolist()))


def test_superpoint_selection_is_exact_on_its_own_scores():
    """Integer part of the op: the CUDA top-k equals the canonical (score desc, flat index asc) top-k of the CUDA
    score matrix, bit for bit -- including ties (quantised features produce many equal scores) and masks."""
    g = torch.Generator().manual_seed(7)
    sizes_r, sizes_s = [57, 410, 8], [300, 33, 5]
    c = 32
    rf = torch.nn.functional.normalize(torch.randint(-2, 3, (sum(sizes_r), c), generator=g).float() + 0.01, dim=1)
    sf = torch.nn.functional.normalize(torch.randint(-2, 3, (sum(sizes_s), c), generator=g).float() + 0.01, dim=1)
    rm = torch.rand(sum(sizes_r), generator=g) > 0.1
    sm = torch.rand(sum(sizes_s), generator=g) > 0.1
    spm = MT.SuperPointMatching(256, True)
    rs, ss = np.asarray(sizes_r), np.asarray(sizes_s)
    rcu, scu = np.concatenate([[0], np.cumsum(rs)]), np.concatenate([[0], np.cumsum(ss)])
    eo = np.concatenate([[0], np.cumsum(rs * ss)])
    problems = torch.tensor(np.stack([rcu[:-1], rs, scu[:-1], ss, eo[:-1]], 1), dtype=torch.int64, device=DEV)
    ri, si, sc, cnt, e = T.superpoint_matching(rf.to(DEV), sf.to(DEV), rm.to(DEV), sm.to(DEV), problems, int(rs.max()),
                                               int(ss.max()), int(eo[-1]), 256, True)
    ri, si, sc, cnt, e = ri.cpu(), si.cpu(), sc.cpu(), cnt.cpu(), e.cpu()
    for p in range(3):
        n, m = sizes_r[p], sizes_s[p]
        E = e[eo[p]:eo[p + 1]].view(n, m)
        r_on, s_on = rm[rcu[p]:rcu[p + 1]], sm[scu[p]:scu[p + 1]]
        score = (E / E.sum(1, keepdim=True)) * (E / E.sum(0, keepdim=True))
        score = torch.where(r_on[:, None] & s_on[None, :], score, torch.full_like(score, -1.0))
        k = min(256, int(r_on.sum()) * int(s_on.sum()))
        assert int(cnt[p]) == k
        order = torch.argsort(-score.reshape(-1), stable=True)[:k]
        assert torch.equal(ri[p, :k], order // m) and torch.equal(si[p, :k], order % m)
        assert torch.allclose(sc[p, :k], score.reshape(-1)[order], rtol=1e-5)
        assert (ri[p, k:] == -1).all()


# ---- SE3ET-E block list on the CUDA path --------------------------------------------------------------------------
BLOCKS_E = ['self_eq', 'cross_a_soft', 'self_eq', 'cross_r_soft', 'self', 'cross']


@pytest.fixture(scope="module")
def gold_e(golden_dir):
    return np.load(os.path.join(golden_dir, "model_e_small.npz"))


def test_anchor_statistics_and_mixing_match_torch():
    """anchor_pair_stats / anchor_mix_weights / anchor_mix against the dense torch formulation, two problems."""
    from se3et_b200.ops import transformer_ops as T
    from oracle import transformer as ot
    g = torch.Generator().manual_seed(5)
    a, c, h = 6, 64, 4
    nq, nk = [37, 70], [50, 21]
    q = (torch.randn(sum(nq), a, c, generator=g) * 0.5).bfloat16()
    k = (torch.randn(sum(nk), a, c, generator=g) * 0.5).bfloat16()
    qo, ko = np.concatenate([[0], np.cumsum(nq)]), np.concatenate([[0], np.cumsum(nk)])
    problems = torch.tensor([[qo[i], nq[i], ko[i], nk[i], 0] for i in range(2)], dtype=torch.int64, device=DEV)
    perms = ot.octahedral_rotation_perms()
    qd, kd = q.to(DEV).view(-1, c), k.to(DEV).view(-1, c)
    for positive in ("sq", "softplus"):
        got = T.anchor_pair_stats(qd, a * c, c, kd, a * c, c, problems, max(nq), a, c, h, positive).cpu()
        for i in range(2):
            qq, kk = q[qo[i]:qo[i + 1]].float(), k[ko[i]:ko[i + 1]].float()
            s = torch.einsum("nac,mec->aenm", qq, kk) / (h * (c // h) ** 0.5)
            want = (s ** 2 if positive == "sq" else torch.nn.functional.softplus(s)).sum((-2, -1))
            assert torch.allclose(got[i], want, rtol=2e-3, atol=1e-3), positive
            for r_soft in (False, True):
                w, attn_r = T.anchor_mix_weights(got.to(DEV), problems, perms.to(torch.int32).to(DEV), r_soft)
                ww, ar = ot.anchor_mixing_weights(want / (nq[i] * nk[i]), "r_soft" if r_soft else "a_soft", perms)
                assert torch.allclose(w[i].cpu(), ww, rtol=2e-3, atol=1e-5)
                if r_soft:
                    assert torch.allclose(attn_r[i].cpu(), ar, rtol=2e-3, atol=1e-6)
    w = torch.rand(2, a, a, generator=g)
    x = torch.randn(a, sum(nq), a, c, generator=g).bfloat16()
    cloud_off = torch.tensor(qo, dtype=torch.int64, device=DEV)
    out = T.anchor_mix(x.to(DEV), sum(nq) * a * c, a * c, c, w.to(DEV), cloud_off, a, c, sum(nq)).cpu().float()
    wn = torch.cat([w[0][None].expand(nq[0], -1, -1), w[1][None].expand(nq[1], -1, -1)])  # (n, a, e)
    want = torch.einsum("nae,enac->nac", wn, x.float()).reshape(-1, c)
    assert torch.allclose(out, want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("tag,nlev", [("nosh", 0), ("sh", 2)])
def test_transformer_e_matches_oracle_and_reference(gold, gold_e, tag, nlev):
    """GeometricTransformer with the SE3ET-E schedule (bf16 tensor cores) vs the fp32 oracle and the reference fixture."""
    from test_oracle_transformer import coarse_inputs, transformer_e_state_dict
    from oracle import transformer as ot
    S = helpers.SMALL_CFG
    rp, sp, rf, sf = coarse_inputs(gold)
    sd = transformer_e_state_dict(gold_e, prefix="")
    if nlev == 0:
        sd = {k: v for k, v in sd.items() if "proj_eq" not in k}
    tr = MT.GeometricTransformer(16 * S["init_dim"], S["tr_output_dim"], S["hidden_dim"], S["num_heads"], BLOCKS_E,
                                 S["sigma_d"], S["sigma_a"], S["angle_k"], na=6, n_level_equiv=nlev)
    missing, unexpected = tr.load_state_dict(sd, strict=False)
    assert not unexpected and all(("anchors" in k or "trace_idx" in k or "div_term" in k) for k in missing), missing
    tr = tr.to(DEV).eval()
    with torch.no_grad():
        r, s, *_ = tr(rp[None].to(DEV), sp[None].to(DEV), rf[None].to(DEV), sf[None].to(DEV))
    want_r, want_s = torch.from_numpy(gold_e["ref_feats_" + tag]), torch.from_numpy(gold_e["src_feats_" + tag])
    cos = torch.nn.functional.cosine_similarity
    assert cos(r[0].float().cpu(), want_r, dim=-1).min() > 0.99
    assert cos(s[0].float().cpu(), want_s, dim=-1).min() > 0.99
    assert rel_err(r[0].float().cpu(), want_r) < 5e-2 and rel_err(s[0].float().cpu(), want_s) < 5e-2


def test_tabulated_embedding_equals_the_sinusoid_gemm(gold):
    """se3et_geo_embed_lookup (tabulated W emb(u) + b, step 1/512) against se3et_geo_embed_project (sinusoids generated
    in the kernel, tcgen05 GEMM) and the fp32 oracle on the same superpoints."""
    S = helpers.SMALL_CFG
    rp, sp, _, _ = coarse_inputs(gold)
    net, sd = build_transformer()
    pts = rp[None].to(DEV)
    MT._EMBED_TABLE['on'] = False
    try:
        exact = net.embedding(pts)[0].cpu()
    finally:
        MT._EMBED_TABLE['on'] = True
    table = net.embedding(pts)[0].cpu()
    p = ot.Params(sd, "transformer.embedding.")
    want = ot.geometric_structure_embedding(p, rp, S["hidden_dim"], S["sigma_d"], S["sigma_a"], S["angle_k"])
    assert rel_err(table, want) < 6e-3 and rel_err(exact, want) < 6e-3, (rel_err(table, want), rel_err(exact, want))
    assert torch.allclose(table, want, rtol=2e-2, atol=2e-2)
    assert rel_err(table, exact) < 8e-3
