"""Score-tensor construction (/root/reference/src/model/ldndmv.py:184-209): the numpy restatement
against fixtures generated from the reference's own modules (tests/golden/gen_golden_scores.py), and the CUDA path
against both."""
import os

import numpy as np
import pytest
import torch

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["scores_small", "scores_fmask", "scores_mid"]


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    z = _load(name)
    hm = z["head_mask"] if bool(z["function_mask"]) else None
    attach, dec, root, md, ma = oracle.dmv_scores(z["x1"], z["x2"], z["token"], z["dec_score"], z["root_score"], hm)
    np.testing.assert_allclose(attach, z["attach"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(dec, z["dec"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(root, z["root"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(md, z["merged_dec"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(ma, z["merged_attach"], rtol=0, atol=1e-5)
    # structure: -INF rows exactly where the reference has them, padding value of merge bit-identical
    assert np.array_equal(ma <= -1e11, z["merged_attach"] <= -1e11)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_forward_and_backward(name):
    from vlgae_b200.scores import dmv_scores, split_merged

    dev = torch.device("cuda:0")
    z = _load(name)
    t = {k: torch.from_numpy(v).to(dev) for k, v in z.items() if v.dtype != np.bool_ or k == "head_mask"}
    hm = t["head_mask"] if bool(z["function_mask"]) else None
    x1, x2, ds, rs = (t[k].clone().requires_grad_() for k in ("x1", "x2", "dec_score", "root_score"))
    md, ma = dmv_scores(x1, x2, t["token"], ds, rs, hm)
    # log-probabilities: 1e-5 absolute (fp32 dot products of r terms in a different order + a different log-sum-exp split)
    np.testing.assert_allclose(md.detach().cpu().numpy(), z["merged_dec"], rtol=0, atol=2e-6)
    got = ma.detach().cpu().numpy()
    assert np.array_equal(got <= -1e11, z["merged_attach"] <= -1e11)
    big = z["merged_attach"] <= -1e11
    assert np.array_equal(got[big], z["merged_attach"][big]), "fill values (merge's zero, -INF of masked heads) must be bit-identical"
    np.testing.assert_allclose(got[~big], z["merged_attach"][~big], rtol=0, atol=1e-5)
    dec, attach, root = split_merged(md, ma)
    np.testing.assert_allclose(attach.detach().cpu().numpy()[z["attach"] > -1e11], z["attach"][z["attach"] > -1e11], rtol=0, atol=1e-5)
    np.testing.assert_allclose(root.detach().cpu().numpy(), z["root"], rtol=0, atol=1e-5)
    g1, g2, gd, gr = torch.autograd.grad([md, ma], [x1, x2, ds, rs], [t["grad_merged_dec"], t["grad_merged_attach"]])
    for got_g, key in ((g1, "grad_x1"), (g2, "grad_x2"), (gd, "grad_dec_score"), (gr, "grad_root_score")):
        want = z[key]
        scale = max(1.0, float(np.abs(want).max()))
        np.testing.assert_allclose(got_g.cpu().numpy(), want, rtol=0, atol=2e-5 * scale, err_msg=key)


@pytest.mark.gpu
def test_cuda_matches_oracle_at_cfg2_shape_and_feeds_the_chart():
    """B = 128, n = 40, 10 k tokens, r = 16 (the shape of BASELINE.json configs[1]): forward against the numpy restatement
    on a slice, and the merged tensors go straight into DMV1o."""
    from vlgae_b200.scores import dmv_scores
    from vlgae_b200.torch_struct import DMV1o

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    B, n, T, r = 128, 40, 10000, 16
    x1 = torch.randn(B, n, 2, 2, r, generator=g) * 0.5
    x2 = torch.randn(T, 2, 2, r, generator=g) * 0.5
    token = torch.randint(0, T, (B, n), generator=g)
    ds = torch.randn(B, n, 2, 2, 2, generator=g)
    rs = torch.randn(T, generator=g)
    lengths = torch.randint(4, n + 1, (B,), generator=g).sort(descending=True).values
    x1d, x2d = x1.to(dev).requires_grad_(), x2.to(dev).requires_grad_()
    md, ma = dmv_scores(x1d, x2d, token.to(dev), ds.to(dev), rs.to(dev))
    nb = 4
    _, _, _, omd, oma = oracle.dmv_scores(x1[:nb].numpy(), x2.numpy(), token[:nb].numpy(), ds[:nb].numpy(), rs.numpy())
    np.testing.assert_allclose(md[:nb].detach().cpu().numpy(), omd, rtol=0, atol=2e-6)
    np.testing.assert_allclose(ma[:nb].detach().cpu().numpy(), oma, rtol=0, atol=2e-5)
    dist = DMV1o([md, ma], lengths.to(dev))
    Z = dist.partition
    assert torch.isfinite(Z).all()
    g1, g2 = torch.autograd.grad(Z.sum(), [x1d, x2d])
    assert torch.isfinite(g1).all() and torch.isfinite(g2).all()
    # a head beyond the end of its sentence takes part in no tree: its rows of the merged tensors get zero marginals
    pad = torch.arange(n, device=dev)[None, :] >= lengths.to(dev)[:, None]
    assert float(g1[pad].abs().max()) == 0.0, "heads beyond the sentence receive no gradient"


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,T,r", [(1, 1, 1, 4), (2, 3, 65, 8), (3, 5, 130, 32), (2, 40, 1000, 16), (5, 2, 7, 4)])
def test_cuda_edge_shapes_against_oracle(B, n, T, r):
    """Single word, single token, vocabulary sizes around the 64-token tile, every supported rank; forward against the
    numpy restatement, backward against autograd through the reference's torch formula in fp64."""
    from vlgae_b200.scores import dmv_scores

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(100 + B + n + T)
    x1 = torch.randn(B, n, 2, 2, r, generator=g)
    x2 = torch.randn(T, 2, 2, r, generator=g)
    token = torch.randint(0, T, (B, n), generator=g)
    ds = torch.randn(B, n, 2, 2, 2, generator=g)
    rs = torch.randn(T, generator=g)
    hm = torch.rand(B, n, generator=g) > 0.7
    t = [v.to(dev).requires_grad_() for v in (x1, x2, ds, rs)]
    md, ma = dmv_scores(t[0], t[1], token.to(dev), t[2], t[3], hm.to(dev))
    _, _, _, omd, oma = oracle.dmv_scores(x1.numpy(), x2.numpy(), token.numpy(), ds.numpy(), rs.numpy(), hm.numpy())
    np.testing.assert_allclose(md.detach().cpu().numpy(), omd, rtol=0, atol=5e-6)
    got = ma.detach().cpu().numpy()
    big = oma <= -1e11
    assert np.array_equal(got <= -1e11, big) and np.array_equal(got[big], oma[big])
    np.testing.assert_allclose(got[~big], oma[~big], rtol=0, atol=2e-5)
    gmd = torch.randn(md.shape, generator=g).to(dev)
    gma = torch.randn(ma.shape, generator=g).to(dev)
    mine = torch.autograd.grad([md, ma], t, [gmd, gma])
    # the reference's formula (ldndmv.py:185-209) in fp64
    d = [v.double().requires_grad_() for v in (x1, x2, ds, rs)]
    rule = torch.einsum("bhdve,cdve->bhcdv", d[0], d[1]).log_softmax(2)
    prob = rule.gather(2, token.reshape(B, 1, n, 1, 1).expand(B, n, n, 2, 2))
    lm = torch.tril(torch.ones(n, n, dtype=torch.float64), -1)[None, :, :, None]
    rm = torch.triu(torch.ones(n, n, dtype=torch.float64), 1)[None, :, :, None]
    att = (prob[..., 0, :] * lm + prob[..., 1, :] * rm).masked_fill(hm[:, :, None, None], -1e20)
    dec = d[2].permute(0, 1, 3, 4, 2).log_softmax(-1)
    root = d[3].log_softmax(-1)[token]
    rmd = torch.full((B, n + 1, 2, 2, 2), -1e12, dtype=torch.float64)
    rma = torch.full((B, n + 1, n + 1, 2), -1e12, dtype=torch.float64)
    rma[:, 0, 1:, 1] = root
    rma[:, 1:, 1:, :] = att
    rmd[:, 0, 1, :, :] = 0
    rmd[:, 1:] = dec
    theirs = torch.autograd.grad([rmd, rma], d, [gmd.double().cpu(), gma.double().cpu()])
    for a, b_, name in zip(mine, theirs, ("x1", "x2", "dec_score", "root_score")):
        scale = max(1.0, float(b_.abs().max()))
        np.testing.assert_allclose(a.cpu().numpy(), b_.numpy(), rtol=0, atol=5e-5 * scale, err_msg=name)


@pytest.mark.gpu
def test_cuda_rejects_what_it_does_not_provide():
    from vlgae_b200._lib import VlgaeError
    from vlgae_b200.scores import dmv_scores

    dev = torch.device("cuda:0")
    x1, x2 = torch.zeros(1, 2, 2, 2, 16, device=dev), torch.zeros(5, 2, 2, 16, device=dev)
    tok, ds, rs = torch.zeros(1, 2, dtype=torch.long, device=dev), torch.zeros(1, 2, 2, 2, 2, device=dev), torch.zeros(5, device=dev)
    with pytest.raises(VlgaeError):
        dmv_scores(x1, x2, tok, ds, rs, extended_valence=False)
    with pytest.raises(VlgaeError):
        dmv_scores(x1[..., :5], x2[..., :5], tok, ds, rs)  # rank 5 is not one of 4 / 8 / 16 / 32
    with pytest.raises(VlgaeError):
        dmv_scores(x1.cpu(), x2.cpu(), tok.cpu(), ds.cpu(), rs.cpu())
