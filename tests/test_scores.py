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
