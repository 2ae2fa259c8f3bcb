"""Visual factor features with the relation MLP collapsed (/root/reference/src/model/vis_encoder/
box_rel.py:42-52 + joint.py:140-179): the literal numpy restatement against fixtures generated from the reference's own MLP
modules (tests/golden/gen_golden_vis.py), and the CUDA path (collapsed) against both, forward and backward."""
import os

import numpy as np
import pytest
import torch

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["vis_small", "vis_noattr"]


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def _inputs(z):
    feat = z["feat"]
    if bool(z["img_feat"]):  # box_rel.py:35-40
        feat = np.concatenate([feat, np.broadcast_to(feat.mean(1, keepdims=True), feat.shape)], -1)
    return feat


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    z = _load(name)
    attr = (z["w_attr"], z["b_attr"]) if bool(z["use_attr"]) else (None, None)
    mid, mask = oracle.vis_factors(_inputs(z), z["w_box"], z["b_box"], z["w_rel"], z["b_rel"], attr[0], attr[1], z["box_mask"],
                                   add_image=bool(z["add_image"]))
    np.testing.assert_allclose(mid, z["mid"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(mask, z["vis_mask"])
    np.testing.assert_allclose(mid.astype(np.float64) @ z["w_pre"].astype(np.float64).T, z["vis"], rtol=1e-4, atol=1e-5)


class _MLP(torch.nn.Module):  # the attribute layout of the reference's MLP (nn/common.py:23-51)
    def __init__(self, w, b):
        super().__init__()
        self.linear = torch.nn.Linear(w.shape[1], w.shape[0])
        with torch.no_grad():
            self.linear.weight.copy_(torch.from_numpy(w)); self.linear.bias.copy_(torch.from_numpy(b))
        self.activation = torch.nn.LeakyReLU()
        self.dropout = torch.nn.Identity()


class _Encoder(torch.nn.Module):
    def __init__(self, z):
        super().__init__()
        self.img_feat, self.use_attr = bool(z["img_feat"]), bool(z["use_attr"])
        self.box_fc, self.rel_fc = _MLP(z["w_box"], z["b_box"]), _MLP(z["w_rel"], z["b_rel"])
        if self.use_attr:
            self.attr_fc = _MLP(z["w_attr"], z["b_attr"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_forward_and_backward(name):
    from vlgae_b200.vis_factors import vis_feat_unprune_collapsed

    dev = torch.device("cuda:0")
    z = _load(name)
    enc = _Encoder(z).to(dev)
    pre = torch.nn.Linear(z["w_pre"].shape[1], z["w_pre"].shape[0], bias=False).to(dev)
    with torch.no_grad():
        pre.weight.copy_(torch.from_numpy(z["w_pre"]))
    feat = torch.from_numpy(z["feat"]).to(dev).requires_grad_()
    torch.backends.cuda.matmul.allow_tf32 = False
    vis, mask, split, mid = vis_feat_unprune_collapsed(enc, pre, feat, torch.from_numpy(z["box_mask"]).to(dev),
                                                       add_image=bool(z["add_image"]), return_mid=True)
    assert vis.names == ("A", "V", "Y") and mask.names == ("A", "V")
    assert split == z["split"].tolist()
    np.testing.assert_allclose(mid.detach().cpu().numpy(), z["mid"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(vis.rename(None).detach().cpu().numpy(), z["vis"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(mask.rename(None).cpu().numpy(), z["vis_mask"])
    params = [feat, enc.box_fc.linear.weight, enc.box_fc.linear.bias, enc.rel_fc.linear.weight, enc.rel_fc.linear.bias, pre.weight]
    keys = ["g_feat", "g_w_box", "g_b_box", "g_w_rel", "g_b_rel", "g_w_pre"]
    if enc.use_attr:
        params += [enc.attr_fc.linear.weight, enc.attr_fc.linear.bias]
        keys += ["g_w_attr", "g_b_attr"]
    grads = torch.autograd.grad([vis.rename(None), mid], params,
                                [torch.from_numpy(z["grad_vis"]).to(dev), torch.from_numpy(z["grad_mid"]).to(dev)])
    for g, k in zip(grads, keys):
        want = z[k]
        np.testing.assert_allclose(g.cpu().numpy(), want, rtol=2e-4, atol=2e-5 * max(1.0, float(np.abs(want).max())), err_msg=k)


@pytest.mark.gpu
def test_cuda_matches_literal_oracle_at_cfg2_shape():
    """36 boxes, H = 256 (the VLParse shape; a slice of the batch goes through the LITERAL n^2-pair restatement)."""
    from vlgae_b200.vis_factors import pairwise_factors

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    B, n, F, H = 16, 36, 64, 256
    x = torch.randn(B, n, F, generator=g)
    w = [torch.randn(H, F, generator=g) * 0.2 for _ in range(3)]
    b = [torch.randn(H, generator=g) * 0.1 for _ in range(3)]
    bm = torch.rand(B, n, generator=g) > 0.2
    u = [(x @ wi.T + bi).to(dev) for wi, bi in zip(w, b)]
    mid, mask, split = pairwise_factors(u[0], u[1], u[2], bm.to(dev))
    assert split == [n, n * n, n, 1] and mid.shape == (B, n + n * n + n + 1, H)
    nb = 2
    omid, omask = oracle.vis_factors(x[:nb].numpy(), w[0].numpy(), b[0].numpy(), w[1].numpy(), b[1].numpy(), w[2].numpy(), b[2].numpy(),
                                     bm[:nb].numpy())
    np.testing.assert_allclose(mid[:nb].cpu().numpy(), omid, rtol=1e-4, atol=2e-5)
    assert np.array_equal(mask[:nb].cpu().numpy(), omask)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,H,attr,img", [(1, 1, 5, True, True), (2, 3, 7, False, True), (3, 4, 130, True, False), (2, 2, 64, False, False)])
def test_cuda_edge_shapes_against_torch_formula(B, n, H, attr, img):
    """One box, channel counts that are not multiples of 4 (scalar path), every combination of the optional factor groups;
    forward and backward against the pair formula in torch fp64."""
    from vlgae_b200.vis_factors import pairwise_factors

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7 + B + n + H)
    u = [torch.randn(B, n, H, generator=g) for _ in range(3)]
    bm = torch.rand(B, n, generator=g) > 0.3
    t = [v.to(dev).requires_grad_() for v in u]
    mid, mask, split = pairwise_factors(t[0], t[1], t[2] if attr else None, bm.to(dev), add_image=img)
    d = [v.double().requires_grad_() for v in u]
    act = torch.nn.functional.leaky_relu
    box = act(d[0], 0.01)
    rel = act((d[1].unsqueeze(1) + d[1].unsqueeze(2)) / 2, 0.01).reshape(B, n * n, H)
    feats = [box, rel] + ([act(d[2], 0.01)] if attr else []) + ([box.mean(1, keepdim=True)] if img else [])
    ref = torch.cat(feats, 1)
    rmask = torch.cat([bm, (bm.unsqueeze(1) & bm.unsqueeze(2)).triu(1).reshape(B, -1)] + ([bm] if attr else [])
                      + ([torch.ones(B, 1, dtype=torch.bool)] if img else []), 1)
    assert split == [n, n * n] + ([n] if attr else []) + ([1] if img else [])
    np.testing.assert_allclose(mid.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(mask.cpu(), rmask)
    go = torch.randn(ref.shape, generator=g)
    ins = t if attr else t[:2]
    mine = torch.autograd.grad(mid, ins, go.to(dev))
    theirs = torch.autograd.grad(ref, d if attr else d[:2], go.double())
    for a, b_ in zip(mine, theirs):
        np.testing.assert_allclose(a.cpu().numpy(), b_.numpy(), rtol=1e-5, atol=1e-5)
