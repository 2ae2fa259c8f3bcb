"""Parity of the tcgen05 alignment kernel (through the C ABI) with the oracle / the reference's golden vectors."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def run(vis, vm, txt, tm, dev, split=3, pad_rows=True):
    from vlgae_b200.alignment import gather_logit_simple

    out = gather_logit_simple(_t(vis, dev), _t(vm, dev), _t(txt, dev), _t(tm, dev), split=split, pad_rows=pad_rows)
    assert out.is_contiguous() or pad_rows
    torch.cuda.synchronize()
    assert out.names == ("B", "A", "Q", "V")
    return out.rename(None).cpu().numpy()


def check(got, want, vis, txt, split):
    masked = want == -1e20
    assert ((got == -1e20) == masked).all(), "mask pattern differs"
    # error model of the split-bf16 product: 3 terms kept -> relative 2^-16 per product, accumulated in fp32
    scale = np.sqrt((vis.astype(np.float64) ** 2).sum(-1)).max() * np.sqrt((txt.astype(np.float64) ** 2).sum(-1)).max()
    tol = scale * (2.0 ** -15 if split == 3 else 2.0 ** -7)
    err = np.abs(got - want)[~masked].max() if (~masked).any() else 0.0
    assert err <= tol, (err, tol)
    return err


@pytest.mark.parametrize("name", ["align_small", "align_mid"])
@pytest.mark.parametrize("split", [3, 1])
def test_golden_reference_vectors(golden, dev, name, split):
    g = golden(name)
    got = run(g["vis_feat"], g["vis_mask"], g["txt_feat"], g["txt_mask"], dev, split)
    check(got, g["attmap"], g["vis_feat"], g["txt_feat"], split)
    if split == 3:
        np.testing.assert_allclose(got.max(-1), g["max_v"], rtol=1e-4, atol=2e-3)
        np.testing.assert_allclose(got.max(2), g["max_q"], rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("A,V,B,Q,D", [(2, 1369, 3, 82, 128), (1, 128, 1, 16, 64), (3, 129, 2, 130, 128),
                                       (5, 300, 7, 33, 100), (2, 40, 2, 1, 8)])
def test_shapes_against_oracle(dev, A, V, B, Q, D):
    rng = np.random.default_rng(A * 1000 + V)
    vis = rng.normal(size=(A, V, D)).astype(np.float32)
    txt = rng.normal(size=(B, Q, D)).astype(np.float32)
    vm = rng.random((A, V)) > 0.2
    tm = rng.random((B, Q)) > 0.2
    want = oracle.gather_logit_simple(vis, vm, txt, tm)
    got = run(vis, vm, txt, tm, dev)
    check(got, want, vis, txt, 3)
    got = run(vis, vm, txt, tm, dev, pad_rows=False)
    check(got, want, vis, txt, 3)


def test_reduced_and_drop_in_signature(golden, dev):
    from vlgae_b200.alignment import gather_logit_reduced_impl, gather_logit_simple_impl

    g = golden("align_mid")
    vis = (_t(g["vis_feat"], dev).refine_names("A", "V", "D"), _t(g["vis_mask"], dev).refine_names("A", "V"), None)
    txt = (_t(g["txt_feat"], dev).refine_names("B", "Q", "D"), _t(g["txt_mask"], dev).refine_names("B", "Q"),
           _t(g["txt_marginal"], dev))
    att = gather_logit_simple_impl(None, {}, vis, txt, None)
    assert att.names == ("B", "A", "Q", "V")
    logit = att.max("V").values.log_softmax("A")  # the consumer's first steps (joint.py:473-476)
    assert logit.names == ("B", "A", "Q")
    att.rename(None)[0, 0, 0, 0] = 1.0  # writable, not aliased
    red = gather_logit_reduced_impl(None, {}, vis, txt, None)
    np.testing.assert_allclose(red.cpu().numpy(), g["reduced"], rtol=1e-4, atol=2e-3)


def test_gradients_match_reference_formula(dev):
    """The reference's attmap is differentiable (einsum + masked_fill_, joint.py:413-418); so is the replacement."""
    from vlgae_b200.alignment import gather_logit_reduced, gather_logit_simple

    g = torch.Generator(device=dev).manual_seed(3)
    A, V, B, Q, D = 3, 150, 4, 20, 96
    vis = torch.randn(A, V, D, generator=g, device=dev)
    txt = torch.randn(B, Q, D, generator=g, device=dev)
    vm = torch.rand(A, V, generator=g, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g, device=dev) > 0.2
    marg = torch.rand(B, Q, generator=g, device=dev) + 0.1
    w = torch.randn(B, A, Q, V, generator=g, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False

    def reference(v, t):
        att = torch.einsum("avd,bqd->baqv", v, t)
        att = att.masked_fill(~vm[None, :, None, :], -1e20).masked_fill(~tm[:, None, :, None], -1e20)
        return att

    v1, t1 = vis.clone().requires_grad_(), txt.clone().requires_grad_()
    v2, t2 = vis.clone().requires_grad_(), txt.clone().requires_grad_()
    keep = vm[None, :, None, :] & tm[:, None, :, None]
    (gather_logit_simple(v1, vm, t1, tm, named=False) * w * keep).sum().backward()
    (reference(v2, t2) * w * keep).sum().backward()
    np.testing.assert_allclose(v1.grad.cpu().numpy(), v2.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(t1.grad.cpu().numpy(), t2.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    # the reduced form (max over V, marginal-weighted mean over Q) back-propagates through the same node
    v3, t3 = vis.clone().requires_grad_(), txt.clone().requires_grad_()
    v4, t4 = vis.clone().requires_grad_(), txt.clone().requires_grad_()
    gather_logit_reduced(v3, vm, t3, tm, marg).sum().backward()
    att = reference(v4, t4)
    (torch.sum(att.max(dim=-1).values * marg.unsqueeze(1), dim=-1) / marg.sum(1, keepdim=True)).sum().backward()
    np.testing.assert_allclose(v3.grad.cpu().numpy(), v4.grad.cpu().numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(t3.grad.cpu().numpy(), t4.grad.cpu().numpy(), rtol=1e-3, atol=1e-3)
    # no graph is built when nothing requires grad
    assert not gather_logit_simple(vis, vm, txt, tm, named=False).requires_grad


@pytest.mark.parametrize("A,V,B,Q,D", [(2, 1369, 3, 82, 128), (3, 129, 2, 130, 128), (5, 300, 7, 33, 100), (2, 40, 2, 1, 8),
                                       (9, 700, 20, 50, 64)])
def test_max_over_factors_equals_max_of_materialised(dev, A, V, B, Q, D):
    """The fused epilogue (no [B,A,Q,V] tensor) gives bit for bit the max over V of the materialised logits, and the
    arg-max it reports attains that max at the smallest factor index."""
    from vlgae_b200.alignment import gather_logit_simple, max_over_factors

    g = torch.Generator(device=dev).manual_seed(A * 100 + Q)
    vis = torch.randn(A, V, D, generator=g, device=dev)
    txt = torch.randn(B, Q, D, generator=g, device=dev)
    vm = torch.rand(A, V, generator=g, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g, device=dev) > 0.2
    vm[0, :] = False  # an image without any valid factor: every max is the fill value
    for split in (3, 1):
        att = gather_logit_simple(vis, vm, txt, tm, split=split, named=False)
        maxv, argv = max_over_factors(vis, vm, txt, tm, split=split)
        ref = att.max(dim=-1)
        assert torch.equal(maxv, ref.values)
        picked = att.gather(-1, argv.long().unsqueeze(-1)).squeeze(-1)
        assert torch.equal(picked, ref.values)
        first = (att == ref.values.unsqueeze(-1)).int().argmax(dim=-1)   # first index attaining the maximum
        keep = tm.view(B, 1, Q).expand(B, A, Q)
        assert torch.equal(argv.long()[keep], first[keep])
    # against the oracle (fp32 einsum + masks) within the split-bf16 error model
    want = oracle.gather_logit_simple(vis.cpu().numpy(), vm.cpu().numpy(), txt.cpu().numpy(), tm.cpu().numpy()).max(-1)
    got = max_over_factors(vis, vm, txt, tm)[0].cpu().numpy()
    masked = want == -1e20
    assert ((got == -1e20) == masked).all()
    scale = float(vis.norm(dim=-1).max() * txt.norm(dim=-1).max())
    assert np.abs(got - want)[~masked].max(initial=0.0) <= scale * 2.0 ** -15


@pytest.mark.parametrize("B,V,n,D,H", [(4, 300, 12, 128, 256), (3, 1369, 40, 128, 256), (2, 33, 1, 16, 8), (2, 31, 64, 64, 100)])
def test_word_factor_attention(dev, B, V, n, D, H):
    """Row a10 (joint.py:668-673): the online-softmax kernel against the numpy restatement, and its backward against
    autograd through the reference's own two einsums + softmax evaluated in fp64."""
    from vlgae_b200.alignment import word_factor_attention

    g = torch.Generator(device=dev).manual_seed(5)
    vis = (torch.randn(B, V, D, generator=g, device=dev) * 0.3).requires_grad_()
    txt = (torch.randn(B, n, D, generator=g, device=dev) * 0.3).requires_grad_()
    mid = torch.randn(B, V, H, generator=g, device=dev).requires_grad_()
    out = word_factor_attention(vis, txt, mid)
    want = oracle.word_factor_attention(vis.detach().cpu().numpy(), txt.detach().cpu().numpy(), mid.detach().cpu().numpy())
    np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=1e-4, atol=1e-5)
    go = torch.randn(B, n, H, generator=g, device=dev)
    gv, gt, gm = torch.autograd.grad(out, [vis, txt, mid], go)
    v64, t64, m64 = (x.detach().double().requires_grad_() for x in (vis, txt, mid))
    ref = torch.einsum("bqv,bvh->bqh", torch.einsum("bvd,bqd->bqv", v64, t64).softmax(2), m64)  # joint.py:670-673
    rv, rt, rm = torch.autograd.grad(ref, [v64, t64, m64], go.double())
    for got, exp in ((gv, rv), (gt, rt), (gm, rm)):
        np.testing.assert_allclose(got.cpu().numpy(), exp.cpu().numpy(), rtol=2e-4, atol=2e-5)


def test_word_factor_attention_only_some_grads(dev):
    from vlgae_b200.alignment import word_factor_attention

    g = torch.Generator(device=dev).manual_seed(6)
    vis = torch.randn(2, 50, 32, generator=g, device=dev)
    txt = torch.randn(2, 7, 32, generator=g, device=dev).requires_grad_()
    mid = torch.randn(2, 50, 16, generator=g, device=dev)
    out = word_factor_attention(vis, txt, mid)
    (gt,) = torch.autograd.grad(out.sum(), [txt])
    t64 = txt.detach().double().requires_grad_()
    ref = torch.einsum("bqv,bvh->bqh", torch.einsum("bvd,bqd->bqv", vis.double(), t64).softmax(2), mid.double())
    (rt,) = torch.autograd.grad(ref.sum(), [t64])
    np.testing.assert_allclose(gt.cpu().numpy(), rt.cpu().numpy(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("A,V,B,Q,D,pad", [(2, 150, 3, 20, 128, False), (3, 300, 4, 82, 128, True), (2, 129, 2, 130, 64, False),
                                           (2, 40, 2, 5, 8, True), (5, 1369, 6, 82, 128, True)])
def test_backward_kernels_against_fp64(dev, A, V, B, Q, D, pad):
    """vlgae_align_logits_backward (tcgen05, split bf16) against the two transposed contractions in fp64, with the masks
    folded in; `pad` passes the upstream gradient with padded rows (stride > V), as autograd does for the padded output."""
    from vlgae_b200._lib import check, lib

    g_ = torch.Generator(device=dev).manual_seed(A * 7 + Q)
    vis = torch.randn(A, V, D, generator=g_, device=dev)
    txt = torch.randn(B, Q, D, generator=g_, device=dev)
    vm = torch.rand(A, V, generator=g_, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g_, device=dev) > 0.2
    ld = (V + 7) // 8 * 8 if pad else V
    gfull = torch.randn(B, A, Q, ld, generator=g_, device=dev)
    g = gfull[..., :V]
    gv = torch.full((A, V, D), 7.0, device=dev)
    gt = torch.full((B, Q, D), 7.0, device=dev)
    ws = torch.empty(lib().vlgae_align_workspace_bytes(A, V, B, Q, D), dtype=torch.uint8, device=dev)
    vmu, tmu = vm.view(torch.uint8), tm.view(torch.uint8)
    check(lib().vlgae_align_logits_backward(gfull.data_ptr(), ld, vis.data_ptr(), vmu.data_ptr(), txt.data_ptr(), tmu.data_ptr(),
                                            A, V, B, Q, D, 3, gv.data_ptr(), gt.data_ptr(), ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream().cuda_stream), "bwd")
    torch.cuda.synchronize()
    gm = (g * vm[None, :, None, :] * tm[:, None, :, None]).double()
    rv = torch.einsum("baqv,bqd->avd", gm, txt.double())
    rt = torch.einsum("baqv,avd->bqd", gm, vis.double())
    # error model of a split-bf16 product summed over K terms: 2^-16 per product, random-walk accumulation
    tol_v = 2.0 ** -14 * float(g.abs().max() * txt.abs().max()) * (B * Q) ** 0.5
    tol_t = 2.0 ** -14 * float(g.abs().max() * vis.abs().max()) * (A * V) ** 0.5
    assert float((gv.double() - rv).abs().max()) <= tol_v
    assert float((gt.double() - rt).abs().max()) <= tol_t
    assert (gv[~vm] == 0).all() and (gt[~tm] == 0).all()   # masked rows receive exact zeros


@pytest.mark.parametrize("pad", [True, False])
def test_inplace_write_into_attmap_then_backward(dev, pad):
    """The reference's default training config writes into attmap in place (use_pos_prior, joint.py:466-469:
    ``attmap[arange, arange, ...] -= mask * 100``) and then back-propagates through it.  With requires_grad inputs the
    result comes from the autograd node; it must be writable (not a view made inside the custom Function) and the
    gradients must equal those of the reference formula with the same in-place write."""
    from vlgae_b200.alignment import gather_logit_simple

    g_ = torch.Generator(device=dev).manual_seed(21)
    A = B = 3
    V, Q, D = 45, 10, 32   # V = 45 is odd: padded rows (ldv = 48) when pad
    vis = torch.randn(A, V, D, generator=g_, device=dev).requires_grad_()
    txt = torch.randn(B, Q, D, generator=g_, device=dev).requires_grad_()
    vm = torch.rand(A, V, generator=g_, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g_, device=dev) > 0.2
    prior = (torch.rand(B, Q, V, generator=g_, device=dev) > 0.5).float()
    w = torch.randn(B, A, Q, V, generator=g_, device=dev)

    att = gather_logit_simple(vis, vm, txt, tm, named=False, pad_rows=pad)
    ar = torch.arange(B, device=dev)
    att[ar, ar] -= prior * 100          # must not raise
    keep = vm[None, :, None, :] & tm[:, None, :, None]
    (att * w * keep).sum().backward()
    gv, gt = vis.grad.clone(), txt.grad.clone()

    vis2, txt2 = vis.detach().double().requires_grad_(), txt.detach().double().requires_grad_()
    ref = torch.einsum("avd,bqd->baqv", vis2, txt2)
    ref = ref.masked_fill(~vm[None, :, None, :], -1e20).masked_fill(~tm[:, None, :, None], -1e20)
    ref[ar, ar] -= prior.double() * 100
    (ref * w.double() * keep).sum().backward()
    tol = 2.0 ** -13 * float(w.abs().max()) * float(max(vis.abs().max(), txt.abs().max())) * (max(A * V, B * Q)) ** 0.5
    assert float((gv.double() - vis2.grad).abs().max()) <= tol
    assert float((gt.double() - txt2.grad).abs().max()) <= tol
    got, want = att.detach().cpu().numpy(), ref.detach().cpu().numpy()
    masked = ~keep.expand(B, A, Q, V).cpu().numpy()
    assert np.abs(got - want)[~masked].max() < 1e-2 and (got[masked] <= -1e19).all()


# ---- fused grounding consumers ------------------------------------------------------------------
def _loss_inputs(g, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    off = np.concatenate([[0], np.cumsum(g["vis_split"])])
    prior = [(t(g["pos_obj"]), int(off[0]), int(off[1])), (t(g["pos_rel"]), int(off[1]), int(off[2])),
             (t(g["pos_attr"]), int(off[2]), int(off[3]))]
    return t(g["vis_feat"]), t(g["vis_mask"]), t(g["txt_feat"]), t(g["txt_mask"]), t(g["txt_marginal"]), prior


@pytest.mark.parametrize("A,V,B,Q,D", [(3, 150, 4, 20, 64), (5, 1369, 6, 82, 128), (2, 129, 2, 128, 32), (4, 40, 3, 5, 8)])
def test_fused_maxima_bit_identical(dev, A, V, B, Q, D):
    """Both maxima of one pass (vlgae_align_maxima) equal torch.max of the materialised logits bit for bit, arg-max =
    smallest attaining index, incl. fully masked rows / columns."""
    from vlgae_b200.alignment import fused_maxima, gather_logit_simple

    g_ = torch.Generator(device=dev).manual_seed(V + Q)
    vis = torch.randn(A, V, D, generator=g_, device=dev)
    txt = torch.randn(B, Q, D, generator=g_, device=dev)
    vm = torch.rand(A, V, generator=g_, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g_, device=dev) > 0.2
    vm[0, :] = False       # an image with every factor masked
    tm[B - 1, :] = False   # a caption with every query masked
    att = gather_logit_simple(vis, vm, txt, tm, named=False, pad_rows=False)
    maxv, argv, maxq, argq = fused_maxima(vis, vm, txt, tm)
    torch.cuda.synchronize()
    rv, rq = att.max(-1), att.max(2)
    assert torch.equal(maxv, rv.values) and torch.equal(maxq, rq.values)
    # first maximal index (torch.max on CUDA may return any index on exact ties; compare through the values it points to)
    assert torch.equal(att.gather(-1, argv.long().unsqueeze(-1)).squeeze(-1), rv.values)
    assert torch.equal(att.gather(2, argq.long().unsqueeze(2)).squeeze(2), rq.values)
    first_v = (att == rv.values.unsqueeze(-1)).float().argmax(-1)
    first_q = (att == rq.values.unsqueeze(2)).float().argmax(2)
    assert torch.equal(argv.long(), first_v) and torch.equal(argq.long(), first_q)


def test_diagonal_slab_and_topk(golden, dev):
    from vlgae_b200.alignment import diagonal_slab, topk_rows

    g = golden("align_loss")
    vis, vm, txt, tm, _, _ = _loss_inputs(g, dev)
    diag = diagonal_slab(vis, vm, txt, tm)
    got, want = diag.cpu().numpy(), g["diag"]
    masked = want <= -1e19
    assert ((got <= -1e19) == masked).all()
    np.testing.assert_allclose(got[~masked], want[~masked], rtol=1e-5, atol=1e-5)
    # top 5 of the decode logits (joint.py:594) on the reference's own tensor: same sets, same order where values differ
    dl = torch.from_numpy(g["decode_logit"]).to(dev)
    top = topk_rows(dl, 5).cpu().numpy()
    vals = np.take_along_axis(g["decode_logit"], top, -1)
    want_vals = np.take_along_axis(g["decode_logit"], g["top5"], -1)
    np.testing.assert_array_equal(vals, want_vals)
    assert (np.diff(vals, axis=-1) <= 0).all()
    x = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0, 0.0]], device=dev)
    assert topk_rows(x, 5).tolist() == [[1, 2, 4, 3, 0]]   # ties: smaller index first


@pytest.mark.parametrize("use_prior", [False, True])
def test_fused_grounding_loss_matches_reference(golden, dev, use_prior):
    """loss_grounding_factor_ce through the fused consumers (no [B, A, Q, V] tensor) against the reference's own lines."""
    from vlgae_b200.alignment import FusedMatch, grounding_loss_fused

    g = golden("align_loss")
    vis, vm, txt, tm, marg, prior = _loss_inputs(g, dev)
    match = FusedMatch(vis, vm, txt, tm)
    total, loss, (t2v, v2t) = grounding_loss_fused(match, marg, vm, 37, prior=prior if use_prior else None, vis2txt=1.0)
    k = "prior" if use_prior else "plain"
    np.testing.assert_allclose(float(t2v), float(g["txt2vis_" + k]), rtol=2e-4)
    np.testing.assert_allclose(float(v2t), float(g["vis2txt_" + k]), rtol=2e-4)
    np.testing.assert_allclose(float(total), 2 * 37, rtol=1e-5)   # each term is self-normalised to num_token (:477-483)
    # decode accesses of the UNMODIFIED reference method (joint.py:520-524)
    factor2img = match.max("V").values.max("A").indices
    np.testing.assert_array_equal(factor2img.rename(None).cpu().numpy(), g["factor2img"])
    d = match.diagonal().refine_names("Q", "V", "B").align_to("B", "Q", "V").rename(None)
    assert d.shape == g["diag"].shape


def test_max_over_factors_backward_kernel(dev):
    """gather_logit_reduced's backward (kernel, vlgae_align_max_over_factors_backward) against autograd through the
    reference formula in fp64."""
    from vlgae_b200.alignment import gather_logit_reduced

    g_ = torch.Generator(device=dev).manual_seed(3)
    A, V, B, Q, D = 4, 70, 5, 12, 48
    vis = torch.randn(A, V, D, generator=g_, device=dev).requires_grad_()
    txt = torch.randn(B, Q, D, generator=g_, device=dev).requires_grad_()
    vm = torch.rand(A, V, generator=g_, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g_, device=dev) > 0.2
    marg = torch.rand(B, Q, generator=g_, device=dev) * tm
    w = torch.randn(B, A, generator=g_, device=dev)
    (gather_logit_reduced(vis, vm, txt, tm, marg) * w).sum().backward()
    vis2, txt2 = vis.detach().double().requires_grad_(), txt.detach().double().requires_grad_()
    att = torch.einsum("avd,bqd->baqv", vis2, txt2)
    att = att.masked_fill(~vm[None, :, None, :], -1e20).masked_fill(~tm[:, None, :, None], -1e20)
    ref = (att.max(-1).values * marg.double().unsqueeze(1)).sum(-1) / marg.double().sum(1, keepdim=True)
    (ref * w.double()).sum().backward()
    assert float((vis.grad.double() - vis2.grad).abs().max()) < 1e-4
    assert float((txt.grad.double() - txt2.grad).abs().max()) < 1e-4


@pytest.mark.parametrize("V", [121, 127, 129, 135, 241, 247, 249, 255, 361, 369, 1369])
def test_dense_layout_sector_tiles_equal_padded(dev, V):
    """The reference's dense layout (rows of odd length) is written by overlapping tiles 120 factors apart, every tile storing
    whole 32-byte sectors of a row (align_gemm_kernel VSTEP): bit-identical to the padded result at every tile-boundary case."""
    from vlgae_b200.alignment import gather_logit_simple

    g_ = torch.Generator(device=dev).manual_seed(V)
    A, B, Q, D = 3, 5, 19, 64
    vis = torch.randn(A, V, D, generator=g_, device=dev)
    txt = torch.randn(B, Q, D, generator=g_, device=dev)
    vm = torch.rand(A, V, generator=g_, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g_, device=dev) > 0.2
    dense = gather_logit_simple(vis, vm, txt, tm, named=False, pad_rows=False)
    padded = gather_logit_simple(vis, vm, txt, tm, named=False, pad_rows=True)
    torch.cuda.synchronize()
    assert dense.is_contiguous() and dense.shape == (B, A, Q, V)
    assert torch.equal(dense, padded)
    ref = torch.einsum("avd,bqd->baqv", vis.double(), txt.double())
    keep = (vm[None, :, None, :] & tm[:, None, :, None]).expand(B, A, Q, V)
    assert float((dense.double() - ref)[keep].abs().max()) < 2e-3 and bool((dense[~keep] == -1e20).all())
