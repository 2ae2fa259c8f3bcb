"""tests/golden/gen_golden.py -- generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/gen_golden.py

It imports the reference's own ``src/model/torch_struct`` package in isolation (the package needs
only torch; ``import src`` itself needs hydra/lightning/fastNLP, which are absent), emulates
``src.setup_inf(1e20)`` (/root/reference/src/__init__.py:113-120) and records, for seeded synthetic
inputs built exactly as SURVEY.md section 8(d) prescribes:

  * ``DMV1o.merge`` outputs, ``.partition``, ``.max``, ``autograd.grad(partition.sum(), [dec, attach])``,
    ``autograd.grad(max.sum(), [dec, attach])`` and ``.argmax`` (stored compactly as heads + arc valence);
  * ``DependencyCRF`` partition / max / marginals / argmax, incl. the MBR chain used by
    /root/reference/src/model/ldndmv.py:294-299;
  * the alignment operator ``gather_logit_simple`` / ``gather_logit_reduced``.  ``src/model/joint.py``
    cannot be imported here (fastNLP, hydra), so those two methods are re-typed below line for line
    from /root/reference/src/model/joint.py:406-432 using the same torch calls (einsum, named tensors,
    masked_fill_); they run on torch's CPU kernels, which is the arithmetic the reference uses.

Nothing from the reference is copied into the repository: only inputs and outputs are stored.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src/model/torch_struct"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_torch_struct", REF + "/__init__.py",
                                                  submodule_search_locations=[REF])
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_torch_struct"] = m
    spec.loader.exec_module(m)
    m.semirings.semirings.NEGINF = -1e20  # src.setup_inf(1e20)
    return m


def synth(B, n, seed, quant=None, zero=False):
    g = torch.Generator().manual_seed(seed)
    dec = torch.randn(B, n, 2, 2, 2, generator=g).log_softmax(-1)
    attach = torch.randn(B, n, n, 2, generator=g).log_softmax(2)
    root = torch.randn(B, n, generator=g).log_softmax(-1)
    if quant is not None:
        dec, attach, root = [(t / quant).round() * quant for t in (dec, attach, root)]
    if zero:
        dec, attach, root = [torch.zeros_like(t) for t in (dec, attach, root)]
    return dec, attach, root


def run_dmv(ref, name, dec, attach, root, lengths, compact=False):
    lengths = torch.as_tensor(lengths, dtype=torch.long)
    md, ma = ref.DMV1o.merge(dec, attach, root)
    a = ma.detach().clone().requires_grad_()
    d = md.detach().clone().requires_grad_()
    dist = ref.DMV1o([d, a], lengths)
    Z = dist.partition
    gd, ga = torch.autograd.grad(Z.sum(), [d, a])
    dist2 = ref.DMV1o([d, a], lengths)
    mx = dist2.max
    vgd, vga = torch.autograd.grad(mx.sum(), [d, a])
    dist3 = ref.DMV1o([d, a], lengths)
    arg = dist3.argmax.detach()  # [B, N, N, 2]
    assert torch.equal(arg, vga), "argmax and grad of max disagree in the reference"
    B, N = md.shape[:2]
    heads = np.zeros((B, N), dtype=np.int64)
    val = np.full((B, N), -1, dtype=np.int8)
    for b, h, c, v in arg.nonzero().tolist():
        heads[b, c] = h
        val[b, c] = v
    # compact: the big fixtures store the raw inputs only (merge is pinned bit-exactly by the small ones)
    inputs = dict(dec=dec.numpy(), attach=attach.numpy(), root=root.numpy())
    if not compact:
        inputs.update(merged_dec=md.numpy(), merged_attach=ma.numpy())
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), **inputs,
        lengths=lengths.numpy(), partition=Z.detach().numpy(),
        max=mx.detach().numpy(), grad_dec=gd.numpy(), grad_attach=ga.numpy(), vgrad_dec=vgd.numpy(), heads=heads,
        arc_valence=val)
    print(name, "B", B, "N", N, "Z[0]", float(Z[0]), "max[0]", float(mx[0]))
    return ga


def run_deptree(ref, name, arc, lengths):
    lengths = torch.as_tensor(lengths, dtype=torch.long)
    a = arc.detach().clone().requires_grad_()
    dist = ref.DependencyCRF(a, lengths)
    Z = dist.partition
    marg = torch.autograd.grad(Z.sum(), a)[0]
    mx = ref.DependencyCRF(a, lengths).max
    arg = ref.DependencyCRF(a, lengths).argmax.detach()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), arc=arc.numpy(), lengths=lengths.numpy(),
                        partition=Z.detach().numpy(), max=mx.detach().numpy(), marginals=marg.numpy(),
                        argmax=arg.numpy().astype(np.int8))
    print(name, tuple(arc.shape), "Z", Z.detach().numpy()[:3])


INF = 1e20  # /root/reference/src/__init__.py:110


def gather_logit_simple(vis, txt):
    # re-typed from /root/reference/src/model/joint.py:406-419
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, txt_marginal = txt
    attmap = torch.einsum("avd, bqd -> baqv", vis_feat.rename(None), txt_feat.rename(None))
    attmap = attmap.refine_names("B", "A", "Q", "V")
    attmap.masked_fill_(~vis_mask.align_as(attmap), -INF)
    attmap.masked_fill_(~txt_mask.align_as(attmap), -INF)
    return attmap


def gather_logit_reduced(vis, txt):
    # re-typed from /root/reference/src/model/joint.py:421-432
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, txt_marginal = txt
    attmap = gather_logit_simple(vis, txt)
    maxatt = attmap.max(dim=-1).values
    logit = torch.sum(maxatt * txt_marginal.unsqueeze(1), dim=-1) / txt_marginal.sum(1, keepdim=True)
    return logit


def run_align(name, A, V, B, Q, D, seed):
    g = torch.Generator().manual_seed(seed)
    vis_feat = torch.randn(A, V, D, generator=g)
    txt_feat = torch.randn(B, Q, D, generator=g)
    vis_mask = torch.rand(A, V, generator=g) > 0.2
    vis_mask[:, 0] = True
    half = Q // 2
    lens = torch.randint(1, half, (B,), generator=g)
    m = torch.arange(half)[None, :] <= lens[:, None]
    m[:, 0] = False  # ROOT slot is masked (joint.py:248-249)
    txt_mask = torch.cat([m, m], dim=1)
    txt_marginal = torch.rand(B, Q, generator=g) * txt_mask
    vis = (vis_feat.refine_names("A", "V", "D"), vis_mask.refine_names("A", "V"), None)
    txt = (txt_feat.refine_names("B", "Q", "D"), txt_mask.refine_names("B", "Q"), txt_marginal)
    att = gather_logit_simple(vis, txt)
    red = gather_logit_reduced(vis, (txt[0], txt[1], txt_marginal))
    att_u = att.rename(None)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), vis_feat=vis_feat.numpy(), vis_mask=vis_mask.numpy(),
        txt_feat=txt_feat.numpy(), txt_mask=txt_mask.numpy(), txt_marginal=txt_marginal.numpy(),
        attmap=att_u.numpy(), reduced=red.rename(None).numpy(), max_v=att_u.max(-1).values.numpy(),
        argmax_v=att_u.max(-1).indices.numpy(), max_q=att_u.max(2).values.numpy())
    print(name, tuple(att.shape))


def run_align_loss(name, B, nb, T, D, seed):
    """loss_grounding_factor_ce (joint.py:439-491) and the head of decode_grounding_on_factor (:518-551, :594), re-typed
    line for line on attmap from gather_logit_simple above; A = B, V = nb + nb^2 + nb + 1 (obj, rel, attr, img),
    Q = 2 (T + 1) (words incl. ROOT, then arcs)."""
    g = torch.Generator().manual_seed(seed)
    vis_split = [nb, nb * nb, nb, 1]
    names_f = ["obj", "rel", "attr", "img"]
    V, Q = sum(vis_split), 2 * (T + 1)
    vis_feat = torch.randn(B, V, D, generator=g) * 0.5
    txt_feat = torch.randn(B, Q, D, generator=g) * 0.5
    vis_mask = torch.rand(B, V, generator=g) > 0.15
    vis_mask[:, 0] = True
    lens = torch.randint(2, T + 1, (B,), generator=g)
    m = torch.arange(T + 1)[None, :] <= lens[:, None]
    m[:, 0] = False
    txt_mask = torch.cat([m, m], dim=1)
    txt_marginal = torch.rand(B, Q, generator=g) * txt_mask
    pos_mask = {n: torch.rand(B, T, 1, generator=g) > 0.6 for n in ("obj", "rel", "attr")}
    vis = (vis_feat.refine_names("A", "V", "D"), vis_mask.refine_names("A", "V"), None)
    txt = (txt_feat.refine_names("B", "Q", "D"), txt_mask.refine_names("B", "Q"), txt_marginal)
    out = {}
    for use_prior in (False, True):
        attmap = gather_logit_simple(vis, txt)
        # ---- joint.py:446-470
        if use_prior:
            names = attmap.names
            attmap = attmap.rename(None)
            arange = torch.arange(len(attmap))
            offset = 0
            for i, (fname, width) in enumerate(zip(names_f, vis_split)):
                if fname in pos_mask:
                    mask = pos_mask[fname]
                else:
                    offset += width
                    continue
                attmap[arange, arange, 1: mask.shape[1] + 1, :offset] -= mask * 100
                attmap[arange, arange, 1: mask.shape[1] + 1, offset + width:] -= (mask * 100)
                offset += width
            attmap = attmap.refine_names(*names)
        # ---- joint.py:473-483
        logit = attmap.max("V").values
        logit = logit.log_softmax("A")
        txt2vis = -(logit.rename(None).diagonal().T * txt_marginal).sum()
        logit = attmap.max("Q").values
        logit = logit.log_softmax("B")
        vis2txt = -(logit.rename(None).diagonal().T * vis_mask).sum()
        k = "prior" if use_prior else "plain"
        out["txt2vis_" + k], out["vis2txt_" + k] = txt2vis.numpy(), vis2txt.numpy()
    # ---- decode head, joint.py:518-551 (+ :594)
    match_logit = gather_logit_simple(vis, txt)
    factor2img = match_logit.max("V").values.max("A").indices
    match_logit = match_logit.diagonal().refine_names("Q", "V", "B").align_to("B", "Q", "V").rename(None)
    out["diag"] = match_logit.clone().numpy()
    arange = torch.arange(len(match_logit))
    offset = 0
    for i, (fname, width) in enumerate(zip(names_f, vis_split)):
        if fname in pos_mask:
            mask = pos_mask[fname]
        else:
            offset += width
            continue
        match_logit[arange, 1: mask.shape[1] + 1, :offset] -= 1e10 * mask
        match_logit[arange, 1: mask.shape[1] + 1, offset + width:] -= (1e10 * mask)
        offset += width
    top5 = match_logit.argsort(-1, descending=True)[..., :5]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), vis_feat=vis_feat.numpy(), vis_mask=vis_mask.numpy(), txt_feat=txt_feat.numpy(),
        txt_mask=txt_mask.numpy(), txt_marginal=txt_marginal.numpy(), vis_split=np.array(vis_split),
        pos_obj=pos_mask["obj"].numpy(), pos_rel=pos_mask["rel"].numpy(), pos_attr=pos_mask["attr"].numpy(),
        factor2img=factor2img.rename(None).numpy(), decode_logit=match_logit.numpy(), top5=top5.numpy(), **out)
    print(name, "B", B, "V", V, "Q", Q, {k: float(v) for k, v in out.items() if k != "diag"})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref = load_reference()
    # tiny ragged batch incl. length 1 and 2
    d, a, r = synth(8, 6, 11)
    run_dmv(ref, "dmv_tiny_ragged", d, a, r, [6, 5, 4, 3, 2, 1, 6, 2])
    # BASELINE cfg1: B=64, n=16, full length, seed 1
    d, a, r = synth(64, 16, 1)
    run_dmv(ref, "dmv_cfg1", d, a, r, [16] * 64)
    # cfg1 ragged variant
    g = torch.Generator().manual_seed(101)
    L = torch.randint(1, 17, (16,), generator=g)
    L[0] = 16
    d, a, r = synth(16, 16, 12)
    run_dmv(ref, "dmv_cfg1_ragged", d, a, r, L)
    # tie stress: scores quantised to multiples of 0.25 / 1.0, and the all-zero KAT
    g = torch.Generator().manual_seed(102)
    L = torch.randint(1, 11, (32,), generator=g)
    L[0] = 10
    d, a, r = synth(32, 10, 13, quant=0.25)
    run_dmv(ref, "dmv_ties_q025", d, a, r, L)
    d, a, r = synth(32, 10, 14, quant=1.0)
    run_dmv(ref, "dmv_ties_q1", d, a, r, L)
    d, a, r = synth(4, 7, 15, zero=True)
    run_dmv(ref, "dmv_ties_zero", d, a, r, [7, 5, 3, 1])
    # long sentences (cfg2 upper end)
    d, a, r = synth(4, 40, 2)
    ga = run_dmv(ref, "dmv_len40", d, a, r, [40, 33, 17, 4])
    # the full cfg2 batch bench.py times (BASELINE.json configs[1]: 128 captions, len 4..40 ragged sorted desc, seed 2)
    # -- same draws as bench.py:make_batch_cpu / make_lengths
    g = torch.Generator().manual_seed(2)
    L = torch.randint(4, 41, (128,), generator=g).sort(descending=True).values
    L[0] = 40
    d, a, r = synth(128, 40, 2)
    run_dmv(ref, "dmv_cfg2_full", d, a, r, L, compact=True)
    # cfg3 upper end (BASELINE.json configs[2]) at small B: n = 64 and n = 128, full length
    d, a, r = synth(6, 64, 3)
    run_dmv(ref, "dmv_n64", d, a, r, [64] * 6, compact=True)
    d, a, r = synth(3, 128, 3)
    run_dmv(ref, "dmv_n128", d, a, r, [128] * 3, compact=True)
    # DependencyCRF: random potentials, and the MBR chain on real arc marginals
    g = torch.Generator().manual_seed(103)
    run_deptree(ref, "deptree_rand", torch.randn(8, 9, 9, generator=g), [8, 7, 5, 3, 2, 1, 8, 4])
    run_deptree(ref, "deptree_mbr", ga.sum(-1), [40, 33, 17, 4])
    run_deptree(ref, "deptree_ties", (torch.randn(8, 9, 9, generator=g)).round(), [8, 7, 5, 3, 2, 1, 8, 4])
    # alignment
    run_align("align_small", A=3, V=11, B=4, Q=10, D=16, seed=21)
    run_align("align_mid", A=5, V=150, B=6, Q=18, D=128, seed=22)
    run_align_loss("align_loss", B=6, nb=7, T=9, D=64, seed=23)


if __name__ == "__main__":
    main()
