"""tests/golden/gen_golden_scores.py -- golden vectors for the score-tensor construction.

Run in the build container only (reads /root/reference):

    python tests/golden/gen_golden_scores.py

``src/model/ldndmv.py`` cannot be imported (hydra, fastNLP, lightning), so lines 184-209 of its ``_forward`` are
re-typed below with the same torch calls, on top of the reference's OWN modules, imported unmodified by file path:
``DMVFactorizedBilinear`` (/root/reference/src/model/nn/dmv_spec.py:59-76, needs only torch) for the three scorers and
``DMV1o.merge`` from the reference's torch_struct package.  The projected operands (outputs of ``project1`` /
``project2``, the seam of vlgae_dmv_scores) are captured with forward hooks, so the scorer's forward runs as is; the
gradients w.r.t. them come from the reference's autograd under a fixed random cotangent of the merged tensors.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import load_reference  # noqa: E402

INF = 1e20  # /root/reference/src/__init__.py:110
LEFT, RIGHT = 0, 1


def load_spec():
    spec = importlib.util.spec_from_file_location("ref_dmv_spec", "/root/reference/src/model/nn/dmv_spec.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def run(ref, spec_mod, name, B, n, T, hid, r, seed, function_mask):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    attach_scorer = spec_mod.DMVFactorizedBilinear(n_in=hid, r=r)
    dec_scorer = spec_mod.DMVFactorizedBilinear(n_in=hid, r=r)
    root_scorer = spec_mod.DMVFactorizedBilinear(n_in=hid, r=r)
    h_parent = torch.randn(B, n, 2, 2, hid, generator=g)
    h_child = torch.randn(1, T, 2, 2, hid, generator=g)
    h_root = torch.randn(1, 1, 2, 2, hid, generator=g)
    h_dec = torch.randn(1, 2, 2, 2, hid, generator=g)
    token = torch.randint(0, T, (B, n), generator=g)
    tag = torch.randint(0, 6, (B, n), generator=g)
    fmask = torch.tensor([1, 4])
    captured = {}

    def capture(key):
        def hook(_mod, _inp, out):
            out.retain_grad()
            captured[key] = out  # (returns None: the output is passed on unchanged)
        return hook

    hooks = [attach_scorer.project1.register_forward_hook(capture("x1")),
             attach_scorer.project2.register_forward_hook(capture("x2"))]
    b = B
    # ---- ldndmv.py:184-209, re-typed ----
    attach_rule = attach_scorer(h_parent, h_child).log_softmax(2)
    target_size = torch.Size([b, n, n, 2, 2])
    attach_prob = attach_rule.gather(2, token.reshape(b, 1, n, 1, 1).expand(target_size))
    left_mask = torch.tril(torch.ones(n, n), diagonal=-1)
    right_mask = torch.triu(torch.ones(n, n), diagonal=1)
    attach_prob = attach_prob[..., LEFT, :] * left_mask.unsqueeze(0).unsqueeze(-1) \
        + attach_prob[..., RIGHT, :] * right_mask.unsqueeze(0).unsqueeze(-1)
    in_mask = None
    if function_mask:
        tag_array = tag.unsqueeze(-1).unsqueeze(-1)
        mask_set = fmask.view(1, 1, 1, -1)
        in_mask = tag_array.eq(mask_set).any(dim=-1, keepdims=True)
        attach_prob.masked_fill_(in_mask, -INF)
    dec_raw = dec_scorer(h_parent, h_dec)
    dec_raw.retain_grad()
    dec_prob = dec_raw.permute(0, 1, 3, 4, 2).log_softmax(-1)
    root_raw = root_scorer(h_root, h_child).sum([-1, -2])
    root_raw.retain_grad()
    root_prob = root_raw.log_softmax(-1).squeeze(1).expand(b, -1)
    root = torch.gather(root_prob, 1, token)
    merged_dec, merged_attach = ref.DMV1o.merge(dec_prob, attach_prob, root)
    # ---- gradients under a fixed cotangent ----
    gmd = torch.randn(merged_dec.shape, generator=g)
    gma = torch.randn(merged_attach.shape, generator=g)
    ((merged_dec * gmd).sum() + (merged_attach * gma).sum()).backward()
    for h in hooks:
        h.remove()
    x1, x2 = captured["x1"], captured["x2"]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), x1=x1.detach().numpy(), x2=x2.detach().numpy()[0], token=token.numpy(),
        head_mask=(in_mask.reshape(B, n).numpy() if in_mask is not None else np.zeros((B, n), dtype=bool)),
        function_mask=np.array(function_mask), dec_score=dec_raw.detach().numpy(), root_score=root_raw.detach().numpy().reshape(T),
        attach=attach_prob.detach().numpy(), dec=dec_prob.detach().numpy(), root=root.detach().numpy(),
        merged_dec=merged_dec.detach().numpy(), merged_attach=merged_attach.detach().numpy(), grad_merged_dec=gmd.numpy(),
        grad_merged_attach=gma.numpy(), grad_x1=x1.grad.numpy(), grad_x2=x2.grad.numpy()[0], grad_dec_score=dec_raw.grad.numpy(),
        grad_root_score=root_raw.grad.numpy().reshape(T))
    print(name, "B", B, "n", n, "T", T, "r", r, "attach[0,0,1]", attach_prob[0, 0, 1].tolist())


def main():
    torch.set_num_threads(1)
    ref = load_reference()
    spec_mod = load_spec()
    run(ref, spec_mod, "scores_small", B=3, n=7, T=50, hid=12, r=16, seed=21, function_mask=False)
    run(ref, spec_mod, "scores_fmask", B=4, n=9, T=133, hid=10, r=8, seed=22, function_mask=True)
    run(ref, spec_mod, "scores_mid", B=6, n=20, T=700, hid=16, r=16, seed=23, function_mask=False)


if __name__ == "__main__":
    main()
