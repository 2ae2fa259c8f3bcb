"""tests/golden/gen_golden_vis.py -- golden vectors for the visual factor features.

Run in the build container only (reads /root/reference):   python tests/golden/gen_golden_vis.py

The reference's ``MLP`` (/root/reference/src/model/nn/common.py:23-51) is imported unmodified by file path (its one
``src.`` import, ``src.model.nn.dropout``, is satisfied by loading that file -- torch only -- under the same module name);
``VisBoxRelSimpleEncoder`` itself and ``joint.py`` need hydra / fastNLP, so box_rel.py:31-54 and joint.py:140-179 are
re-typed below with the same torch calls on top of those MLPs.  Stored: inputs, the MLP weights, the reference's outputs
(``vis``, ``vis_mask``, ``_mid``) and its autograd gradients w.r.t. the box features and the weights under a fixed cotangent.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_NN = "/root/reference/src/model/nn"


def load_mlp():
    for name in ("src", "src.model", "src.model.nn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    for name, path in (("src.model.nn.dropout", REF_NN + "/dropout.py"), ("ref_nn_common", REF_NN + "/common.py")):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
    return sys.modules["ref_nn_common"].MLP


def run(MLP, name, B, n, F, H, Y, seed, use_attr, add_image, img_feat):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    n_in = F * 2 if img_feat else F
    box_fc, rel_fc = MLP(n_in, H, 0.0, True), MLP(n_in, H, 0.0, True)
    attr_fc = MLP(n_in, H, 0.0, True) if use_attr else None
    pre = nn.Linear(H, Y, bias=False)
    feat = torch.randn(B, n, F, generator=g).requires_grad_()
    _box_mask = torch.rand(B, n, generator=g) > 0.25
    _box_mask[:, 0] = True
    # ---- box_rel.py:31-54, re-typed ----
    if img_feat:
        inputs = torch.cat([feat, feat.mean(1, keepdim=True).expand(-1, feat.shape[1], -1)], dim=-1)
    else:
        inputs = feat
    _rel_inp = (inputs.unsqueeze(1) + inputs.unsqueeze(2)) / 2
    x_rel = rel_fc(_rel_inp)
    rel = x_rel.view(len(x_rel), -1, H)
    encoded = {"box": box_fc(inputs), "rel": rel}
    if use_attr:
        encoded["attr"] = attr_fc(inputs)
    # ---- joint.py:140-179, re-typed (inputs["vis_rel_mask"] present, cfg.add_rel) ----
    feats = [encoded["box"]]
    mask = [_box_mask]
    split = [_box_mask.shape[1]]
    feats.append(encoded["rel"])
    rel_mask = _box_mask.unsqueeze(1) * _box_mask.unsqueeze(2)
    rel_mask = rel_mask.triu(1)
    rel_mask = rel_mask.view(B, -1)
    mask.append(rel_mask)
    split.append(_box_mask.shape[1] * _box_mask.shape[1])
    if use_attr:
        feats.append(encoded["attr"]); mask.append(_box_mask); split.append(_box_mask.shape[1])
    if add_image:
        feats.append(encoded["box"].mean(1, keepdim=True))
        mask.append(torch.ones(len(encoded["box"]), 1, dtype=torch.bool))
        split.append(1)
    vis = _mid = torch.cat(feats, dim=1)
    vis = pre(vis)
    vis_mask = torch.cat(mask, dim=1)
    # ---- gradients under a fixed cotangent of vis and _mid (the word attention reads _mid) ----
    gv, gm = torch.randn(vis.shape, generator=g), torch.randn(_mid.shape, generator=g) * 0.1
    params = [feat, box_fc.linear.weight, box_fc.linear.bias, rel_fc.linear.weight, rel_fc.linear.bias, pre.weight]
    if use_attr:
        params += [attr_fc.linear.weight, attr_fc.linear.bias]
    grads = torch.autograd.grad([vis, _mid], params, [gv, gm])
    out = dict(feat=feat.detach().numpy(), box_mask=_box_mask.numpy(), w_box=box_fc.linear.weight.detach().numpy(),
               b_box=box_fc.linear.bias.detach().numpy(), w_rel=rel_fc.linear.weight.detach().numpy(),
               b_rel=rel_fc.linear.bias.detach().numpy(), w_pre=pre.weight.detach().numpy(), vis=vis.detach().numpy(),
               mid=_mid.detach().numpy(), vis_mask=vis_mask.numpy(), split=np.array(split), grad_vis=gv.numpy(), grad_mid=gm.numpy(),
               g_feat=grads[0].numpy(), g_w_box=grads[1].numpy(), g_b_box=grads[2].numpy(), g_w_rel=grads[3].numpy(),
               g_b_rel=grads[4].numpy(), g_w_pre=grads[5].numpy(), use_attr=np.array(use_attr), add_image=np.array(add_image),
               img_feat=np.array(img_feat))
    if use_attr:
        out.update(w_attr=attr_fc.linear.weight.detach().numpy(), b_attr=attr_fc.linear.bias.detach().numpy(),
                   g_w_attr=grads[6].numpy(), g_b_attr=grads[7].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "vis", tuple(vis.shape), "split", split)


def main():
    torch.set_num_threads(1)
    MLP = load_mlp()
    run(MLP, "vis_small", B=3, n=5, F=24, H=16, Y=8, seed=31, use_attr=True, add_image=True, img_feat=True)
    run(MLP, "vis_noattr", B=2, n=7, F=20, H=12, Y=8, seed=32, use_attr=False, add_image=False, img_feat=False)


if __name__ == "__main__":
    main()
