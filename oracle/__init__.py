"""oracle -- CPU checker for the VLGAE structured-inference hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; nothing under ``vlgae_b200/`` does (the product path fails loudly
when its CUDA library is missing instead of falling back to this).

It wraps ``oracle/dmv_oracle.c`` (a plain-C restatement of the reference's
chart DP, see that file's header for the file:line map) and restates in numpy
the small tensor glue around it:

* ``merge``                 -> /root/reference/src/model/torch_struct/distributions.py:253-265
* ``gather_logit_simple``   -> /root/reference/src/model/joint.py:406-419
* ``gather_logit_reduced``  -> /root/reference/src/model/joint.py:421-432

Parity pin: the reference has no tests or golden vectors for this path, so
the restatement is pinned against outputs of the reference itself (imported
in the build container by ``tests/golden/gen_golden.py``; fixtures committed
under ``tests/golden/``) and against ``oracle/bruteforce.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_dmv.so")

# /root/reference/src/model/torch_struct/dmv.py:7-15
NOCHILD, HASCHILD, LEFT, RIGHT, GO, STOP = 1, 0, 0, 1, 0, 1
# /root/reference/src/model/torch_struct/semirings/semirings.py:16 (import-time value)
NEGINF = -1e12
# /root/reference/src/__init__.py:110
INF = 1e20


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "dmv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


_REF_DIR = os.path.join(_HERE, "_ref", "torch_struct")
_ref_mod = None


def load_reference():
    """The UNMODIFIED reference package (src/model/torch_struct), imported in isolation from ``oracle/_ref`` (staged by
    ``oracle/make_ref.py``) with ``src.setup_inf(1e20)`` emulated (/root/reference/src/__init__.py:113-120).
    Returns the module, or None when the package has not been staged."""
    global _ref_mod
    if _ref_mod is None:
        init = os.path.join(_REF_DIR, "__init__.py")
        if not os.path.exists(init):
            return None
        import importlib.util
        import sys

        spec = importlib.util.spec_from_file_location("ref_torch_struct", init, submodule_search_locations=[_REF_DIR])
        m = importlib.util.module_from_spec(spec)
        sys.modules["ref_torch_struct"] = m
        spec.loader.exec_module(m)
        m.semirings.semirings.NEGINF = -1e20
        _ref_mod = m
    return _ref_mod


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def merge(dec, attach, root, one=0.0, zero=NEGINF):
    """Prepend ROOT as position 0 (distributions.py:253-265).  Always float32."""
    dec, attach, root = _f32(dec), _f32(attach), _f32(root)
    B, n = dec.shape[:2]
    N = n + 1
    attach_w = np.full((B, N, N, 2), zero, dtype=np.float32)
    dec_w = np.full((B, N, 2, 2, 2), zero, dtype=np.float32)
    attach_w[:, 0, 1:, NOCHILD] = root
    attach_w[:, 1:, 1:, :] = attach
    dec_w[:, 0, RIGHT, :, :] = one
    dec_w[:, 1:] = dec
    return dec_w, attach_w


def dmv_log(dec, attach, lengths, *, mask_zero=NEGINF, trim=False, gZ=None, want_grad=True, f64=False):
    """Inside pass in the log semiring + explicit reverse pass.

    Returns ``(Z [B], gdec [B,N,2,2,2] | None, gattach [B,N,N,2] | None)``;
    ``gattach`` is what ``DMV1o.marginals`` / ``autograd.grad(partition.sum(), attach)`` return.
    """
    dec, attach = _f32(dec), _f32(attach)
    lengths = np.ascontiguousarray(np.asarray(lengths, dtype=np.int64))
    B, N = dec.shape[:2]
    assert dec.shape == (B, N, 2, 2, 2) and attach.shape == (B, N, N, 2) and lengths.shape == (B,)
    rt, ct = (np.float64, ctypes.c_double) if f64 else (np.float32, ctypes.c_float)
    Z = np.zeros(B, dtype=rt)
    gdec = np.zeros((B, N, 2, 2, 2), dtype=rt) if want_grad else None
    gatt = np.zeros((B, N, N, 2), dtype=rt) if want_grad else None
    gz = None if gZ is None else np.ascontiguousarray(np.asarray(gZ, dtype=rt).reshape(B))
    fn = lib().vlgae_oracle_dmv_log_f64 if f64 else lib().vlgae_oracle_dmv_log_f32
    fn.restype = ctypes.c_int
    rc = fn(_p(dec, ctypes.c_float), _p(attach, ctypes.c_float), _p(lengths, ctypes.c_int64), ctypes.c_int(B),
            ctypes.c_int(N), ctypes.c_float(mask_zero), ctypes.c_int(int(trim)), _p(gz, ct), _p(Z, ct), _p(gdec, ct),
            _p(gatt, ct))
    if rc:
        raise ValueError(f"oracle dmv_log failed: rc={rc} (2 = length outside [1, N-1])")
    return Z, gdec, gatt


def dmv_viterbi(dec, attach, lengths, *, mask_zero=NEGINF, trim=False, f64=False):
    """Max semiring + first-max back-pointer decode.

    Returns ``(best [B], heads [B,N] int64, arcs [B,N,N,2], gdec [B,N,2,2,2])`` where ``arcs`` is
    ``DMV1o.argmax`` and ``heads[b, c]`` is the head of word ``c`` (0 = ROOT; column 0 and padding are 0).
    """
    dec, attach = _f32(dec), _f32(attach)
    lengths = np.ascontiguousarray(np.asarray(lengths, dtype=np.int64))
    B, N = dec.shape[:2]
    rt, ct = (np.float64, ctypes.c_double) if f64 else (np.float32, ctypes.c_float)
    best = np.zeros(B, dtype=rt)
    heads = np.zeros((B, N), dtype=np.int64)
    arcs = np.zeros((B, N, N, 2), dtype=rt)
    gdec = np.zeros((B, N, 2, 2, 2), dtype=rt)
    fn = lib().vlgae_oracle_dmv_viterbi_f64 if f64 else lib().vlgae_oracle_dmv_viterbi_f32
    fn.restype = ctypes.c_int
    rc = fn(_p(dec, ctypes.c_float), _p(attach, ctypes.c_float), _p(lengths, ctypes.c_int64), ctypes.c_int(B),
            ctypes.c_int(N), ctypes.c_float(mask_zero), ctypes.c_int(int(trim)), _p(best, ct),
            _p(heads, ctypes.c_int64), _p(arcs, ct), _p(gdec, ct))
    if rc:
        raise ValueError(f"oracle dmv_viterbi failed: rc={rc}")
    return best, heads, arcs, gdec


def deptree(arc, lengths=None, *, semiring="log", fill=NEGINF, mask_zero=NEGINF, f64=False):
    """Arc-factored projective CRF (``DependencyCRF``), deptree.py:25-76.

    Returns ``(value [B], marginals-or-indicator [B,N,N], heads [B,N])``.
    """
    arc = _f32(arc)
    B, N = arc.shape[:2]
    if lengths is None:
        lengths = np.full(B, N - 1)
    lengths = np.ascontiguousarray(np.asarray(lengths, dtype=np.int64))
    rt, ct = (np.float64, ctypes.c_double) if f64 else (np.float32, ctypes.c_float)
    out = np.zeros(B, dtype=rt)
    marg = np.zeros((B, N, N), dtype=rt)
    heads = np.zeros((B, N), dtype=np.int64)
    fn = lib().vlgae_oracle_deptree_f64 if f64 else lib().vlgae_oracle_deptree_f32
    fn.restype = ctypes.c_int
    rc = fn(_p(arc, ctypes.c_float), _p(lengths, ctypes.c_int64), ctypes.c_int(B), ctypes.c_int(N),
            ctypes.c_float(fill), ctypes.c_float(mask_zero), ctypes.c_int(1 if semiring == "max" else 0), _p(out, ct),
            _p(marg, ct), _p(heads, ctypes.c_int64))
    if rc:
        raise ValueError(f"oracle deptree failed: rc={rc}")
    return out, marg, heads


def gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask, neg=-INF):
    """``attmap[b,a,q,v] = <txt[b,q,:], vis[a,v,:]>`` then the two masked fills (joint.py:406-419)."""
    vis_feat, txt_feat = _f32(vis_feat), _f32(txt_feat)
    A, V, D = vis_feat.shape
    B, Q, _ = txt_feat.shape
    att = (txt_feat.reshape(B * Q, D) @ vis_feat.reshape(A * V, D).T).reshape(B, Q, A, V).transpose(0, 2, 1, 3)
    att = np.ascontiguousarray(att)
    vm = np.asarray(vis_mask, dtype=bool)[None, :, None, :]  # align_as -> [1, A, 1, V]
    tm = np.asarray(txt_mask, dtype=bool)[:, None, :, None]  # align_as -> [B, 1, Q, 1]
    att[np.broadcast_to(~vm, att.shape)] = neg
    att[np.broadcast_to(~tm, att.shape)] = neg
    return att


def gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal, neg=-INF):
    """max over V, marginal-weighted mean over Q (joint.py:421-432) -> [B, A]."""
    att = gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask, neg)
    maxatt = att.max(-1)
    tm = _f32(txt_marginal)
    return (maxatt * tm[:, None, :]).sum(-1) / tm.sum(1, keepdims=True)


def loss_grounding_factor_ce(attmap, txt_marginal, vis_mask, prior=None, vis2txt=True):
    """The two cross-entropy sums of ``loss_grounding_factor_ce`` (joint.py:439-491) in float64 numpy:
    ``txt2vis = -sum(diag(log_softmax_A(attmap.max(V))) * txt_marginal)`` (:473-477) and
    ``vis2txt = -sum(diag(log_softmax_B(attmap.max(Q))) * vis_mask)`` (:480-483), after the POS prior
    ``attmap[b, b, 1:T+1, outside the group] -= mask * 100`` (:446-470).  prior: list of (mask [B,T,1] bool, lo, hi)."""
    att = np.array(attmap, dtype=np.float64)
    B = att.shape[0]
    ar = np.arange(B)
    for mask, lo, hi in (prior or []):
        T = mask.shape[1]
        m = mask.astype(np.float64) * 100
        att[ar, ar, 1:T + 1, :lo] -= m
        att[ar, ar, 1:T + 1, hi:] -= m

    def lsm(x, axis):
        mx = x.max(axis=axis, keepdims=True)
        return x - mx - np.log(np.exp(x - mx).sum(axis=axis, keepdims=True))

    logit = lsm(att.max(-1), 1)                      # [B, A, Q], softmax over A
    txt2vis = -(logit[ar, ar] * np.asarray(txt_marginal, dtype=np.float64)).sum()
    v2t = None
    if vis2txt:
        logit = lsm(att.max(2), 0)                   # [B, A, V], softmax over B
        v2t = -(logit[ar, ar] * np.asarray(vis_mask, dtype=np.float64)).sum()
    return txt2vis, v2t


def word_factor_attention(vis_feat, txt_feat, vis_mid):
    """``softmax_v(<txt[b,q,:], vis[b,v,:]>) @ vis_mid[b]`` (joint.py:668-673; the caller drops ROOT with ``[:, 1:]``
    before and adds the residual + LayerNorm after)."""
    vis_feat, txt_feat, vis_mid = _f32(vis_feat), _f32(txt_feat), _f32(vis_mid)
    s = np.einsum("bvd,bqd->bqv", vis_feat, txt_feat)
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(-1, keepdims=True)
    return np.einsum("bqv,bvh->bqh", p, vis_mid)


def dmv_scores(x1, x2, token, dec_score, root_score, head_mask=None, one=0.0, zero=NEGINF, neg=-INF):
    """Score-tensor construction of ``DiscriminativeNDMV._forward`` (/root/reference/src/model/ldndmv.py:184-209) on the
    projected operands of ``DMVFactorizedBilinear`` (/root/reference/src/model/nn/dmv_spec.py:68-76), followed by
    ``merge``.  x1 [B,n,2,2,r], x2 [T,2,2,r], token [B,n], dec_score [B,n,2(decision),2,2], root_score [T].
    Returns (attach [B,n,n,2], dec [B,n,2,2,2], root [B,n], merged_dec, merged_attach), fp32 like the reference."""
    x1, x2, dec_score, root_score = _f32(x1), _f32(x2), _f32(dec_score), _f32(root_score)
    token = np.asarray(token, dtype=np.int64)
    B, n = token.shape

    def lsm(x, axis):
        mx = x.max(axis=axis, keepdims=True)
        return x - mx - np.log(np.exp(x - mx).sum(axis=axis, keepdims=True, dtype=np.float32))

    rule = lsm(np.einsum("bhdve,cdve->bhcdv", x1, x2).astype(np.float32), 2)          # :185
    idx = np.broadcast_to(token.reshape(B, 1, n, 1, 1), (B, n, n, 2, 2))
    prob = np.take_along_axis(rule, idx, axis=2)                                         # :188-189
    left = np.tril(np.ones((n, n), dtype=np.float32), -1)[None, :, :, None]
    right = np.triu(np.ones((n, n), dtype=np.float32), 1)[None, :, :, None]
    attach = prob[..., LEFT, :] * left + prob[..., RIGHT, :] * right                     # :190-193
    if head_mask is not None:
        attach = np.where(np.asarray(head_mask, dtype=bool)[:, :, None, None], np.float32(neg), attach)  # :194-198
    dec = lsm(np.transpose(dec_score, (0, 1, 3, 4, 2)), -1)                              # :202
    root = lsm(root_score, -1)[token]                                                    # :206-207
    md, ma = merge(dec, attach, root, one, zero)                                         # :209
    return attach.astype(np.float32), dec.astype(np.float32), root.astype(np.float32), md, ma


def vis_factors(inputs, w_box, b_box, w_rel, b_rel, w_attr, b_attr, box_mask, add_image=True, slope=0.01):
    """``VisBoxRelSimpleEncoder.forward`` (/root/reference/src/model/vis_encoder/box_rel.py:42-52) followed by the
    concatenation and mask of ``vis_feat_unprune`` (/root/reference/src/model/joint.py:140-172), LITERALLY: the relation
    MLP is applied to the n^2 pairwise means of the inputs (no collapse -- this is the checker).
    inputs [B, n, F]; w_* [H, F], b_* [H]; returns (mid [B, V, H] float32, mask [B, V] bool)."""
    x = _f32(inputs)
    B, n, _ = x.shape

    def mlp(t, w, b):
        y = (t.astype(np.float64) @ np.asarray(w, dtype=np.float64).T + np.asarray(b, dtype=np.float64))
        return np.where(y > 0, y, y * slope)

    pair = (x[:, None, :, :].astype(np.float64) + x[:, :, None, :]) / 2          # box_rel.py:45
    rel = mlp(pair, w_rel, b_rel).reshape(B, n * n, -1)                           # :46-48
    box = mlp(x, w_box, b_box)                                                    # :50
    feats, masks = [box, rel], []
    bm = np.asarray(box_mask, dtype=bool)
    rel_mask = np.triu(bm[:, None, :] & bm[:, :, None], 1).reshape(B, -1)         # joint.py:151-155
    masks = [bm, rel_mask]
    if w_attr is not None:
        feats.append(mlp(x, w_attr, b_attr)); masks.append(bm)                    # :161-164
    if add_image:
        feats.append(box.mean(1, keepdims=True)); masks.append(np.ones((B, 1), dtype=bool))  # :165-176
    return np.concatenate(feats, 1).astype(np.float32), np.concatenate(masks, 1)
