"""oracle/bruteforce.py -- exhaustive enumeration of single-root projective trees (n <= 7).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  An independent pin for the chart DP: it never
builds a chart, it scores every tree with the valence rule that the reference's gold-tree rule
counter uses (/root/reference/src/model/dmv_helper/good_init_nn.py:34-78):

  for a head h and a direction d with children K (ordered by distance from h)
    K empty : dec[h, d, NOCHILD, STOP]
    else    : the OUTERMOST child c pays dec[h, d, NOCHILD, GO] + attach[h, c, NOCHILD],
              every other child    pays dec[h, d, HASCHILD, GO] + attach[h, c, HASCHILD],
              and the head finally pays dec[h, d, HASCHILD, STOP];
  ROOT (position 0) only has a right side and takes exactly one child.

Inputs are the *merged* tensors (ROOT at position 0) of one sentence.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

NOCHILD, HASCHILD, LEFT, RIGHT, GO, STOP = 1, 0, 0, 1, 0, 1


def _is_single_root_projective_tree(heads):
    """heads[c-1] = head of word c (1-based words, 0 = ROOT)."""
    n = len(heads)
    if sum(1 for h in heads if h == 0) != 1:
        return False
    for c in range(1, n + 1):  # acyclic: every word reaches ROOT
        seen, x = set(), c
        while x != 0:
            if x in seen:
                return False
            seen.add(x)
            x = heads[x - 1]
    arcs = [(min(h, c), max(h, c)) for c, h in enumerate(heads, 1)]
    for (a, b), (c, d) in itertools.combinations(arcs, 2):
        if a < c < b < d or c < a < d < b:
            return False
    return True


def tree_score(dec, attach, heads):
    """Score of one tree under merged dec [N,2,2,2] / attach [N,N,2] (float64 accumulation)."""
    n = len(heads)
    total = 0.0
    for h in range(0, n + 1):
        for d in (LEFT, RIGHT):
            if h == 0 and d == LEFT:
                continue
            kids = [c for c in range(1, n + 1) if heads[c - 1] == h and ((c < h) if d == LEFT else (c > h))]
            kids.sort(key=lambda c: abs(c - h))
            if not kids:
                total += float(dec[h, d, NOCHILD, STOP])
                continue
            for k, c in enumerate(kids):
                v = NOCHILD if k == len(kids) - 1 else HASCHILD
                total += float(dec[h, d, v, GO]) + float(attach[h, c, v])
            total += float(dec[h, d, HASCHILD, STOP])
    return total


def enumerate_trees(n):
    return [hs for hs in itertools.product(range(0, n + 1), repeat=n)
            if all(h != c for c, h in enumerate(hs, 1)) and _is_single_root_projective_tree(hs)]


def brute(dec, attach, length):
    """Returns (logZ, best, best_heads, arc_marginals [N,N], n_trees) for one sentence of `length` words."""
    dec = np.asarray(dec, dtype=np.float64)
    attach = np.asarray(attach, dtype=np.float64)
    N = dec.shape[0]
    trees = enumerate_trees(length)
    scores = np.array([tree_score(dec, attach, t) for t in trees])
    m = scores.max()
    logZ = m + math.log(np.exp(scores - m).sum())
    post = np.exp(scores - logZ)
    marg = np.zeros((N, N))
    for t, p in zip(trees, post):
        for c, h in enumerate(t, 1):
            marg[h, c] += p
    k = int(scores.argmax())
    return logZ, float(m), list(trees[k]), marg, len(trees)
