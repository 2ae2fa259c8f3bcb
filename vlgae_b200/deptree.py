"""Host wrapper of the arc-factored (MBR) chart kernel: ``DependencyCRF`` semantics of the reference
(/root/reference/src/model/torch_struct/deptree.py:25-76,146-162)."""
from __future__ import annotations

import torch

from ._lib import VlgaeError, check, lib
from .ops import MASK_ZERO

_ws = {}


def run(arc, lengths, semiring, want_marg):
    """Returns ``(value [B], marginals-or-indicator [B,N,N] | None)``."""
    from .torch_struct.semirings import semirings as _sr

    if arc.dim() == 4:  # labeled potentials: semiring-sum over labels first (deptree.py:41)
        arc = torch.logsumexp(arc, -1) if semiring.name == "log" else arc.max(-1).values
    dev = arc.device
    if dev.type != "cuda":
        raise VlgaeError("vlgae_b200 operators need CUDA tensors (there is no CPU fallback)")
    B, N, N2 = arc.shape
    if N != N2:
        raise AssertionError("Non-square potentials")  # deptree.py:149
    if lengths is None:
        lengths = torch.full((B,), N - 1, dtype=torch.int64, device=dev)
    lengths = torch.as_tensor(lengths).to(device=dev, dtype=torch.int64).contiguous()
    arc = arc.detach().to(torch.float32).contiguous()
    out = torch.empty(B, dtype=torch.float32, device=dev)
    marg = torch.empty((B, N, N), dtype=torch.float32, device=dev) if want_marg else None
    need = lib().vlgae_deptree_workspace_bytes(B, N)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)  # per stream: concurrent calls must not share scratch
    ws = _ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        _ws[key] = ws
    fill = float(_sr.NEGINF)  # looked up at call time, like the reference's zero_() (quirk Q1)
    with torch.cuda.device(dev):
        check(lib().vlgae_deptree(arc.data_ptr(), lengths.data_ptr(), B, N, fill, MASK_ZERO,
                                  1 if semiring.name == "max" else 0, out.data_ptr(),
                                  None if marg is None else marg.data_ptr(), None, ws.data_ptr(), ws.numel(),
                                  torch.cuda.current_stream(dev).cuda_stream), "vlgae_deptree")
    return out, marg
