"""CPU-side checks of the boundary: the library builds/loads and exports exactly what include/vlgae_b200.h
declares; the Python mirror exposes the reference's names; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vlgae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlgae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as entry
    from vlgae_b200 import _lib

    entry.build()
    declared = _declared_symbols()
    assert len(declared) >= 10
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == declared, "ctypes prototypes out of sync with the header"
    assert _lib.lib().vlgae_version() >= 1


def test_invalid_arguments_return_codes():
    from vlgae_b200._lib import lib

    L = lib()
    # null pointers / bad N are rejected before any CUDA call
    assert L.vlgae_dmv_inside_outside(None, None, None, 1, 4, -1e12, None, None, None, None, None, 0, None) == 1
    assert b"non-null" in L.vlgae_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.vlgae_dmv_viterbi(p, p, p, 1, 1000, -1e12, p, None, None, None, None, 0, None) == 1
    assert L.vlgae_dmv_workspace_bytes(0, 10) == 0


def test_reference_surface_names():
    import vlgae_b200.torch_struct as ts
    from vlgae_b200.torch_struct import DMV1o, DependencyCRF  # noqa: F401
    from vlgae_b200.torch_struct.dmv import GO, HASCHILD, LEFT, NOCHILD, RIGHT, STOP

    assert (LEFT, RIGHT, HASCHILD, NOCHILD, GO, STOP) == (0, 1, 0, 1, 0, 1)
    assert ts.version == "0.4"
    # src.setup_inf rebinding (reference src/__init__.py:113-117) must work on the mirror
    old = ts.semirings.semirings.NEGINF
    ts.semirings.semirings.NEGINF = -1e20
    assert ts.LogSemiring.zero == -1e12  # class attribute stays frozen, as in the reference
    ts.semirings.semirings.NEGINF = old


def test_no_cpu_fallback():
    from vlgae_b200._lib import VlgaeError
    from vlgae_b200.torch_struct import DMV1o

    dec = torch.zeros(1, 3, 2, 2, 2)
    attach = torch.zeros(1, 3, 3, 2)
    with pytest.raises(VlgaeError):
        DMV1o([dec, attach], torch.tensor([2])).partition
    with pytest.raises(VlgaeError):
        DMV1o.merge(torch.zeros(1, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2), torch.zeros(1, 2))


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vlgae_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src:
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


def test_integration_alias_recipe():
    """aliasing vlgae_b200.torch_struct under the reference's module paths makes the
    reference's own import statements (ldndmv.py:21-22, joint.py:20, src/__init__.py:113-117) resolve to the mirror."""
    import importlib
    import sys
    import types

    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    try:
        for k in saved:
            del sys.modules[k]
        # stand-ins for the reference's package skeleton (src/__init__.py needs hydra / lightning, absent here)
        for pkg in ("src", "src.model"):
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
        for name in ("", ".dmv", ".distributions", ".semirings", ".semirings.semirings"):
            sys.modules["src.model.torch_struct" + name] = importlib.import_module("vlgae_b200.torch_struct" + name)
        ns = {}
        exec("from src.model.torch_struct import DMV1o, DependencyCRF\n"
             "from src.model.torch_struct.dmv import LEFT, RIGHT, NOCHILD, HASCHILD, GO, STOP\n"
             "import src.model.torch_struct as stt\n", ns)
        assert (ns["LEFT"], ns["RIGHT"], ns["HASCHILD"], ns["NOCHILD"], ns["GO"], ns["STOP"]) == (0, 1, 0, 1, 0, 1)
        import vlgae_b200.torch_struct as vts

        assert ns["DMV1o"] is vts.DMV1o and ns["DependencyCRF"] is vts.DependencyCRF
        old = ns["stt"].semirings.semirings.NEGINF
        ns["stt"].semirings.semirings.NEGINF = -1e20   # what src.setup_inf(1e20) does
        assert vts.semirings.semirings.NEGINF == -1e20
        ns["stt"].semirings.semirings.NEGINF = old
        # unsupported methods raise instead of returning something else
        import pytest
        import torch

        d = vts.DMV1o([torch.zeros(1, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2)], torch.tensor([1]))
        for call in (lambda: d.kmax(2), lambda: d.topk(2), d.entropy, d.sample):
            with pytest.raises(NotImplementedError):
                call()
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
