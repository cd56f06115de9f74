"""CPU: the C-ABI library loads and exports every symbol include/sow_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|size_t|char\s*\*|const char\s*\*)\s+\*?\s*((?:sow|tt)_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ["sow_group_fwd", "sow_group_bwd", "sow_group_workspace_bytes", "sow_merge_grouped", "sow_thin_qr", "sow_thin_qr_workspace_bytes",
                 "tt_project", "tt_project_workspace_bytes", "tt_project2_workspace_bytes", "tt_adam2_fused_workspace_bytes", "tt_interleave", "tt_deinterleave", "tt_matmul_rk", "tt_gather2", "tt_project2", "tt_reconstruct2", "tt_adam_fused2", "tt_adam2_head", "tt_adam2_fused", "tt_adam2_workspace_bytes", "tt_adam2_step", "tt_adam_interleaved", "tt_adam_dense",
                 "sow_adam_multi", "sow_last_error", "sow_abi_version"]:
        assert must in names, must


def test_library_exports_every_declared_symbol():
    from sow_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/sow_b200.h but not exported"


def test_ctypes_signatures_cover_the_header():
    from sow_b200 import _lib
    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols()
    lib = _lib.load()
    assert lib.sow_abi_version() == 3
    assert lib.sow_rank_pad(50) == 64 and lib.sow_rank_pad(8) == 64 and lib.sow_rank_pad(65) == 128
    gm = (_lib.GroupMember * 1)()
    gm[0].out_features, gm[0].r, gm[0].scale = 2736, 50, 1.0
    assert lib.sow_group_workspace_bytes(_lib.OP_LINEAR_BWD, 4096, 1024, gm, 1) >= 1024 * 64 * 4
    assert lib.sow_merge_table_stride() % 128 == 0


def test_missing_library_fails_loudly(monkeypatch):
    from sow_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsow_b200.so")
    with pytest.raises(_lib.SowB200Error, match="no CPU fallback"):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    import torch
    from sow_b200 import ops
    from sow_b200._lib import SowB200Error
    x = torch.zeros(8, 16, dtype=torch.bfloat16)
    with pytest.raises(SowB200Error, match="no CPU fallback"):
        ops.linear_fwd(x, None, torch.zeros(16, 4, dtype=torch.bfloat16), torch.zeros(4, 8, dtype=torch.bfloat16), None, 1.0)
    with pytest.raises(SowB200Error):
        ops.thin_qr(torch.zeros(8, 4), 2)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under sow_b200/ or tn_gradient/ may reference it."""
    for pkg in ("sow_b200", "tn_gradient"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in src.replace("sow_oracle_free", ""), os.path.join(dirpath, f)


def test_workspace_queries_run_without_a_gpu():
    """The *_workspace_bytes entry points are pure host arithmetic: callable here, positive for supported shapes, 0 where the
    one-call TT entry points do not take the ranks (callers then keep the op-by-op path)."""
    from sow_b200 import _lib
    lib = _lib.load()
    assert lib.sow_thin_qr_workspace_bytes(4096, 64, 2) > lib.sow_thin_qr_workspace_bytes(4096, 8, 2) > 0
    assert lib.sow_thin_qr_workspace_bytes(4096, 100, 1) == 4096 * 100 * 4          # Gram-Schmidt path above rank 64
    assert lib.tt_project_workspace_bytes(4096, 4096, 8, 1) > 0
    assert lib.tt_project2_workspace_bytes(64, 64, 16) >= lib.tt_project_workspace_bytes(4096, 4096, 16, 1)
    assert lib.tt_adam2_workspace_bytes(64, 64) > lib.tt_adam2_fused_workspace_bytes(64, 64) > 0
    ok = (ctypes.c_int * 4)(1, 8, 8, 1)
    bad_rank = (ctypes.c_int * 4)(1, 128, 8, 1)
    bad_edge = (ctypes.c_int * 4)(2, 8, 8, 1)
    assert lib.tt_nd_workspace_bytes(16, 16, 3, ok) > 0
    assert lib.tt_adam_nd_workspace_bytes(16, 16, 3, ok) > lib.tt_nd_workspace_bytes(16, 16, 3, ok)
    assert lib.tt_nd_workspace_bytes(16, 16, 3, bad_rank) == 0
    assert lib.tt_adam_nd_workspace_bytes(16, 16, 3, bad_edge) == 0
    assert lib.tt_adam_nd_workspace_bytes(16, 16, 2, (ctypes.c_int * 3)(1, 8, 1)) == 0   # order 2 has its own entry point
