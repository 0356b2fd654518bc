"""The C-ABI shared library loads on a host without a GPU and exports every symbol include/causalgen_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "causalgen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from causalgen_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first: python __graft_entry__.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from causalgen_b200 import _lib
    assert sorted(_lib.EXPORTED) == declared_symbols()
    lib = _lib.load()
    assert lib.cg_version() >= 100


def test_struct_layouts_match_the_header_sizes():
    """ctypes mirrors of the ABI structs: sizes that the C side static-asserts through its own layout"""
    from causalgen_b200 import _lib as L
    assert ctypes.sizeof(L.Src) == 24
    assert ctypes.sizeof(L.Seg) == 104
    assert ctypes.sizeof(L.ConvArgs) == 32 + 3 * 24 + 4 * 104 + 32
    assert ctypes.sizeof(L.PackDesc) % 8 == 0
    assert ctypes.sizeof(L.WgradArgs) == 176 and L.WgradArgs.min_tiles.offset == 172   # include/causalgen_b200.h cg_wgrad_args
    # weight-slab planner is pure host arithmetic and must be callable without a GPU
    lib = L.load()
    nc = lib.cg_conv_nchunk(9 * 4, 16)
    assert nc == 16
    assert lib.cg_packed_weight_bytes(9 * 4, 16) == 9 * 4 * 16 * 32
    assert lib.cg_conv_nchunk(9 * 2, 224) == 64           # four chunks of <= 64 output channels per CTA
    assert lib.cg_conv_nchunk(9 * 2, 128) == 64
    # a huge-K first conv of a Block (posterior, 272 -> 32 channels, 3x3): no operand ring needed -> one chunk of 32
    assert lib.cg_conv_nchunk_ex(9 * 17, 32, 0) == 32 and lib.cg_conv_nchunk_ex(9 * 17, 32, 1) == 16
    assert lib.cg_packed_weight_bytes_nc(9 * 17, 32, 32) == 9 * 17 * 32 * 32
    # column-folded 3x3 (cg_conv_args.fold): one GEMM-N chunk of <= 32 output channels, nine-tap K, slab must fit
    assert lib.cg_conv_fold_ok(9 * 4, 16, 0) == 1 and lib.cg_conv_fold_ok(9 * 8, 32, 1) == 1
    assert lib.cg_conv_fold_ok(9 * 4, 48, 0) == 0 and lib.cg_conv_fold_ok(4, 16, 0) == 0
    assert lib.cg_conv_fold_ok(9 * 40, 32, 1) == 0          # 640 input channels: the slab leaves no room for the rings


def test_weight_gradient_grid_policy():
    """ops.wgrad_min_tiles: the measured optimum of profiles/r4g_wgrad_grid_and_stem.txt (192 tiles at 128 images per GPU, 96
    at 32) for models of light Blocks, 48 for the four-conv GELU configs"""
    from causalgen_b200 import ops
    assert ops.wgrad_min_tiles(128, True) == 192 and ops.wgrad_min_tiles(32, True) == 96
    assert ops.wgrad_min_tiles(1, True) == 24 and ops.wgrad_min_tiles(4096, True) == 256
    assert ops.wgrad_min_tiles(1024, False) == 48 and ops.wgrad_min_tiles(64, False) == 48


def test_fold_policy_follows_the_measured_layers():
    """ops.fold_pays: the A/B of profiles/r3r_conv_ab_folded_final_kernel.txt as a rule (wide K, tile count within 20 %)"""
    from causalgen_b200 import ops
    if ops.FOLD != 1:
        pytest.skip("CAUSALGEN_B200_FOLD overrides the policy")
    assert ops.fold_pays(96, 64) and ops.fold_pays(24, 128) and ops.fold_pays(192, 64) and ops.fold_pays(12, 160)
    assert not ops.fold_pays(48, 96) and not ops.fold_pays(48, 208)     # 24 folded tiles against 18 per image
    assert not ops.fold_pays(192, 32)                                    # two K-blocks: nothing to save
    assert not ops.fold_pays(None, 256)                                  # resolution unknown at construction


def test_compute_entry_fails_loudly_without_b200():
    import torch
    if torch.cuda.is_available():
        return
    from causalgen_b200 import _lib as L
    lib = L.load()
    rc = lib.cg_cf_combine(None, None, None, None, None, None, None, None, 0, None)
    assert rc == -2  # CG_ERR_ARCH
    assert b"no CPU" in lib.cg_last_error() or b"sm_100" in lib.cg_last_error()
