"""ctypes binding of libcausalgen_b200.so (include/causalgen_b200.h).

The product path fails loudly when the CUDA library is missing or the device is not sm_100:
there is no CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAUSALGEN_B200_LIB", os.path.join(_HERE, "libcausalgen_b200.so"))

CG_MAX_SRC = 3
CG_MAX_SEG = 4
ACT_NONE, ACT_RELU, ACT_GELU, ACT_LRELU = 0, 1, 2, 3
BF16, F32 = 0, 1


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ns", C.c_int64), ("C", C.c_int32), ("c8", C.c_int32)]


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("add", C.c_void_p), ("add2", C.c_void_p), ("mul", C.c_void_p),
                ("ns", C.c_int64), ("add_ns", C.c_int64), ("add2_ns", C.c_int64), ("mul_ns", C.c_int64),
                ("c0", C.c_int32), ("cn", C.c_int32), ("dtype", C.c_int32), ("mul_act", C.c_int32),
                ("out_act", C.c_int32), ("_pad", C.c_int32), ("act_copy", C.c_void_p), ("act_copy_ns", C.c_int64)]


class ConvArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("ksize", C.c_int32), ("act", C.c_int32),
                ("nsrc", C.c_int32), ("nseg", C.c_int32), ("cout", C.c_int32),
                ("src", Src * CG_MAX_SRC), ("seg", Seg * CG_MAX_SEG),
                ("wpack", C.c_void_p), ("bias", C.c_void_p), ("bias_n", C.c_int32), ("nc", C.c_int32),
                ("fold", C.c_int32), ("_pad2", C.c_int32)]


class PackDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("cout_l", C.c_int32), ("cin_l", C.c_int32), ("k", C.c_int32),
                ("transpose", C.c_int32), ("taps", C.c_int32), ("n_pad", C.c_int32), ("nc", C.c_int32),
                ("n_off", C.c_int32), ("n_log", C.c_int32), ("nsrc", C.c_int32),
                ("src_c", C.c_int32 * CG_MAX_SRC), ("src_log", C.c_int32 * CG_MAX_SRC),
                ("src_off", C.c_int32 * CG_MAX_SRC), ("fold", C.c_int32), ("n_scale", C.c_void_p)]


class WgradArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("ksize", C.c_int32), ("act", C.c_int32),
                ("nsrc", C.c_int32), ("src", Src * CG_MAX_SRC), ("dy", C.c_void_p), ("dy_ns", C.c_int64),
                ("dy_c", C.c_int32), ("dy_c8", C.c_int32), ("dw", C.c_void_p), ("dbias", C.c_void_p), ("cout_l", C.c_int32),
                ("cin_l", C.c_int32), ("src_log", C.c_int32 * CG_MAX_SRC), ("src_off", C.c_int32 * CG_MAX_SRC),
                ("taps", C.c_int32), ("min_tiles", C.c_int32)]


class LatentArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("p", C.c_void_p), ("q_ld", C.c_int32), ("p_ld", C.c_int32),
                ("eps", C.c_void_p), ("seed", C.c_uint64), ("offset", C.c_uint64), ("seed_dev", C.c_void_p),
                ("log_t", C.c_float),
                ("z_bf16", C.c_void_p), ("z_ns", C.c_int64), ("z_f32", C.c_void_p), ("eps_out", C.c_void_p),
                ("kl_out", C.c_void_p), ("N", C.c_int32), ("HW", C.c_int32), ("zdim", C.c_int32),
                ("mode", C.c_int32), ("kl_ch", C.c_void_p), ("kl_elem", C.c_void_p)]


class LatentBwdArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("p", C.c_void_p), ("q_ld", C.c_int32), ("p_ld", C.c_int32),
                ("eps", C.c_void_p), ("seed", C.c_uint64), ("offset", C.c_uint64), ("seed_dev", C.c_void_p),
                ("dz", C.c_void_p), ("dz_ns", C.c_int64), ("g_kl", C.c_float),
                ("dq", C.c_void_p), ("dq_ns", C.c_int64), ("dp", C.c_void_p), ("dp_ns", C.c_int64),
                ("N", C.c_int32), ("HW", C.c_int32), ("zdim", C.c_int32), ("mode", C.c_int32),
                ("g_kl_dev", C.c_void_p), ("log_t", C.c_float), ("_pad", C.c_int32),
                ("kl_gate", C.c_void_p)]


class DGaussArgs(C.Structure):
    _fields_ = [("h", C.c_void_p), ("h_ns", C.c_int64), ("Cw", C.c_int32), ("_pad0", C.c_int32), ("x", C.c_void_p),
                ("w_loc", C.c_void_p), ("b_loc", C.c_void_p), ("w_ls", C.c_void_p), ("b_ls", C.c_void_p),
                ("w_co", C.c_void_p), ("b_co", C.c_void_p),
                ("N", C.c_int32), ("HW", C.c_int32), ("C", C.c_int32),
                ("nll", C.c_void_p), ("g", C.c_float), ("dh", C.c_void_p), ("dh_ns", C.c_int64),
                ("dw_loc", C.c_void_p), ("db_loc", C.c_void_p), ("dw_ls", C.c_void_p), ("db_ls", C.c_void_p),
                ("dw_co", C.c_void_p), ("db_co", C.c_void_p)]


class DmolArgs(C.Structure):
    _fields_ = [("h", C.c_void_p), ("h_ns", C.c_int64), ("Cw", C.c_int32), ("_pad0", C.c_int32), ("x", C.c_void_p),
                ("w", C.c_void_p), ("b", C.c_void_p), ("N", C.c_int32), ("HW", C.c_int32), ("nll", C.c_void_p),
                ("g", C.c_float), ("dh", C.c_void_p), ("dh_ns", C.c_int64), ("dw", C.c_void_p), ("db", C.c_void_p)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "cg_version": (C.c_int, []),
    "cg_last_error": (C.c_char_p, []),
    "cg_device_sms": (C.c_int, []),
    "cg_conv2d": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "cg_conv_nchunk": (C.c_int32, [C.c_int32, C.c_int32]),
    "cg_conv_nchunk_ex": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "cg_conv_fold_ok": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "cg_packed_weight_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "cg_packed_weight_bytes_nc": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "cg_pack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "cg_conv2d_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    "cg_conv2d_wgrad_launches": (C.c_int32, [C.POINTER(WgradArgs)]),
    "cg_stem_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                              C.c_int32, C.c_int64, C.c_void_p]),
    "cg_stem_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_int64, C.c_void_p]),
    "cg_avgpool_fwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_int64] * 2 + [C.c_int32, C.c_void_p]),
    "cg_avgpool_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_int64] * 2 + [C.c_int32] * 2 +
                       [C.c_void_p]),
    "cg_upsample_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_int64] * 2 + [C.c_void_p]),
    "cg_upsample_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_int64] * 2 +
                        [C.c_int32, C.c_void_p]),
    "cg_latent_fwd": (C.c_int, [C.POINTER(LatentArgs), C.c_void_p]),
    "cg_latent_bwd": (C.c_int, [C.POINTER(LatentBwdArgs), C.c_void_p]),
    "cg_free_bits": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "cg_latent_mix": (C.c_int, [C.c_void_p] * 6 + [C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "cg_dgauss_nll_fwd": (C.c_int, [C.POINTER(DGaussArgs), C.c_void_p]),
    "cg_dgauss_nll_bwd": (C.c_int, [C.POINTER(DGaussArgs), C.c_void_p]),
    "cg_dgauss_sample": (C.c_int, [C.POINTER(DGaussArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "cg_dgauss_sample_bwd": (C.c_int, [C.POINTER(DGaussArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_dmol_loss_fwd": (C.c_int, [C.POINTER(DmolArgs), C.c_void_p]),
    "cg_dmol_loss_bwd": (C.c_int, [C.POINTER(DmolArgs), C.c_void_p]),
    "cg_dmol_predict": (C.c_int, [C.POINTER(DmolArgs), C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "cg_cf_combine": (C.c_int, [C.c_void_p] * 8 + [C.c_int64, C.c_void_p]),
    "cg_cf_combine_bwd": (C.c_int, [C.c_void_p] * 10 + [C.c_int64, C.c_void_p]),
    "cg_normalise_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cg_augment_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p]),
    "cg_parents_plane": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "cg_nchw_f32_to_planar": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_int64, C.c_void_p]),
    "cg_planar_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_int64, C.c_void_p]),
    "cg_stats_to_nchw": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]),
    "cg_fill_planar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "cg_colsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "cg_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_int64] * 3 + [C.c_void_p]),
    "cg_elbo_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p]),
    "cg_bn_fold": (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "cg_conv_direct_fwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 9 + [C.c_int64, C.c_void_p]),
    "cg_pool_max_fwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_int64, C.c_int64, C.c_void_p]),
    "cg_groupnorm_fwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_float, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p] +
                         [C.c_int32] * 3 + [C.c_int64, C.c_int64, C.c_void_p]),
    "cg_global_avgpool": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_int64, C.c_int32, C.c_void_p]),
    "cg_linear": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p] + [C.c_int32] * 4 +
                  [C.c_void_p]),
    "cg_sumsq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "cg_optim_advance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32] +
                         [C.c_float] * 6 + [C.c_int32, C.c_void_p]),
    "cg_adamw_ema_step": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_float] * 4 +
                          [C.c_void_p]),
}

EXPORTED = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (no compute, no GPU needed)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "causalgen_b200 has no CPU/PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().cg_last_error().decode()


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"causalgen_b200 {what} failed (status {rc}): {last_error()}")


class Launch:
    """One recorded kernel launch: a bound C function plus its (mutable) argument tuple."""
    __slots__ = ("fn", "args", "name", "keep", "side", "lane", "algo_bytes")

    def __init__(self, name, *args):
        self.fn = getattr(load(), name)
        self.args = tuple(args)
        self.name = name
        self.keep = None
        self.side = False  # True: nothing later in the program reads its result -> may run on a side stream
        self.lane = 0      # 0: main stream, 1: auxiliary lane of the program (engine.Program)
        self.algo_bytes = 0  # algorithmic HBM bytes of the launch (logical channels), set by ops.ConvLayer

    def __call__(self, stream):
        rc = self.fn(*self.args, stream)
        if rc != 0:
            raise RuntimeError(f"causalgen_b200 {self.name} failed (status {rc}): {last_error()}")
