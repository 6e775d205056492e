"""ctypes binding of libdwn_b200.so (include/dwn_b200.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# DWN_LIB: another build of the same library (A/B timing of two revisions on one GPU box, tests/gpu_checks/run_ab.sh)
_LIB_PATH = Path(os.environ["DWN_LIB"]).resolve() if os.environ.get("DWN_LIB") else \
    Path(__file__).resolve().parent / "libdwn_b200.so"
_lib = None


class DwnError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("A", C.c_void_p), ("B", C.c_void_p), ("a_mn", C.c_int), ("b_mn", C.c_int),
        ("lda", C.c_long), ("ldb", C.c_long), ("a_zstride", C.c_long), ("b_zstride", C.c_long),
        ("a_zmode", C.c_int), ("b_zmode", C.c_int), ("b_batch_rows", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("Z", C.c_int),
        ("epi", C.c_int), ("D", C.c_void_p), ("d_dtype", C.c_int), ("ldd", C.c_long), ("d_zstride", C.c_long),
        ("m_limit", C.c_int), ("n_limit", C.c_int), ("bias", C.c_void_p), ("beta", C.c_float), ("Tn", C.c_int),
        ("n_out_total", C.c_int), ("row_offset_per_z", C.c_int), ("block_n", C.c_int),
        ("A2", C.c_void_p), ("B2", C.c_void_p), ("lda2", C.c_long), ("ldb2", C.c_long), ("K2", C.c_int),
        ("dbg_lbo_a", C.c_uint), ("dbg_sbo_a", C.c_uint), ("dbg_lbo_b", C.c_uint), ("dbg_sbo_b", C.c_uint),
        ("split", C.c_int), ("a_pstride", C.c_long), ("b_pstride", C.c_long),
    ]


_T = {"p": C.c_void_p, "i": C.c_int, "l": C.c_long, "f": C.c_float, "d": C.c_double}

# name -> argument type string (p pointer, i int, l long, f float, d double); mirrors include/dwn_b200.h
SIGNATURES = {
    "dwn_input_moments": "piilpipp",
    "dwn_stem_coef": "pidppppppffpip",
    "dwn_stem_fwd": "ppppppppp" + "iiiiiiii" + "p",
    "dwn_bn_finalize": "pidpppppffipiiip",
    "dwn_colstats": "pliipiip",
    "dwn_sdw_fwd": "ppppp" + "iiiiiii" + "p",
    "dwn_tdw_fwd": "ppppp" + "iiiiii" + "p",
    "dwn_se_pool": "pppp" + "iiiii" + "p",
    "dwn_se_mlp": "pii" + "ppppppp" + "iii" + "p",
    "dwn_fold_gate": "ppp" + "iiii" + "p",
    "dwn_block_out": "ppppp" + "ppp" + "ppp" + "iiiiiiiiiiii" + "p",
    "dwn_pool_hw": "ppp" + "iii" + "p",
    "dwn_cortex_out": "ppppppp" + "iiiiii" + "p",
    "dwn_readout_prep": "pppp" + "iiii" + "p",
    "dwn_cast_bf16": "pplp",
    "dwn_split3": "pplp",
    # backward
    "dwn_bn_bwd_finalize": "piiidpppip",
    "dwn_bn_bwd_finalize2": "piiiidppppppip",
    "dwn_block_bwd_reduce": "ppppppp" + "iiiiiiiiiii" + "p",
    "dwn_block_bwd_dy": "pppppp" + "llii" + "p",
    "dwn_block_in_bwd": "ppppppp" + "iiiiiii" + "p",
    "dwn_block_in_bwd_stem": "pppppppp" + "iiiiiiii" + "p",
    "dwn_stem_bwd_finalize": "pi" + "pppppp" + "iili" + "p",
    "dwn_pool_bwd": "pp" + "lii" + "p",
    "dwn_se_bwd": "ppppppp" + "pppppppp" + "iiii" + "p",
    "dwn_tdw_bwd_reduce": "pppp" + "i" + "p" + "iiii" + "p",
    "dwn_tdw_bwd": "ppppppppp" + "iiiiii" + "p",
    "dwn_sdw_bwd": "ppppppppp" + "iiiiiii" + "p",
    "dwn_bn_bwd_apply": "pppp" + "lii" + "p",
    "dwn_reduce_rows": "pilpp",
    "dwn_dw_wgrad_finalize": "piiiipip",
    "dwn_stem_bwd": "ppp" + "i" + "pppppp" + "iili" + "p",
    "dwn_cortex_bwd_reduce": "ppppppp" + "iiiiiii" + "p",
    "dwn_cortex_bwd_dy": "pppppp" + "iiiii" + "p",
    "dwn_cortex_in_bwd": "pppppp" + "iii" + "p",
    # head / loss / distillation / predictor
    "dwn_poisson_fwd": "ppp" + "iil" + "f" + "p" + "i" + "p",
    "dwn_poisson_bwd": "pppi" + "p" + "il" + "f" + "pp",
    "dwn_readout_bwd_prep": "pp" + "f" + "ppp" + "iiiiiii" + "p",
    "dwn_readout_dx_combine": "pp" + "i" + "p" + "iii" + "p",
    "dwn_distill_prepare": "pifppp",
    "dwn_distill_fill": "ppp" + "iii" + "l" + "p",
    "dwn_distill_weights": "pppip",
    "dwn_window_blend": "ppp" + "iiiiii" + "l" + "p",
    "dwn_window_gather": "pp" + "ii" + "l" + "iiii" + "p",
    "dwn_assemble_clips": "pippp" + "iiiii" + "f" + "iiii" + "p",
    "dwn_assemble_batch": "pipp" + "iiiiii" + "f" + "p",
    "dwn_corr_update": "ppp" + "iiii" + "pp" + "p",
    "dwn_corr_finalize": "pp" + "i" + "d" + "pp" + "p",
    "dwn_cutmix": "pppp" + "i" + "l" + "ii" + "p",
    "dwn_lerp_rows": "pppp" + "i" + "l" + "p",
    "dwn_scatter_mouse_targets": "pp" + "i" + "p" + "iiii" + "p",
    # conv_pw algebra (Gram statistics, BN1-backward folded into GEMMs)
    "dwn_partial_colsum": "piiiipp",
    "dwn_gram_finalize": "piipdppp",
    "dwn_pw_stats": "pppdpppppffpiip",
    "dwn_pw_bwd_prep": "ppppppp" + "ii" + "p",
    "dwn_pw_wgrad_finalize": "ppppppp" + "ii" + "p",
    # optimizer / EMA
    "dwn_adamw": "ppp" + "i" + "pp" + "i" + "ffffff" + "ppp",
    "dwn_ema": "ppp" + "i" + "f" + "p",
    "dwn_scale": "plfp",
    # NCCL behind the C ABI (hosts without torch.distributed)
    "dwn_comm_unique_id": "p",
    "dwn_comm_init": "iip",
    "dwn_allreduce_bucket": "pliiip",
    "dwn_comm_group_begin": "",
    "dwn_comm_group_end": "",
    "dwn_comm_destroy": "",
}


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise DwnError(
                f"{_LIB_PATH} is missing: build it with `python -m sensorium_b200.build` "
                "(sensorium_b200 has no CPU / eager fallback)")
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.dwn_last_error.restype = C.c_char_p
        _lib.dwn_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        _lib.dwn_gemm.restype = C.c_int
        for name, sig in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = [_T[ch] for ch in sig]
            fn.restype = C.c_int
    return _lib


def exported_symbols():
    return ["dwn_last_error", "dwn_abi_version", "dwn_sm_count", "dwn_set_sm_budget", "dwn_gemm", "dwn_block_in_bwd_stem_rows",
            "dwn_pw_bwd_prep_scratch", "dwn_opt_chunk", *SIGNATURES.keys()]


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x


# kernels launched per C-ABI call (default 1) — used for the bench's gpu_launches claim
_LAUNCHES_PER_CALL = {"dwn_input_moments": 2, "dwn_se_bwd": 3, "dwn_adamw": 2, "dwn_stem_bwd": 2, "dwn_pw_bwd_prep": 3}
LAUNCHES = 0
PROF = None  # when set to a list, every call appends (name, tag, bytes, flops, start_event, end_event)


def call(name: str, *args, _bytes: int = 0, _flops: int = 0, _tag: str = ""):
    global LAUNCHES
    fn = getattr(lib(), name)
    prof = PROF
    if prof is not None:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(*[_ptr(a) for a in args])
    if rc != 0:
        raise DwnError(f"{name} failed: {lib().dwn_last_error().decode()}")
    LAUNCHES += _LAUNCHES_PER_CALL.get(name, 1)
    if prof is not None:
        e1.record()
        prof.append((name, _tag, _bytes, _flops, e0, e1))


def gemm(stream, _bytes: int = 0, _flops: int = 0, _tag: str = "", **kw):
    global LAUNCHES
    d = GemmDesc()
    for k, v in kw.items():
        setattr(d, k, _ptr(v))
    prof = PROF
    if prof is not None:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = lib().dwn_gemm(C.byref(d), stream)
    if rc != 0:
        raise DwnError(f"dwn_gemm failed: {lib().dwn_last_error().decode()}")
    LAUNCHES += 1
    if prof is not None:
        e1.record()
        if not _flops:
            _flops = 2 * d.M * d.N * d.K * d.Z
        prof.append(("dwn_gemm", _tag, _bytes, _flops, e0, e1))
