"""ctypes binding of libb200vqa.so (include/b200vqa.h).

There is no CPU fallback: if the shared library is missing, or a kernel entry point is called without a
B200, the call raises.  Tensors are passed as raw device pointers plus the current CUDA stream, so every
call is asynchronous and CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

from . import _build

F32, BF16 = 0, 1
LAYOUT_K, LAYOUT_MN = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_SILU, ACT_TANH = 0, 1, 2, 3, 4
EPI_NONE, EPI_ACT, EPI_ADD, EPI_DACT, EPI_ACCUM, EPI_ACT_D, EPI_MUL = 0, 1, 2, 3, 4, 5, 6
GROUP_TILE = 128

ACT_CODES = {"none": ACT_NONE, "gelu": ACT_GELU, "relu": ACT_RELU, "silu": ACT_SILU, "tanh": ACT_TANH}

_C = {"p": ctypes.c_void_p, "i": ctypes.c_int, "f": ctypes.c_float, "l": ctypes.c_longlong, "z": ctypes.c_size_t}

# name -> (restype code, argument codes); mirrors include/b200vqa.h one to one
SIGNATURES = {
    "b200_init": ("i", "i"),
    "b200_last_error_string": ("s", ""),
    "b200_abi_version": ("i", ""),
    "b200_launch_count": ("l", ""),
    "b200_reset_launch_count": ("v", ""),
    "b200_dropout_mask": ("i", "plpp"),
    "b200_dropout_apply": ("i", "pplipp"),
    "b200_cast": ("i", "pipilp"),
    "b200_colsum_ws": ("z", "ii"),
    "b200_colsum": ("i", "piiipippzp"),
    "b200_dropout_colsum": ("i", "ppiiipppzp"),
    "b200_gemm": ("i", "piipiipiiiiiipiippipp"),
    "b200_ggemm": ("i", "pipipiiiiippiipiippipp"),
    "b200_ggemm_wgrad": ("i", "pipipiiiipip"),
    "b200_glu_fwd": ("i", "pipiiippp"),
    "b200_glu_bwd": ("i", "ppipiiippp"),
    "b200_add_ln_fwd": ("i", "pppppfpppiiipip"),
    "b200_add_ln_bwd_ws": ("z", "ii"),
    "b200_add_ln_bwd": ("i", "pppppppippppiiipippzp"),
    "b200_attn_fwd": ("i", "pipipipipipiiiiifipp"),
    "b200_attn_bwd": ("i", "pipipipipipippipipiiiiiifipp"),
    "b200_embed_fwd": ("i", "ppppiiiiipp"),
    "b200_embed_bwd": ("i", "pppiiiipp"),
    "b200_ce_fwd": ("i", "plpiiifippppp"),
    "b200_ce_bwd": ("i", "plppiiifippplp"),
    "b200_router_ws": ("z", "ii"),
    "b200_router_fwd": ("i", "pipppffiiiippppppppppzp"),
    "b200_router_bwd_ws": ("z", "iii"),
    "b200_router_bwd": ("i", "pipppffiiiipppppppppppppzp"),
    "b200_moe_max_rows": ("i", "ii"),
    "b200_moe_plan_ws": ("z", "ii"),
    "b200_moe_plan": ("i", "piiipppppppppzp"),
    "b200_moe_capacity": ("i", "pppppiiippp"),
    "b200_moe_permute": ("i", "pppiiiiipp"),
    "b200_moe_unpermute": ("i", "pppiiiipp"),
    "b200_moe_combine_fwd": ("i", "pppppfiiiipppp"),
    "b200_moe_combine_bwd_ws": ("z", "ii"),
    "b200_moe_combine_bwd": ("i", "ppppppppiiiiipppppzp"),
    "b200_ep_push_counts": ("i", "ppiiip"),
    "b200_ep_layout": ("i", "piiiiippppp"),
    "b200_ep_dispatch": ("i", "ppppppiiiiiiiip"),
    "b200_ep_return": ("i", "ppppiiiiiip"),
    "b200_p2p_allreduce_f32": ("i", "piillfp"),
    "b200_nvls_allreduce_f32": ("i", "piillfip"),
}

class DropoutT(ctypes.Structure):
    """b200_dropout_t: device pointer to (seed, offset), probability, site id."""
    _fields_ = [("rng_state", ctypes.c_void_p), ("p", ctypes.c_float), ("site", ctypes.c_uint)]


def dropout_arg(drop):
    """(state_tensor, p, site) or None -> ctypes struct or None."""
    if drop is None:
        return None
    st, p, site = drop
    return DropoutT(st.data_ptr(), float(p), int(site) & 0xFFFFFFFF)


_lib = None
_initialized_devices: set[int] = set()


def lib_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libb200vqa.so (building it in-tree with nvcc when absent/stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and os.environ.get("B200VQA_NO_BUILD") != "1":
        try:
            if not _build.is_fresh():
                _build.build()
        except Exception as e:  # nvcc missing on the target box: fall through to the prebuilt file
            if not path.exists():
                raise RuntimeError(f"libb200vqa.so is missing and could not be built: {e}") from e
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: the b200vqa CUDA extension is required (there is no CPU fallback). "
            "Run `python -m vqa_model_builder_b200._build`.")
    lib = ctypes.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = {"s": ctypes.c_char_p, "v": None}.get(res, _C.get(res))
        fn.argtypes = [_C[c] for c in args]
    if lib.b200_abi_version() != 2:
        raise RuntimeError("libb200vqa.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200_last_error_string().decode(errors="replace")


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    if isinstance(x, DropoutT):
        return ctypes.addressof(x)
    return x


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ensure_device(t: torch.Tensor) -> None:
    """The product path is CUDA-only; refuse anything else loudly."""
    if not t.is_cuda:
        raise RuntimeError("b200vqa kernels need CUDA tensors on a B200 (sm_100a); there is no CPU fallback")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _initialized_devices:
        lib = load()
        rc = lib.b200_init(dev)
        if rc != 0:
            raise RuntimeError(f"b200_init({dev}) failed: {last_error()}")
        _initialized_devices.add(dev)


#: set to a list to time every entry-point call with CUDA events on the launching stream:
#: entries are (name, start_event, end_event, scalar_args)
PROFILE = None


def call(name: str, *args):
    """Invoke an int-returning entry point; tensors -> device pointers; non-zero return raises."""
    lib = load()
    fn = getattr(lib, name)
    if PROFILE is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*[_ptr(a) for a in args])
        e.record()
        PROFILE.append((name, s, e, tuple(a for a in args if isinstance(a, (int, float)) and not isinstance(a, bool))))
    else:
        rc = fn(*[_ptr(a) for a in args])
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")


def query(name: str, *args) -> int:
    """Invoke a size/count query (no error protocol)."""
    return int(getattr(load(), name)(*args))


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise RuntimeError(f"b200vqa supports float32 and bfloat16 activations, got {dt}")


def launch_count() -> int:
    return int(load().b200_launch_count())


def reset_launch_count() -> None:
    load().b200_reset_launch_count()
