"""CPU: the C-ABI library builds, loads without a GPU and exports exactly what include/b200vqa.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "b200vqa.h"


def parse_header():
    src = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"(int|size_t|long long|void|const char\*)\s+(b200_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.groups()
        codes = ""
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    codes += "p"
                elif a.startswith("long long"):
                    codes += "l"
                elif a.startswith("size_t"):
                    codes += "z"
                elif a.startswith("float"):
                    codes += "f"
                elif a.startswith("int"):
                    codes += "i"
                else:
                    raise ValueError(a)
        out[name] = ({"int": "i", "size_t": "z", "long long": "l", "void": "v", "const char*": "s"}[ret], codes)
    return out


def test_header_declares_the_hot_path():
    names = set(parse_header())
    for required in ("b200_gemm", "b200_ggemm", "b200_ggemm_wgrad", "b200_attn_fwd", "b200_attn_bwd",
                     "b200_router_fwd", "b200_router_bwd", "b200_moe_plan", "b200_moe_permute",
                     "b200_moe_combine_fwd", "b200_moe_combine_bwd", "b200_add_ln_fwd", "b200_add_ln_bwd"):
        assert required in names


def test_library_builds_loads_and_exports_every_symbol():
    from vqa_model_builder_b200 import _build
    path = _build.build()
    lib = ctypes.CDLL(str(path))
    for name in parse_header():
        assert hasattr(lib, name), f"{name} declared in b200vqa.h but not exported"
    lib.b200_abi_version.restype = ctypes.c_int
    assert lib.b200_abi_version() == 2


def test_ctypes_signatures_match_header():
    from vqa_model_builder_b200 import _lib
    hdr = parse_header()
    assert set(hdr) == set(_lib.SIGNATURES)
    for name, sig in hdr.items():
        assert _lib.SIGNATURES[name] == sig, name


def test_size_queries_work_without_gpu():
    from vqa_model_builder_b200 import _lib
    assert _lib.query("b200_moe_max_rows", 64, 8) == 1152          # 64 + 8*127 rounded up to 128
    assert _lib.query("b200_moe_max_rows", 29184, 8) % 128 == 0
    assert _lib.query("b200_router_ws", 1024, 8) > 0
    assert _lib.query("b200_add_ln_bwd_ws", 2048, 768) >= (2048 // 8) * 3 * 768 * 4
    assert _lib.query("b200_add_ln_bwd_ws", 32, 768) > 4 * 3 * 768 * 4     # the staged kernel also keeps a group map


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vqa_model_builder_b200.moe import MOELayer
    m = MOELayer(input_dim=64, hidden_dim=128, output_dim=64, num_experts=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4, 64))


def test_sass_uses_blackwell_tensor_path():
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM in the shipped binary."""
    import shutil
    import subprocess
    from vqa_model_builder_b200 import _build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_build.build())], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
