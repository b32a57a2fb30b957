"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/fvla.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols():
    text = (ROOT / "include" / "fvla.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fvla_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for must in ("fvla_create", "fvla_load_tensor", "fvla_finalize", "fvla_forward", "fvla_set_tap", "fvla_last_error",
                 "fvla_op_gemm", "fvla_op_dwconv", "fvla_op_attention"):
        assert must in syms


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    raw = ctypes.CDLL(str(native.lib_path()))
    for s in _header_symbols():
        assert hasattr(raw, s), f"{s} declared in include/fvla.h but not exported by libfvla.so"
    assert sorted(native.EXPORTED_SYMBOLS) == _header_symbols(), "ctypes table and header out of sync"
    assert lib.fvla_abi_version() == 1


def test_no_cpu_fallback(native):
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less machines")
    lib = native.load()
    cfg = native.FvlaConfig()
    h = ctypes.c_void_p()
    assert lib.fvla_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert "no CUDA device" in native.last_error() or "CPU" in native.last_error()
    with pytest.raises(native.NativeError):
        native.require_cuda()
    from vla_fastvlm.model.fastvlm_adapter import FastVLMBackbone, FastVLMBackboneConfig

    with pytest.raises(native.NativeError):
        FastVLMBackbone(FastVLMBackboneConfig(model_id="synthetic:tiny"))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it."""
    for p in (ROOT / "vla-from-fastvlm_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and p.is_file():
            text = p.read_text(errors="ignore")
            assert "oracle." not in text.replace("oracle.fastvla_oracle header", "") or "import oracle" not in text, p
            assert "from oracle" not in text and "import oracle" not in text, p
