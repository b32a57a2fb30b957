"""ctypes binding of libfvla.so (C ABI declared in include/fvla.h).

There is no CPU or eager-PyTorch fallback: if the shared library is missing or no CUDA device is
visible the compute entry points raise.  torch is used only as the owner of device memory and
streams; every call below hands raw pointers + sizes to the extension.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional

import torch

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_lib" / "libfvla.so"
_CSRC = _HERE.parent / "csrc"

FVLA_F32, FVLA_BF16, FVLA_U8 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_SILU, ACT_RELU = 0, 1, 2, 3
ACT_GELU_HALF, ACT_GELU_HALF_F16 = 4, 5  # operand pre-halved; 5 stores the result as fp16
POOL_LAST_TOKEN, POOL_MEAN = 0, 1
IMAGE_TOKEN_INDEX = -200
MAX_VIS_STAGES = 8

TAP_PREPROCESS, TAP_STEM, TAP_VIS_STAGE0 = 0, 1, 10
TAP_IMAGE_FEATURES, TAP_PROJECTOR, TAP_EMBEDS, TAP_LAYER0 = 30, 31, 32, 100
TAP_POOLED, TAP_STATE_FEAT, TAP_FUSED = 1000, 1001, 1002


class NativeError(RuntimeError):
    """Raised when a libfvla call returns a non-zero status."""


class FvlaConfig(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("image_size", C.c_int32),
        ("vis_num_stages", C.c_int32),
        ("vis_layers", C.c_int32 * MAX_VIS_STAGES),
        ("vis_dims", C.c_int32 * MAX_VIS_STAGES),
        ("vis_attention", C.c_int32 * MAX_VIS_STAGES),
        ("vis_pos_emb", C.c_int32 * MAX_VIS_STAGES),
        ("vis_mlp_ratio", C.c_int32),
        ("vis_head_dim", C.c_int32),
        ("vis_se_reduced", C.c_int32),
        ("hidden", C.c_int32),
        ("n_layers", C.c_int32),
        ("n_q_heads", C.c_int32),
        ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32),
        ("intermediate", C.c_int32),
        ("vocab", C.c_int32),
        ("rms_eps", C.c_float),
        ("rope_theta", C.c_float),
        ("state_dim", C.c_int32),
        ("action_dim", C.c_int32),
        ("hidden_dim", C.c_int32),
        ("fusion_dim", C.c_int32),
        ("pool_mode", C.c_int32),
        ("vision_chunk", C.c_int32),
        ("skip_unused_vision", C.c_int32),
    ]


class FvlaForwardArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("images", C.c_void_p),
        ("img_dtype", C.c_int32),
        ("img_nhwc", C.c_int32),
        ("img_c", C.c_int32),
        ("img_h", C.c_int32),
        ("img_w", C.c_int32),
        ("letterbox", C.c_int32),
        ("pad_value", C.c_float),
        ("img_scale", C.c_float),
        ("normalize", C.c_int32),
        ("mean", C.c_float * 3),
        ("inv_std", C.c_float * 3),
        ("token_ids", C.POINTER(C.c_int32)),
        ("text_len", C.POINTER(C.c_int32)),
        ("n_tokens", C.c_int32),
        ("pool_idx", C.POINTER(C.c_int32)),
        ("states", C.c_void_p),
        ("actions", C.c_void_p),
        ("pooled", C.c_void_p),
    ]


# symbol -> (restype, argtypes); mirrors include/fvla.h one to one
_VP, _I32, _I64, _F32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGNATURES = {
    "fvla_abi_version": (C.c_int, []),
    "fvla_last_error": (C.c_char_p, []),
    "fvla_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fvla_create": (C.c_int, [C.POINTER(FvlaConfig), C.POINTER(_VP)]),
    "fvla_destroy": (None, [_VP]),
    "fvla_load_tensor": (C.c_int, [_VP, C.c_char_p, _VP, _I32, _I32, C.POINTER(_I64)]),
    "fvla_missing_tensors": (C.c_int, [_VP, C.c_char_p, _I64, C.POINTER(_I32)]),
    "fvla_finalize": (C.c_int, [_VP]),
    "fvla_reserve": (C.c_int, [_VP, _I32, _I32]),
    "fvla_workspace_bytes": (_I64, [_VP]),
    "fvla_weight_bytes": (_I64, [_VP]),
    "fvla_forward": (C.c_int, [_VP, C.POINTER(FvlaForwardArgs), _VP]),
    "fvla_last_launch_count": (_I64, [_VP]),
    "fvla_last_forward_flops": (C.c_double, [_VP]),
    "fvla_set_tap": (C.c_int, [_VP, _I32, _VP, _I64]),
    "fvla_merged_len": (C.c_int, [_VP]),
    "fvla_set_io_normalization": (C.c_int, [_VP, C.POINTER(_F32), C.POINTER(_F32), C.POINTER(_F32), C.POINTER(_F32)]),
    "fvla_set_profile": (C.c_int, [_VP, _I32]),
    "fvla_profile_report": (C.c_int, [_VP, C.c_char_p, _I64]),
    "fvla_op_gemm": (C.c_int, [_I32, _VP, _I32, _VP, _I32, _VP, _I32, _I32, _I32, _I32, _VP, _VP, _VP,
                               _I32, _I32, _I32, _I32, _VP]),
    "fvla_op_preprocess": (C.c_int, [_I32, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _F32,
                                     _F32, _I32, C.POINTER(_F32), C.POINTER(_F32), _VP, _VP]),
    "fvla_op_stem_conv": (C.c_int, [_I32, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _VP]),
    "fvla_op_dwconv": (C.c_int, [_I32, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                 _I32, _VP]),
    "fvla_op_se_gelu": (C.c_int, [_I32, _VP, _VP, _I32, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP,
                                  _VP]),
    "fvla_op_attention": (C.c_int, [_I32, _I32, _VP, _VP, _VP, _I32, _VP, _I32, _I32, _I32, _I32, _I32,
                                    _I32, _F32, _I32, _VP, _VP, _VP]),
    "fvla_op_rmsnorm": (C.c_int, [_I32, _VP, _VP, _VP, _I32, _I32, _F32, _VP]),
    "fvla_op_layernorm_rows": (C.c_int, [_I32, _VP, _VP, _I32, _I32, _F32, _VP]),
    "fvla_op_ffn_fused": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _VP]),
    "fvla_head_train_scratch_floats": (_I64, [_I32, _I32, _I32, _I32, _I32, _I32]),
    "fvla_head_forward_backward": (C.c_int, [_I32, _I32, _I32, _I32, _I32, _I32, C.POINTER(_VP), _VP, _VP, _VP, _VP,
                                             _F32, _VP, _VP, _VP, _VP, _I64, _VP]),
    "fvla_op_convert": (C.c_int, [_I32, _VP, _I32, _VP, _I64, _VP]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def lib_path() -> Path:
    return _LIB_PATH


def build(verbose: bool = False) -> Path:
    """Compile libfvla.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(_CSRC), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0 or not _LIB_PATH.is_file():
        raise RuntimeError("building libfvla.so failed (see output above)")
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libfvla.so and type every symbol declared in include/fvla.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.is_file():
        raise ImportError(
            f"{_LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C vla-from-fastvlm_b200/csrc`. There is no CPU / eager fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.fvla_abi_version() != 1:
        raise ImportError("libfvla.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().fvla_last_error().decode("utf-8", "replace")


def check(status: int, what: str = "libfvla call") -> None:
    if status != 0:
        raise NativeError(f"{what} failed (status {status}): {last_error()}")


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device visible: the FastVLA B200 path has no CPU fallback")
    n = C.c_int(0)
    check(load().fvla_device_count(C.byref(n)), "fvla_device_count")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return FVLA_F32
    if dt == torch.bfloat16:
        return FVLA_BF16
    if dt == torch.uint8:
        return FVLA_U8
    raise TypeError(f"unsupported dtype for the native path: {dt}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------------------------
# Single-kernel wrappers (tests, micro-benchmarks)
# --------------------------------------------------------------------------------------------
def op_gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
            resid: Optional[torch.Tensor] = None, act: int = ACT_NONE, swiglu: bool = False,
            row_scale: Optional[torch.Tensor] = None, block_n: int = 0,
            out: Optional[torch.Tensor] = None, out_f32: bool = False, split_k: int = 0) -> torch.Tensor:
    """D = act(row_scale * A @ W^T + bias) + resid, A [M,K], W [N,K].

    bf16 operands with an fp32 `out` (or out_f32=True) and fp32 `resid`: the decoder's residual-stream epilogue.
    fp16 operands (A and W both torch.float16) select the tensor-core path with fp16 x fp16 MMAs; the output is
    bf16 unless act == ACT_GELU_HALF_F16 (5), whose result is fp16 (the ConvFFN hidden tensor)."""
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.dtype == w.dtype
    assert a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if swiglu else N
    ab_f16 = a.dtype == torch.float16
    code = dtype_code(torch.bfloat16) if ab_f16 else dtype_code(a.dtype)
    if out is None:
        odt = torch.float16 if act == ACT_GELU_HALF_F16 else (torch.bfloat16 if ab_f16 else a.dtype)
        if out_f32:
            odt = torch.float32
        out = torch.empty((M, n_out), device=a.device, dtype=odt)
    f32_stream = a.dtype != torch.float32 and out.dtype == torch.float32  # bf16 operands, fp32 D / resid
    if f32_stream:
        assert resid is None or resid.dtype == torch.float32
    check(load().fvla_op_gemm(code, ptr(a), K, ptr(w), K, ptr(out), out.stride(0), M, N,
                              K, ptr(bias), ptr(row_scale), ptr(resid),
                              resid.stride(0) if resid is not None else 0, act,
                              int(swiglu) | (2 if ab_f16 else 0) | (4 if f32_stream else 0) | ((split_k & 0xff) << 8),
                              block_n, stream_ptr()), "fvla_op_gemm")
    return out


def op_preprocess(src: torch.Tensor, S: int, out_dtype: torch.dtype, nhwc: bool = False,
                  letterbox: bool = True, pad_value: float = 0.0, scale: float = 1.0,
                  mean=None, std=None) -> torch.Tensor:
    assert src.dim() == 4 and src.is_contiguous()
    if nhwc:
        B, h, w, Cc = src.shape
    else:
        B, Cc, h, w = src.shape
    out = torch.empty((B, S, S, 4), device=src.device, dtype=out_dtype)
    normalize = mean is not None
    m3 = (C.c_float * 3)(*(mean if normalize else (0.0, 0.0, 0.0)))
    s3 = (C.c_float * 3)(*([1.0 / v for v in std] if normalize else (1.0, 1.0, 1.0)))
    check(load().fvla_op_preprocess(dtype_code(out_dtype), ptr(src), dtype_code(src.dtype), int(nhwc), B,
                                    Cc, h, w, S, int(letterbox), pad_value, scale, int(normalize), m3,
                                    s3, ptr(out), stream_ptr()), "fvla_op_preprocess")
    return out


def op_stem_conv(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    B, H, W, c4 = x.shape
    assert c4 == 4 and x.is_contiguous()
    cout = bias.numel()
    out = torch.empty((B, H // 2, W // 2, cout), device=x.device, dtype=x.dtype)
    check(load().fvla_op_stem_conv(dtype_code(x.dtype), ptr(x), ptr(w_packed), ptr(bias), ptr(out), B, H,
                                   W, cout, stream_ptr()), "fvla_op_stem_conv")
    return out


def op_ffn_fused(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                 resid: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = resid + w2 . gelu(w1 . x + b1) + b2 (bf16).  Takes the un-halved fc1 parameters and halves them
    here, the way the engine packs them at finalize (x/2 is exact in bf16)."""
    M, Cc = x.shape
    hidden = w1.shape[0]
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and resid.is_contiguous()
    w1h = (w1.float() * 0.5).to(torch.bfloat16).contiguous()
    b1h = (b1.float() * 0.5).contiguous()
    w2c = w2.to(torch.float16).contiguous()  # the on-chip hidden tensor and fc2 run in fp16
    b2c = b2.float().contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(load().fvla_op_ffn_fused(ptr(x), ptr(w1h), ptr(b1h), ptr(w2c), ptr(b2c), ptr(resid), ptr(out), M, Cc,
                                   hidden, stream_ptr()), "fvla_op_ffn_fused")
    return out


def op_dwconv(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, ksize: int, stride: int,
              mult: int = 1, act: int = ACT_NONE) -> torch.Tensor:
    B, H, W, cin = x.shape
    assert x.is_contiguous()
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty((B, Ho, Wo, cin * mult), device=x.device, dtype=x.dtype)
    check(load().fvla_op_dwconv(dtype_code(x.dtype), ptr(x), ptr(w_packed), ptr(bias), ptr(out), B, H, W,
                                cin, mult, ksize, stride, act, stream_ptr()), "fvla_op_dwconv")
    return out


def op_se_gelu(x: torch.Tensor, w1, b1, w2, b2) -> torch.Tensor:
    B, HW, Cc = x.shape
    out = torch.empty_like(x)
    mean = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    gate = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    check(load().fvla_op_se_gelu(dtype_code(x.dtype), ptr(x), ptr(out), B, HW, Cc, b1.numel(), ptr(w1),
                                 ptr(b1), ptr(w2), ptr(b2), ptr(mean), ptr(gate), stream_ptr()),
          "fvla_op_se_gelu")
    return out


def op_attention(qkv: torch.Tensor, B: int, N: int, heads_q: int, heads_kv: int, head_dim: int,
                 scale: float, causal: bool, rope_cos: Optional[torch.Tensor] = None,
                 rope_sin: Optional[torch.Tensor] = None, impl: int = 0) -> torch.Tensor:
    """qkv: [B*N, (heads_q + 2*heads_kv) * head_dim] fused projection output."""
    assert qkv.dim() == 2 and qkv.is_contiguous()
    ld = qkv.shape[1]
    es = qkv.element_size()
    q = qkv.data_ptr()
    k = q + heads_q * head_dim * es
    v = k + heads_kv * head_dim * es
    out = torch.empty((B * N, heads_q * head_dim), device=qkv.device, dtype=qkv.dtype)
    check(load().fvla_op_attention(dtype_code(qkv.dtype), impl, q, k, v, ld, ptr(out), out.shape[1], B, N,
                                   heads_q, heads_kv, head_dim, scale, int(causal), ptr(rope_cos),
                                   ptr(rope_sin), stream_ptr()), "fvla_op_attention")
    return out


def op_layernorm_rows(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    out = torch.empty_like(x)
    rows, Cc = x.shape
    check(load().fvla_op_layernorm_rows(dtype_code(x.dtype), ptr(x), ptr(out), rows, Cc, eps, stream_ptr()),
          "fvla_op_layernorm_rows")
    return out


def op_rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    out = torch.empty_like(x)
    rows, H = x.shape
    check(load().fvla_op_rmsnorm(dtype_code(x.dtype), ptr(x), ptr(w), ptr(out), rows, H, eps,
                                 stream_ptr()), "fvla_op_rmsnorm")
    return out
