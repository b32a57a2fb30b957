"""Per-kernel parity: every sm_100a kernel, called through the C ABI, against plain PyTorch fp32 math
on the same seeded inputs.  Tolerances: fp32 kernels 1e-4 relative-to-scale; bf16 kernels are compared
with an fp32 reference computed from the same bf16-rounded inputs and must agree to bf16 output
rounding (2^-8 relative to the tensor scale)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 1.0 / 128  # output rounding (2^-9) + accumulation-order slack, relative to max|ref|
F32_TOL = 2e-5


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda")


def _close(out, ref, tol, what=""):
    out = out.float()
    ref = ref.float()
    assert out.shape == ref.shape, f"{what}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-12
    err = (out - ref).abs().max().item() / scale
    assert err <= tol, f"{what}: max err {err:.3e} of scale {scale:.3e} exceeds {tol:.1e}"


def _act_ref(x, act):
    if act == 1:
        return F.gelu(x)
    if act == 2:
        return F.silu(x)
    if act == 3:
        return F.relu(x)
    return x


GEMM_SHAPES = [
    # M, N, K                       what it stands for
    (128, 128, 64),                 # single tile
    (256, 256, 128),
    (300, 96, 96),                  # M/N/K tails (stage-0 fc2: K=4*96... here K=96 partial k-block)
    (1000, 192, 384),               # stage-1 fc2 shape, ragged M
    (544, 1152, 896),               # Qwen2-0.5B fused qkv, M = 2 x 272
    (4096, 1536, 384),              # stage-2 fc1
    (1024, 2304, 768),              # stage-3 qkv
    (77, 896, 3072),                # projector, tiny ragged M
    (20000, 384, 96),               # many tiles per CTA (persistent loop, both accumulator stages)
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_resid", "silu_rowscale"])
def test_gemm_bf16(native, M, N, K, epi):
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev) if epi != "plain" else None
    resid = torch.randn(M, N, generator=g).to(dev).bfloat16() if epi == "bias_resid" else None
    rs = (torch.rand(M, generator=g) + 0.5).to(dev) if epi == "silu_rowscale" else None
    act = {"plain": 0, "bias_gelu": 1, "bias_resid": 0, "silu_rowscale": 2}[epi]
    out = native.op_gemm(a, w, bias=bias, resid=resid, act=act, row_scale=rs)
    ref = a.float() @ w.float().t()
    if rs is not None:
        ref = ref * rs[:, None]
    if bias is not None:
        ref = ref + bias
    ref = _act_ref(ref, act)
    if resid is not None:
        ref = ref + resid.float()
    _close(out, ref, BF16_TOL, f"gemm bf16 {M}x{N}x{K} {epi}")


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_gemm_bf16_every_tile_width(native, bn):
    dev = _dev()
    g = torch.Generator().manual_seed(bn)
    M, N, K = 777, 448, 320
    a = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    out = native.op_gemm(a, w, bias=bias, block_n=bn)
    _close(out, a.float() @ w.float().t() + bias, BF16_TOL, f"gemm bn={bn}")


def test_gemm_bf16_inplace_residual(native):
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    M, N, K = 640, 384, 1536
    a = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    x = torch.randn(M, N, generator=g).to(dev).bfloat16()
    ref = a.float() @ w.float().t() + x.float()
    out = native.op_gemm(a, w, resid=x, out=x)
    assert out.data_ptr() == x.data_ptr()
    _close(out, ref, BF16_TOL, "gemm in-place residual")


@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
@pytest.mark.parametrize("M,N,K", [(272, 896, 896), (777, 448, 320), (17408, 896, 4864), (300, 40, 64)])
def test_gemm_bf16_fp32_stream(native, M, N, K, bn):
    """Decoder residual-stream epilogue: bf16 operands, FP32 residual in, FP32 out (in place), every tile width;
    ragged M / N (N = 40: partial 32-column store boxes).  The fp32 result carries no output rounding, so the
    tolerance is the accumulation-order slack only."""
    if M > 4096 and bn not in (0, 256):
        pytest.skip("large case only at the widths the engine picks")
    dev = _dev()
    g = torch.Generator().manual_seed(M + N + K + bn)
    a = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    ref = (a.double() @ w.double().t() + x.double()).float()
    out = native.op_gemm(a, w, resid=x, out=x, block_n=bn)          # as the engine calls it: in place, no bias
    assert out.data_ptr() == x.data_ptr() and out.dtype == torch.float32
    _close(out, ref, 2e-5, f"gemm fp32 stream M={M} N={N} K={K} bn={bn}")
    out2 = native.op_gemm(a, w, bias=bias, out_f32=True, block_n=bn)  # no residual, bias
    _close(out2, (a.double() @ w.double().t() + bias.double()).float(), 2e-5, "gemm fp32 out + bias")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,I,K", [(272, 4864, 896), (100, 256, 128), (1000, 640, 64)])
def test_gemm_swiglu(native, dtype, M, I, K):
    dev = _dev()
    g = torch.Generator().manual_seed(I + K)
    a = torch.randn(M, K, generator=g).to(dev).to(dtype)
    gate = (torch.randn(I, K, generator=g) / math.sqrt(K)).to(dev).to(dtype)
    up = (torch.randn(I, K, generator=g) / math.sqrt(K)).to(dev).to(dtype)
    w = torch.stack([gate, up], dim=1).reshape(2 * I, K).contiguous()  # rows: g0,u0,g1,u1,...
    out = native.op_gemm(a, w, swiglu=True)
    ref = F.silu(a.float() @ gate.float().t()) * (a.float() @ up.float().t())
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, f"swiglu {dtype}")


@pytest.mark.parametrize("M,N,K", [(300, 96, 96), (513, 200, 72), (128, 64, 16), (1000, 1152, 896)])
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_resid", "silu_rowscale"])
def test_gemm_f32(native, M, N, K, epi):
    dev = _dev()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    bias = torch.randn(N, generator=g).to(dev) if epi != "plain" else None
    resid = torch.randn(M, N, generator=g).to(dev) if epi == "bias_resid" else None
    rs = (torch.rand(M, generator=g) + 0.5).to(dev) if epi == "silu_rowscale" else None
    act = {"plain": 0, "bias_gelu": 1, "bias_resid": 0, "silu_rowscale": 2}[epi]
    out = native.op_gemm(a, w, bias=bias, resid=resid, act=act, row_scale=rs)
    ref = a.double() @ w.double().t()
    if rs is not None:
        ref = ref * rs[:, None].double()
    if bias is not None:
        ref = ref + bias.double()
    ref = _act_ref(ref, act)
    if resid is not None:
        ref = ref + resid.double()
    _close(out, ref.float(), F32_TOL, f"gemm f32 {M}x{N}x{K} {epi}")


@pytest.mark.parametrize("M,C,hidden", [(256, 192, 768), (1000, 192, 768), (4096, 96, 384), (37, 96, 128),
                                        (20000, 192, 256), (33000, 96, 384)])
def test_ffn_fused(native, M, C, hidden):
    """fc1 -> GELU -> fc2 -> +residual with the hidden tensor kept on chip vs the same math in fp64
    (bf16 operands, hidden rounded to bf16 as the kernel stores it)."""
    dev = _dev()
    g = torch.Generator().manual_seed(M + C + hidden)
    x = torch.randn(M, C, generator=g).to(dev).bfloat16()
    w1 = (torch.randn(hidden, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
    b1 = (0.5 * torch.randn(hidden, generator=g)).to(dev)
    w2 = (torch.randn(C, hidden, generator=g) / math.sqrt(hidden)).to(dev).bfloat16()
    b2 = torch.randn(C, generator=g).to(dev)
    resid = torch.randn(M, C, generator=g).to(dev).bfloat16()
    out = native.op_ffn_fused(x, w1, b1, w2, b2, resid)
    h = F.gelu(x.double() @ w1.double().t() + b1.double()).bfloat16().double()
    ref = (h @ w2.double().t() + b2.double() + resid.double()).float()
    _close(out, ref, BF16_TOL, f"ffn_fused M={M} C={C} hidden={hidden}")
    # in place on the residual buffer (how the engine calls it)
    buf = resid.clone()
    native.op_ffn_fused(x, w1, b1, w2, b2, buf, out=buf)
    assert torch.equal(buf, out), "in-place residual differs"


def _pack_dw(w):  # [Cout,1,k,k] -> [k*k][Cout]
    cout, _, k, _ = w.shape
    return w.reshape(cout, k * k).t().contiguous()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k,stride,mult,act", [(3, 1, 1, 0), (3, 2, 1, 1), (7, 1, 1, 0), (7, 2, 2, 1),
                                               (3, 1, 2, 0), (7, 2, 1, 0)])
@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 16), (1, 37, 29, 24), (3, 64, 64, 96), (2, 8, 128, 32),
                                     (1, 128, 128, 192)])
def test_dwconv(native, dtype, k, stride, mult, act, B, H, W, C):
    dev = _dev()
    if stride == 2 and (H % 2 or W % 2):
        H, W = H + H % 2, W + W % 2
    g = torch.Generator().manual_seed(k * 100 + stride * 10 + mult + C)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    w = (torch.randn(C * mult, 1, k, k, generator=g) / k).to(dev)
    b = torch.randn(C * mult, generator=g).to(dev)
    xin = x.permute(0, 2, 3, 1).contiguous().to(dtype)
    out = native.op_dwconv(xin, _pack_dw(w), b, k, stride, mult, act)
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w, b, stride=stride, padding=k // 2, groups=C)
    ref = _act_ref(ref, act).permute(0, 2, 3, 1)
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, f"dwconv k{k}s{stride}m{mult}")


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 64), (1, 32, 32, 96), (2, 48, 64, 32), (1, 256, 256, 96),
                                     (3, 16, 32, 160)])
def test_dwconv7_tensor_core(native, B, H, W, C):
    """Tensor-core (Toeplitz mma) depthwise 7x7: every tile geometry, image borders inside and between tiles,
    per-channel distinct taps; checked against conv2d with the taps it actually multiplies (bf16)."""
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + W + C)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev)
    b = torch.randn(C, generator=g).to(dev)
    xin = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    out = native.op_dwconv(xin, _pack_dw(w), b, 7, 1, 1, 0)
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=3, groups=C).permute(0, 2, 3, 1)
    _close(out, ref, 1.0 / 256, "dwconv7 tensor core (bf16 taps)")
    ref32 = F.conv2d(xin.float().permute(0, 3, 1, 2), w, b, padding=3, groups=C).permute(0, 2, 3, 1)
    _close(out, ref32, BF16_TOL, "dwconv7 tensor core (fp32 taps)")
    # an impulse image reads the taps back exactly (bf16-rounded), flipped, at the right channel
    imp = torch.zeros(1, H, W, C, device=dev, dtype=torch.bfloat16)
    imp[0, H // 2, W // 2, :] = 1.0
    o = native.op_dwconv(imp, _pack_dw(w), torch.zeros(C, device=dev), 7, 1, 1, 0).float()
    patch = o[0, H // 2 - 3:H // 2 + 4, W // 2 - 3:W // 2 + 4, :]  # [7,7,C]
    want = w.bfloat16().float()[:, 0].flip(1, 2).permute(1, 2, 0)
    assert torch.equal(patch, want), "impulse response != taps"


@pytest.mark.parametrize("B,H,W,C", [(1, 8, 32, 32), (2, 16, 64, 96), (3, 40, 96, 64), (5, 128, 128, 192),
                                     (8, 256, 256, 96), (2, 64, 64, 384), (2, 24, 48, 128), (1, 8, 16, 64)])
def test_dwconv3_tma(native, B, H, W, C, monkeypatch):
    """Persistent TMA-fed depthwise 3x3 (fp32 FFMA2 taps): single-tile and many-tiles-per-CTA grids (the 3-slot
    ring wraps and its barrier parity flips), borders inside / between tiles and between images, per-channel
    distinct taps.  The kernel accumulates in fp32, so each output is the fp32 conv rounded once to bf16."""
    dev = _dev()
    monkeypatch.delenv("FVLA_DISABLE_DWCONV3_TMA", raising=False)
    g = torch.Generator().manual_seed(B * 977 + H + W + C)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(dev)
    b = torch.randn(C, generator=g).to(dev)
    xin = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    out = native.op_dwconv(xin, _pack_dw(w), b, 3, 1, 1, 0).float()
    ref = F.conv2d(xin.double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1, groups=C).permute(0, 2, 3, 1)
    err = (out.double() - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + 1e-5  # one bf16 rounding (half an ulp <= 2^-9 relative) + fp32 summation slack
    assert bool((err <= bound).all()), f"dwconv3: worst excess {(err - bound).max().item():.3e}"
    # an impulse reads the taps back (flipped), bf16-rounded once
    imp = torch.zeros(1, H, W, C, device=dev, dtype=torch.bfloat16)
    imp[0, H // 2, W // 2, :] = 1.0
    o = native.op_dwconv(imp, _pack_dw(w), torch.zeros(C, device=dev), 3, 1, 1, 0).float()
    patch = o[0, H // 2 - 1:H // 2 + 2, W // 2 - 1:W // 2 + 2, :]
    want = w.bfloat16().float()[:, 0].flip(1, 2).permute(1, 2, 0)
    assert torch.equal(patch, want), "impulse response != taps"
    o[0, H // 2 - 1:H // 2 + 2, W // 2 - 1:W // 2 + 2, :] = 0
    assert not bool(o.any()), "impulse leaked outside its 3x3 neighbourhood"


def test_dwconv3_one_and_two_row_kernels_agree(native, monkeypatch):
    """The two-rows-per-thread 3x3 kernel (default) and the one-row kernel sum the same nine fp32 products in the same
    order: their outputs are bit-identical (many tiles per CTA, several images, every channel-block count)."""
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    for (B, H, W, C) in [(3, 64, 64, 96), (2, 128, 128, 192), (1, 40, 96, 64)]:
        x = torch.randn(B, H, W, C, generator=g).to(dev).bfloat16()
        w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(dev)
        b = torch.randn(C, generator=g).to(dev)
        monkeypatch.delenv("FVLA_DWCONV3_ONE_ROW", raising=False)
        two = native.op_dwconv(x, _pack_dw(w), b, 3, 1, 1, 0)
        monkeypatch.setenv("FVLA_DWCONV3_ONE_ROW", "1")
        one = native.op_dwconv(x, _pack_dw(w), b, 3, 1, 1, 0)
        monkeypatch.delenv("FVLA_DWCONV3_ONE_ROW", raising=False)
        assert torch.equal(one, two), (B, H, W, C)


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 16), (1, 64, 96, 48), (3, 128, 128, 96), (1, 256, 256, 96),
                                     (2, 32, 64, 768)])
def test_dwconv7_stride2_tensor_core(native, B, H, W, C, act):
    """PatchEmbed depthwise 7x7, stride 2, two output channels per input channel (+ GELU) on the tensor-core pipeline
    (stride-2 Toeplitz bands, k16 + k8 steps): single- and multi-tile grids, borders inside / between tiles and images,
    per-channel distinct taps for both outputs of a channel; against conv2d with the taps it multiplies (bf16)."""
    dev = _dev()
    g = torch.Generator().manual_seed(B * 131 + H + 3 * W + C + act)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    w = (torch.randn(2 * C, 1, 7, 7, generator=g) / 7).to(dev)
    b = torch.randn(2 * C, generator=g).to(dev)
    xin = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    out = native.op_dwconv(xin, _pack_dw(w), b, 7, 2, 2, act)
    assert out.shape == (B, H // 2, W // 2, 2 * C)

    def ref_of(weights):
        r = F.conv2d(xin.float().permute(0, 3, 1, 2), weights, b, stride=2, padding=3, groups=C)
        return _act_ref(r, act).permute(0, 2, 3, 1)

    _close(out, ref_of(w.bfloat16().float()), 1.0 / 128 if act == 0 else BF16_TOL, "dwconv7 s2 tensor core (bf16 taps)")
    _close(out, ref_of(w), BF16_TOL, "dwconv7 s2 tensor core (fp32 taps)")
    if act == 0:
        # impulses at an even and an odd position read back every second tap of both output channels, exactly
        for (py, px) in [(H // 2, W // 2), (H // 2 + 1, W // 2 - 1)]:
            imp = torch.zeros(1, H, W, C, device=dev, dtype=torch.bfloat16)
            imp[0, py, px, :] = 1.0
            o = native.op_dwconv(imp, _pack_dw(w), torch.zeros(2 * C, device=dev), 7, 2, 2, 0).float()
            want = F.conv2d(imp.float().permute(0, 3, 1, 2), w.bfloat16().float(), None, stride=2, padding=3,
                            groups=C).permute(0, 2, 3, 1)
            assert torch.equal(o, want), "stride-2 impulse response != taps"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_conv(native, dtype):
    dev = _dev()
    g = torch.Generator().manual_seed(11)
    B, H, W, Cout = 2, 64, 48, 96
    x = torch.rand(B, 3, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, 3, 3, 3, generator=g) / 3).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    x4 = torch.zeros(B, H, W, 4, device=dev)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    x4 = x4.to(dtype)
    wp = w.permute(2, 3, 1, 0).reshape(27, Cout).contiguous()  # row = (ky*3+kx)*3+ci
    out = native.op_stem_conv(x4, wp, b)
    ref = F.gelu(F.conv2d(x4[..., :3].float().permute(0, 3, 1, 2), w, b, stride=2, padding=1)).permute(0, 2, 3, 1)
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, "stem conv")


def _resize_with_pad_ref(img, S, pad_value=0.0):
    # restates fastvlm_adapter.py:36-55
    h, w = img.shape[2:]
    ratio = max(w / S, h / S)
    rh, rw = int(h / ratio), int(w / ratio)
    r = F.interpolate(img, size=(rh, rw), mode="bilinear", align_corners=False)
    return F.pad(r, (max(0, S - rw), 0, max(0, S - rh), 0), value=pad_value)


@pytest.mark.parametrize("shape,S", [((2, 3, 480, 480), 1024), ((1, 3, 480, 640), 1024), ((2, 3, 64, 64), 64),
                                     ((1, 1, 50, 70), 128), ((1, 4, 300, 200), 256), ((1, 3, 1500, 1100), 512)])
@pytest.mark.parametrize("nhwc", [False, True])
def test_preprocess_letterbox(native, shape, S, nhwc):
    dev = _dev()
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(*shape, generator=g).to(dev)
    xc = x.repeat(1, 3, 1, 1) if shape[1] == 1 else x[:, :3]
    ref = _resize_with_pad_ref(xc, S, 0.25).permute(0, 2, 3, 1)
    src = x.permute(0, 2, 3, 1).contiguous() if nhwc else x
    out = native.op_preprocess(src, S, torch.float32, nhwc=nhwc, letterbox=True, pad_value=0.25)
    _close(out[..., :3], ref, 1e-5, f"preprocess {shape}->{S}")
    assert (out[..., 3] == 0).all()


def test_preprocess_uint8_stretch_and_normalize(native):
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (2, 3, 40, 60), generator=g, dtype=torch.uint8).to(dev)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out = native.op_preprocess(x, 96, torch.float32, letterbox=False, scale=1 / 255.0, mean=mean, std=std)
    ref = F.interpolate(x.float(), size=(96, 96), mode="bilinear", align_corners=False) / 255.0
    ref = (ref - torch.tensor(mean, device=dev).view(1, 3, 1, 1)) / torch.tensor(std, device=dev).view(1, 3, 1, 1)
    _close(out[..., :3], ref.permute(0, 2, 3, 1), 1e-5, "preprocess u8")


def _rope_tables(T, hd, theta, dev):
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    ang = torch.arange(T, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().to(dev).contiguous(), ang.sin().to(dev).contiguous()


def _attn_ref(qkv, B, N, hq, hkv, hd, scale, causal, cos=None, sin=None):
    qkv = qkv.float().view(B, N, -1)
    q = qkv[..., : hq * hd].view(B, N, hq, hd).transpose(1, 2)
    k = qkv[..., hq * hd: (hq + hkv) * hd].view(B, N, hkv, hd).transpose(1, 2)
    v = qkv[..., (hq + hkv) * hd:].view(B, N, hkv, hd).transpose(1, 2)
    if cos is not None:
        c = torch.cat([cos, cos], -1)[None, None]
        s = torch.cat([sin, sin], -1)[None, None]

        def rot(x):
            x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
            return x * c + torch.cat([-x2, x1], -1) * s

        q, k = rot(q), rot(k)
    k = k.repeat_interleave(hq // hkv, dim=1)
    v = v.repeat_interleave(hq // hkv, dim=1)
    att = (q @ k.transpose(-1, -2)) * scale
    if causal:
        att = att.masked_fill(torch.ones(N, N, device=att.device, dtype=torch.bool).triu(1), float("-inf"))
    o = att.softmax(-1) @ v
    return o.transpose(1, 2).reshape(B * N, hq * hd)


ATTN_CASES = [
    # B, N, hq, hkv, hd, causal, rope
    (2, 1024, 24, 24, 32, False, False),   # FastViTHD stage 3
    (3, 256, 48, 48, 32, False, False),    # FastViTHD stage 4
    (2, 272, 14, 2, 64, True, True),       # Qwen2-0.5B prefill, T' = 256 + 16
    (2, 100, 12, 2, 128, True, True),      # Qwen2-1.5B / 7B head_dim
    (1, 77, 4, 4, 32, False, False),       # ragged N
    (2, 19, 2, 1, 64, True, True),         # tiny test-model shape
    (2, 272, 14, 2, 64, True, False),      # Qwen2-0.5B after in-place RoPE (v2 kernel, causal, ragged tail)
    (1, 300, 12, 2, 128, True, False),     # head_dim 128, v2 kernel
    (2, 130, 4, 4, 64, False, False),      # v2 kernel, non-causal with a 2-key tail tile
    (1, 600, 4, 2, 64, True, False),       # v2 kernel, causal (N >= 512)
    (3, 33, 28, 4, 128, True, False),      # grouped-query kernel: Qwen2-7B grouping (7 heads per kv head), ragged N
    (2, 290, 8, 1, 64, True, False),       # grouped-query kernel: largest group (8 warps), 5 kv tiles
    (2, 64, 6, 3, 64, True, False),        # grouped-query kernel: group of 2, N a multiple of the tile
    (2, 1024, 24, 24, 32, False, False),   # FastViTHD stage-3 MHSA: tcgen05 / tensor-memory kernel (attention_sm100.cu)
    (3, 256, 48, 48, 32, False, False),    # FastViTHD stage-4 MHSA: same kernel, 4 key tiles
    (1, 384, 6, 2, 32, False, False),      # same kernel, grouped kv heads, odd tile count
]


@pytest.mark.parametrize("B,N,hq,hkv,hd,causal,rope", ATTN_CASES)
@pytest.mark.parametrize("impl,dtype", [(0, torch.bfloat16), (1, torch.bfloat16), (0, torch.float32)])
def test_attention(native, B, N, hq, hkv, hd, causal, rope, impl, dtype):
    dev = _dev()
    g = torch.Generator().manual_seed(N + hd)
    qkv = torch.randn(B * N, (hq + 2 * hkv) * hd, generator=g).to(dev).to(dtype)
    cos = sin = None
    if rope:
        cos, sin = _rope_tables(N, hd, 1e6, dev)
    scale = hd ** -0.5
    out = native.op_attention(qkv, B, N, hq, hkv, hd, scale, causal, cos, sin, impl=impl)
    ref = _attn_ref(qkv, B, N, hq, hkv, hd, scale, causal, cos, sin)
    # flash path rounds P and the rotated q/k to bf16: allow 2x the plain bf16 tolerance
    tol = 2 * BF16_TOL if dtype == torch.bfloat16 else 1e-4
    _close(out, ref, tol, f"attention impl={impl} {dtype}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H", [896, 1536, 3584, 5120, 64])  # 5120: beyond the register-resident part of a row
def test_rmsnorm(native, dtype, H):
    dev = _dev()
    g = torch.Generator().manual_seed(9 + H)
    x = (torch.randn(333, H, generator=g) * 3).to(dev).to(dtype)
    w = torch.randn(H, generator=g).to(dev)
    out = native.op_rmsnorm(x, w, 1e-6)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, "rmsnorm")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Cc", [768, 1536, 128, 3072])
def test_layernorm_rows(native, dtype, Cc):
    """LayerNormChannel statistics (FastViTHD AttentionBlock.norm, affine folded elsewhere) vs torch layer_norm."""
    dev = _dev()
    g = torch.Generator().manual_seed(5 + Cc)
    x = (torch.randn(517, Cc, generator=g) * 2 + 0.7).to(dev).to(dtype)
    out = native.op_layernorm_rows(x, 1e-5)
    ref = torch.nn.functional.layer_norm(x.float(), (Cc,), None, None, 1e-5)
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, "layernorm_rows")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,HW,Cc,Cr", [(3, 256, 512, 32), (2, 256, 3072, 192), (9, 64, 104, 8)])
def test_se_gelu(native, dtype, B, HW, Cc, Cr):
    dev = _dev()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, HW, Cc, generator=g).to(dev).to(dtype)
    w1 = (torch.randn(Cr, Cc, generator=g) / math.sqrt(Cc)).to(dev)
    b1 = torch.randn(Cr, generator=g).to(dev)
    w2 = (torch.randn(Cc, Cr, generator=g) / math.sqrt(Cr)).to(dev)
    b2 = torch.randn(Cc, generator=g).to(dev)
    out = native.op_se_gelu(x, w1, b1, w2, b2)
    xf = x.float()
    gate = torch.sigmoid(F.relu(xf.mean(1) @ w1.t() + b1) @ w2.t() + b2)
    ref = F.gelu(xf * gate[:, None, :])
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, "se_gelu")


@pytest.mark.parametrize("M,K,hidden", [(640, 384, 1536), (300, 96, 384), (4096, 768, 3072)])
def test_gemm_fp16_hidden_chain(native, M, K, hidden):
    """ConvFFN as two GEMMs with the hidden tensor in fp16: fc1 (bf16 operands, pre-halved, act 5 -> fp16 GELU
    output) then fc2 (fp16 x fp16 MMA, bf16 output + residual), against fp64 math."""
    dev = _dev()
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w1 = (torch.randn(hidden, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    b1 = (0.5 * torch.randn(hidden, generator=g)).to(dev)
    w2 = (torch.randn(K, hidden, generator=g) / math.sqrt(hidden)).to(dev).bfloat16()
    b2 = torch.randn(K, generator=g).to(dev)
    resid = torch.randn(M, K, generator=g).to(dev).bfloat16()
    h = native.op_gemm(x, (w1.float() * 0.5).bfloat16(), bias=b1 * 0.5, act=5)
    assert h.dtype == torch.float16
    href = F.gelu(x.double() @ w1.double().t() + b1.double())
    _close(h, href.float(), 1.0 / 512, "fp16 GELU hidden")
    # large-magnitude pre-activations stay exact in the saturated regions
    out = native.op_gemm(h, w2.half(), bias=b2, resid=resid)
    assert out.dtype == torch.bfloat16
    ref = (h.double() @ w2.double().t() + b2.double() + resid.double()).float()
    _close(out, ref, BF16_TOL, "fc2 on fp16 operands")


def test_gemm_bf16_gelu_large_magnitudes(native):
    """The MUFU-based GELU must stay exact-in-bf16 far outside the fitted range (|x| up to ~300)."""
    dev = _dev()
    g = torch.Generator().manual_seed(77)
    M, N, K = 512, 256, 64
    a = (torch.randn(M, K, generator=g) * 12).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.4).to(dev).bfloat16()
    out = native.op_gemm(a, w, act=1)
    ref = F.gelu(a.float() @ w.float().t())
    assert ref.abs().max() > 100
    _close(out, ref, BF16_TOL, "gelu large")
    pos = ref > 20
    assert torch.allclose(out.float()[pos], ref[pos], rtol=1e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_gelu_half_operand(native, dtype):
    """ACT_GELU_HALF (4): weights/bias pre-scaled by 1/2, epilogue returns gelu(2h)."""
    dev = _dev()
    g = torch.Generator().manual_seed(31)
    M, N, K = 700, 384, 96
    a = torch.randn(M, K, generator=g).to(dev).to(dtype)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K) * 2).to(dev).to(dtype)
    bias = torch.randn(N, generator=g).to(dev)
    out = native.op_gemm(a, (w * 0.5).contiguous(), bias=bias * 0.5, act=4)
    ref = F.gelu(a.float() @ w.float().t() + bias)
    _close(out, ref, BF16_TOL if dtype == torch.bfloat16 else F32_TOL, "gelu half")
