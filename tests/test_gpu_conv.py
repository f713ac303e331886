"""Conv kernels on the B200: the CUDA-core reference conv against torch (fp32), and the
tcgen05 implicit-GEMM tiles against the reference conv on identical bf16 operands."""
import pytest
import torch
import torch.nn.functional as F

from vlnce_b200 import lib
from vlnce_b200.lib import C, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_conv(x, w, scale, shift, res, stride, pad, relu, y_dtype=torch.float32):
    n, h, ww, c = x.shape
    k, r, s, _ = w.shape
    p, q = (h + 2 * pad - r) // stride + 1, (ww + 2 * pad - s) // stride + 1
    y = torch.empty(n, p, q, k, device=DEV, dtype=y_dtype)
    C.vlnce_conv2d_ref(ptr(x), lib.code(x), ptr(w), lib.code(w), ptr(scale), ptr(shift), ptr(res),
                       lib.code(res) if res is not None else 0, ptr(y), lib.code(y), n, h, ww, c, k, r, s,
                       stride, pad, int(relu), lib.stream())
    return y


@pytest.mark.parametrize("cin,k,r,stride,pad,hw", [(3, 16, 7, 2, 3, 20), (8, 24, 3, 1, 1, 9), (16, 8, 1, 2, 0, 10)])
def test_conv_ref_matches_torch(cin, k, r, stride, pad, hw):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(2, hw, hw, cin, generator=g)
    w = torch.randn(k, r, r, cin, generator=g) * 0.2
    scale = torch.rand(k, generator=g) + 0.5
    shift = torch.randn(k, generator=g)
    # CPU fp32 reference (torch-CUDA convolutions default to TF32)
    want = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), stride=stride, padding=pad)
    want = want * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = torch.randn(want.permute(0, 2, 3, 1).shape, generator=g)
    want = torch.relu(want + res.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    got = _ref_conv(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), res.to(DEV), stride, pad, True)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-4)


CASES = [
    # n, h, w, c, k, r, stride, pad, block_n, epilogue
    (2, 16, 16, 64, 64, 1, 1, 0, 0, "affine_relu"),       # plain GEMM, one k-iter
    (2, 16, 16, 128, 64, 1, 1, 0, 64, "raw"),             # two k-iters
    (3, 12, 12, 64, 128, 3, 1, 1, 128, "affine_res_relu"),  # 3x3 pad, M=432 (ragged last tile)
    (2, 16, 16, 64, 256, 3, 2, 1, 256, "affine_relu"),    # stride 2, widest tile
    (2, 16, 16, 256, 512, 1, 2, 0, 0, "affine"),          # 1x1 stride 2 (downsample)
    (4, 8, 8, 32, 32, 3, 1, 1, 32, "raw"),                # BLOCK_K = 32 (64B swizzle), N = 32
    (4, 8, 8, 32, 128, 1, 1, 0, 0, "affine_relu"),        # BLOCK_K = 32
    (1, 8, 8, 512, 512, 3, 1, 1, 128, "affine_res_relu"),  # long K loop (72 iters), ring wraps
    (1, 1, 1, 192, 64, 1, 1, 0, 64, "affine_relu"),       # M = 1 row (stem-GEMM shape, tiny)
    (5, 7, 7, 64, 64, 3, 1, 1, 64, "raw"),                # odd spatial size (224-input shapes)
]


@pytest.mark.parametrize("half", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_umma_matches_reference(case, half):
    n, h, w, c, k, r, stride, pad, bn, epi = case
    g = torch.Generator(device="cpu").manual_seed(sum(case[:9]))
    x = torch.randn(n, h, w, c, generator=g).to(DEV).to(half)
    wt = (torch.randn(k, r, r, c, generator=g) * (2.0 / (r * r * c)) ** 0.5).to(DEV).to(half)
    p, q = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - r) // stride + 1
    scale = shift = res = None
    if "affine" in epi:
        scale = (torch.rand(k, generator=g) + 0.5).to(DEV)
        shift = torch.randn(k, generator=g).to(DEV) * 0.1
    if "res" in epi:
        res = torch.randn(n, p, q, k, generator=g).to(DEV).to(half)
    relu = "relu" in epi
    want = _ref_conv(x, wt, scale, shift, res, stride, pad, relu)  # fp32 out, same bf16 operands
    for y_dtype in (half, torch.float32):
        if res is not None and y_dtype == torch.float32:
            continue  # a residual add always targets the 16-bit activation tensor
        y = torch.full((n, p, q, k), float("nan"), device=DEV, dtype=y_dtype)
        C.vlnce_conv2d_bf16(ptr(x), ptr(wt), lib.code(x), ptr(scale), ptr(shift), ptr(res), ptr(y), lib.code(y),
                            n, h, w, c, k, r, r, stride, pad, int(relu), bn, lib.stream())
        torch.cuda.synchronize()
        tol = {torch.bfloat16: 2e-2, torch.float16: 3e-3, torch.float32: 2e-3}[y_dtype]
        err = (y.float() - want).abs().max().item()
        ref = want.abs().max().item()
        assert torch.isfinite(y.float()).all(), f"non-finite output ({y_dtype})"
        assert err <= tol * max(ref, 1.0), f"{case} {y_dtype}: max err {err} vs ref max {ref}"


def test_conv_umma_rejects_bad_shapes():
    x = torch.zeros(1, 4, 4, 24, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(32, 1, 1, 24, device=DEV, dtype=torch.bfloat16)
    y = torch.zeros(1, 4, 4, 32, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(lib.VlnceError, match="multiple of 32"):
        C.vlnce_conv2d_bf16(ptr(x), ptr(w), lib.BF16, None, None, None, ptr(y), lib.BF16, 1, 4, 4, 24, 32, 1, 1, 1,
                            0, 0, 0, lib.stream())


@pytest.mark.parametrize("hw", [40, 256], ids=["generic", "row-tile-fast-path"])
@pytest.mark.parametrize("img_dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("half", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_fused_rgb_stem_matches_reference(img_dtype, half, hw):
    """7x7/2 pad-3 stem with the im2col row built inside the tile vs the fp32 reference conv
    (hw=256: one tile = one output row, input rows staged in shared memory)."""
    g = torch.Generator().manual_seed(3)
    n, k = (3 if hw < 100 else 2), 64
    img = torch.randint(0, 256, (n, hw, hw, 3), generator=g, dtype=torch.uint8)
    if img_dtype == torch.float32:
        img = img.float()
    w = torch.randn(k, 7, 7, 3, generator=g) * 0.05
    wp = torch.zeros(k, 192)
    wp[:, :147] = w.reshape(k, -1)
    scale, shift = (torch.rand(k, generator=g) + 0.5) / 255.0, torch.randn(k, generator=g) * 0.1
    img_d, w_d, wp_d = img.to(DEV), w.to(DEV).to(half), wp.to(DEV).to(half)
    scale, shift = scale.to(DEV), shift.to(DEV)
    want = _ref_conv(img_d, w_d.float(), scale, shift, None, 2, 3, True)
    p = (hw + 6 - 7) // 2 + 1
    for y_dtype in (half, torch.float32):
        y = torch.full((n, p, p, k), float("nan"), device=DEV, dtype=y_dtype)
        C.vlnce_conv2d_stem_rgb(ptr(img_d), lib.code(img_d), ptr(wp_d), lib.code(wp_d), ptr(scale), ptr(shift), ptr(y),
                                lib.code(y), n, hw, hw, k, 1, 0, lib.stream())
        torch.cuda.synchronize()
        err = (y.float() - want).abs().max().item()
        assert torch.isfinite(y.float()).all()
        assert err <= (2e-2 if y_dtype == torch.bfloat16 else 4e-3) * max(want.abs().max().item(), 1.0), err


HALO_CASES = [
    # n, h, w, c, k, block_n | ring<<16, epilogue
    (2, 16, 16, 64, 64, 0, "affine_relu"),
    (3, 32, 32, 128, 128, 128 | (3 << 16), "affine_res_relu"),
    (2, 16, 8, 256, 256, 256 | (2 << 16), "raw"),
    (1, 64, 64, 64, 64, 64 | (6 << 16), "affine_relu"),
    (2, 32, 16, 128, 256, 128, "affine_res_relu"),
]


@pytest.mark.parametrize("half", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", HALO_CASES, ids=[str(c) for c in HALO_CASES])
def test_conv3x3_halo_matches_reference(case, half):
    """TMA-halo 3x3 kernel (nine shifted descriptors into one shared-memory tile) vs the fp32
    reference conv; image borders exercise the TMA zero fill = padding."""
    n, h, w, c, k, code, epi = case
    g = torch.Generator(device="cpu").manual_seed(sum(case[:5]))
    x = torch.randn(n, h, w, c, generator=g).to(DEV).to(half)
    wt = (torch.randn(k, 3, 3, c, generator=g) * (2.0 / (9 * c)) ** 0.5).to(DEV).to(half)
    scale = shift = res = None
    if "affine" in epi:
        scale = (torch.rand(k, generator=g) + 0.5).to(DEV)
        shift = torch.randn(k, generator=g).to(DEV) * 0.1
    if "res" in epi:
        res = torch.randn(n, h, w, k, generator=g).to(DEV).to(half)
    relu = "relu" in epi
    want = _ref_conv(x, wt, scale, shift, res, 1, 1, relu)
    for y_dtype in (half, torch.float32):
        if res is not None and y_dtype == torch.float32:
            continue
        y = torch.full((n, h, w, k), float("nan"), device=DEV, dtype=y_dtype)
        C.vlnce_conv3x3_halo(ptr(x), ptr(wt), lib.code(x), ptr(scale), ptr(shift), ptr(res), ptr(y), lib.code(y),
                             n, h, w, c, k, int(relu), code, lib.stream())
        torch.cuda.synchronize()
        tol = {torch.bfloat16: 2e-2, torch.float16: 3e-3, torch.float32: 2e-3}[y_dtype]
        err = (y.float() - want).abs().max().item()
        assert torch.isfinite(y.float()).all(), f"non-finite output ({y_dtype})"
        assert err <= tol * max(want.abs().max().item(), 1.0), f"{case} {y_dtype}: max err {err}"


TMA_CASES = [
    # n, h, w, c, k, R, block_n, epilogue   (grids larger than 148 tiles exercise the persistent loop)
    (2, 16, 16, 64, 64, 1, 0, "affine_relu"),
    (3, 12, 12, 128, 256, 1, 128, "affine_res_relu"),      # rows mode, ragged last tile (M = 432)
    (40, 16, 16, 256, 128, 1, 128, "affine_res_relu"),     # 80 m-tiles x 1 = persistent, multi-chunk
    (64, 16, 16, 64, 256, 1, 64, "raw"),                   # 128 m-tiles x 4 n-tiles = 512 tiles on 148 CTAs
    (2, 16, 16, 64, 64, 3, 0, "affine_relu"),
    (3, 32, 32, 128, 128, 3, 128, "affine_res_relu"),
    (24, 32, 16, 64, 128, 3, 64, "affine_res_relu"),       # 192 patches x 2 n-tiles, persistent halo
    (2, 16, 8, 256, 256, 3, 256, "raw"),
    # ragged border tiles (maps of 224-pixel frames): the TMA unit zero-fills the load boxes and clips the store boxes
    (3, 56, 56, 64, 64, 3, 64, "affine_relu"),             # 4 x 7 tiles per image, the last tile row half outside
    (4, 28, 28, 128, 128, 3, 128, "affine_res_relu"),      # ragged in both directions, residual boxes clipped too
    (5, 14, 14, 256, 256, 3, 128, "raw"),                  # one tile row of 16 with 14 valid rows; odd tile count (no pairs)
    (2, 20, 12, 64, 128, 3, 0, "affine"),
]


MT2, NOSTAT = 1 << 21, 1 << 22  # conv_tma block_n flags (include/vlnce_b200.h)


@pytest.mark.parametrize("flags", [0, MT2, NOSTAT, MT2 | NOSTAT], ids=["auto", "mt2", "stream", "mt2-stream"])
@pytest.mark.parametrize("half", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", TMA_CASES, ids=[str(c) for c in TMA_CASES])
def test_conv_tma_persistent_matches_reference(case, half, flags):
    """Persistent TMA-fed kernel (overlapped epilogue, two TMEM accumulator sets) vs the reference conv;
    flags: paired pixel sub-tiles sharing each weight stage, stationary vs streamed weights."""
    n, h, w, c, k, r, code, epi = case
    code |= flags
    g = torch.Generator(device="cpu").manual_seed(sum(case[:6]))
    x = torch.randn(n, h, w, c, generator=g).to(DEV).to(half)
    wt = (torch.randn(k, r, r, c, generator=g) * (2.0 / (r * r * c)) ** 0.5).to(DEV).to(half)
    scale = shift = res = None
    if "affine" in epi:
        scale = (torch.rand(k, generator=g) + 0.5).to(DEV)
        shift = torch.randn(k, generator=g).to(DEV) * 0.1
    if "res" in epi:
        res = torch.randn(n, h, w, k, generator=g).to(DEV).to(half)
    relu = "relu" in epi
    want = _ref_conv(x, wt, scale, shift, res, 1, r // 2, relu)
    for y_dtype in (half, torch.float32):
        if res is not None and y_dtype == torch.float32:
            continue
        y = torch.full((n, h, w, k), float("nan"), device=DEV, dtype=y_dtype)
        C.vlnce_conv_tma(ptr(x), ptr(wt), lib.code(x), ptr(scale), ptr(shift), ptr(res), ptr(y), lib.code(y),
                         n, h, w, c, k, r, int(relu), code, lib.stream())
        torch.cuda.synchronize()
        tol = {torch.bfloat16: 2e-2, torch.float16: 3e-3, torch.float32: 2e-3}[y_dtype]
        err = (y.float() - want).abs().max().item()
        assert torch.isfinite(y.float()).all(), f"non-finite output ({y_dtype})"
        assert err <= tol * max(want.abs().max().item(), 1.0), f"{case} {y_dtype}: max err {err}"


TF32_CASES = [
    # n, h, w, c, k, R, stride, block_n, epilogue   (fp32 operands multiplied as TF32 by the persistent TMA kernel)
    (2, 16, 16, 64, 64, 1, 1, 0, "affine_relu"),
    (3, 12, 12, 96, 256, 1, 1, 128, "affine_res_relu"),     # 32-channel chunks (C = 96), ragged last tile, fp32 residual
    (40, 16, 16, 256, 128, 1, 1, 64, "affine_res_relu"),    # persistent, multi-chunk
    (3, 32, 32, 128, 128, 3, 1, 64, "affine_res_relu"),     # halo mode
    (2, 16, 8, 256, 256, 3, 1, 128, "raw"),
    (4, 32, 32, 64, 128, 1, 2, 0, "affine"),                # strided-box mode, 1x1
    (4, 32, 32, 64, 64, 3, 2, 64, "affine_relu"),           # strided-box mode, 3x3
    (4, 32, 32, 32, 32, 1, 1, 0, "raw"),                    # K = 32 on a 64-wide tile (depth trunk, first stage)
    (4, 32, 32, 32, 32, 3, 1, 0, "raw"),
    (6, 32, 32, 128, 32, 1, 1, 0, "raw"),
    (8, 8, 8, 256, 256, 3, 1, 64, "affine_relu"),           # 8x8 images: two per tile, tap-shifted boxes
    (3, 28, 28, 128, 128, 3, 1, 128, "affine_res_relu"),    # ragged border tiles, fp32 residual boxes
    (2, 14, 14, 64, 64, 3, 1, 0, "raw"),
]


@pytest.mark.parametrize("flags", [0, MT2, NOSTAT], ids=["auto", "mt2", "stream"])
@pytest.mark.parametrize("case", TF32_CASES, ids=[str(c) for c in TF32_CASES])
def test_conv_tma_tf32_matches_reference(case, flags):
    """fp32 storage / TF32 products (tcgen05 kind::tf32) vs the fp32 CUDA-core conv: TF32 rounds both operands to 10
    mantissa bits (relative 2^-11 each), products accumulate in fp32 -> errors ~1e-3 of the output scale."""
    n, h, w, c, k, r, stride, code, epi = case
    g = torch.Generator(device="cpu").manual_seed(sum(case[:7]))
    x = torch.randn(n, h, w, c, generator=g).to(DEV)
    wt = (torch.randn(k, r, r, c, generator=g) * (2.0 / (r * r * c)) ** 0.5).to(DEV)
    scale = shift = res = None
    if "affine" in epi:
        scale = (torch.rand(k, generator=g) + 0.5).to(DEV)
        shift = torch.randn(k, generator=g).to(DEV) * 0.1
    if "res" in epi:
        res = torch.randn(n, h // stride, w // stride, k, generator=g).to(DEV)
    relu = "relu" in epi
    want = _ref_conv(x, wt, scale, shift, res, stride, r // 2, relu)
    y = torch.full((n, h // stride, w // stride, k), float("nan"), device=DEV)
    if stride == 1:
        C.vlnce_conv_tma(ptr(x), ptr(wt), lib.F32, ptr(scale), ptr(shift), ptr(res), ptr(y), lib.F32, n, h, w, c, k, r,
                         int(relu), code | flags, lib.stream())
    else:
        C.vlnce_conv_s2_tma(ptr(x), ptr(wt), lib.F32, ptr(scale), ptr(shift), ptr(y), lib.F32, n, h, w, c, k, r, int(relu),
                            code | flags, lib.stream())
    torch.cuda.synchronize()
    err = (y - want).abs().max().item()
    assert torch.isfinite(y).all(), "non-finite output"
    assert err <= 3e-3 * max(want.abs().max().item(), 1.0), f"{case}: max err {err}"


WARP_CASES = [
    # n, h, w, c, k, R, stride, pad, epilogue   (fp32 warp-cooperative conv: the tf32 mode's fallback geometries)
    (64, 4, 4, 256, 256, 3, 1, 1, "raw"),                   # depth layer4: 4x4 images
    (64, 8, 8, 256, 256, 3, 2, 1, "raw"),
    (6, 8, 8, 512, 1024, 1, 2, 0, "raw"),
    (3, 7, 7, 96, 130, 3, 1, 1, "affine_res_relu"),         # C = 3 x 32, ragged filter and pixel tiles
    (2, 14, 14, 64, 64, 1, 1, 0, "affine_relu"),
    (1, 5, 3, 32, 2, 3, 1, 1, "affine"),                    # fewer filters than a warp's four
]


@pytest.mark.parametrize("case", WARP_CASES, ids=[str(c) for c in WARP_CASES])
def test_conv2d_warp_matches_reference(case):
    """fp32 products and sums on CUDA cores, per-lane partial sums (order differs from the one-thread reference conv)."""
    n, h, w, c, k, r, stride, pad, epi = case
    g = torch.Generator(device="cpu").manual_seed(sum(case[:8]))
    x = torch.randn(n, h, w, c, generator=g).to(DEV)
    wt = (torch.randn(k, r, r, c, generator=g) * (2.0 / (r * r * c)) ** 0.5).to(DEV)
    p, q = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - r) // stride + 1
    scale = shift = res = None
    if "affine" in epi:
        scale = (torch.rand(k, generator=g) + 0.5).to(DEV)
        shift = torch.randn(k, generator=g).to(DEV) * 0.1
    if "res" in epi:
        res = torch.randn(n, p, q, k, generator=g).to(DEV)
    relu = "relu" in epi
    want = _ref_conv(x, wt, scale, shift, res, stride, pad, relu)
    y = torch.full((n, p, q, k), float("nan"), device=DEV)
    C.vlnce_conv2d_warp(ptr(x), ptr(wt), ptr(scale), ptr(shift), ptr(res), ptr(y), n, h, w, c, k, r, r, stride, pad,
                        int(relu), lib.stream())
    torch.cuda.synchronize()
    assert torch.isfinite(y).all(), "non-finite output"
    err = (y - want).abs().max().item()
    assert err <= 2e-5 * max(want.abs().max().item(), 1.0), f"{case}: max err {err}"


S2_CASES = [
    # n, h, w, c, k, block_n, R   (stride-2 convs through the strided-box TMA mode: 1x1 downsample, 3x3 pad 1)
    (2, 32, 32, 64, 128, 0, 1),        # out 16x16: one image per tile, bh = 16
    (64, 64, 64, 256, 512, 128, 1),    # ResNet-50 layer2 downsample at the benchmark batch: 512 tiles x 4 filter tiles
    (4, 16, 16, 128, 256, 256, 1),     # out 8x8: two images per tile
    (16, 8, 16, 64, 64, 64, 1),        # out 4x8: four images per tile
    (6, 32, 16, 1024, 2048, 128, 1),   # long K loop, wide N, non-square
    (2, 32, 32, 64, 64, 0, 3),         # 3x3: nine tap-shifted boxes, zero-filled borders
    (24, 64, 64, 128, 128, 128, 3),    # ResNet-50 layer2 conv2 (stride 2): persistent, 18 K chunks
    (4, 16, 16, 512, 512, 256, 3),     # layer4 conv2: out 8x8, two images per tile, 72 K chunks
    (8, 8, 16, 64, 128, 64, 3),        # out 4x8: four images per tile (taps cross image boundaries -> zeros)
]


@pytest.mark.parametrize("flags", [0, MT2, NOSTAT], ids=["auto", "mt2", "stream"])
@pytest.mark.parametrize("half", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", S2_CASES, ids=[str(c) for c in S2_CASES])
def test_conv_s2_tma_matches_reference(case, half, flags):
    n, h, w, c, k, code, r = case
    g = torch.Generator(device="cpu").manual_seed(sum(case[:5]))
    x = torch.randn(n, h, w, c, generator=g).to(DEV).to(half)
    wt = (torch.randn(k, r, r, c, generator=g) * (2.0 / (r * r * c)) ** 0.5).to(DEV).to(half)
    scale = (torch.rand(k, generator=g) + 0.5).to(DEV)
    shift = torch.randn(k, generator=g).to(DEV) * 0.1
    want = _ref_conv(x, wt, scale, shift, None, 2, r // 2, False)
    for y_dtype in (half, torch.float32):
        y = torch.full((n, h // 2, w // 2, k), float("nan"), device=DEV, dtype=y_dtype)
        C.vlnce_conv_s2_tma(ptr(x), ptr(wt), lib.code(x), ptr(scale), ptr(shift), ptr(y), lib.code(y), n, h, w, c, k, r, 0,
                            code | flags, lib.stream())
        torch.cuda.synchronize()
        tol = {torch.bfloat16: 2e-2, torch.float16: 3e-3, torch.float32: 2e-3}[y_dtype]
        err = (y.float() - want).abs().max().item()
        assert torch.isfinite(y.float()).all(), f"non-finite output ({y_dtype})"
        assert err <= tol * max(want.abs().max().item(), 1.0), f"{case} {y_dtype}: max err {err}"


@pytest.mark.parametrize("flags", [0, MT2, NOSTAT | MT2], ids=["auto", "mt2", "mt2-stream"])
@pytest.mark.parametrize("hw", [40, 224, 256])
@pytest.mark.parametrize("img_dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("half", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_s2d_stem_matches_reference(img_dtype, half, hw, flags):
    """7x7/2 pad-3 stem as a stride-1 4x4 conv over 2x2 space-to-depth cells (overlapping-window TMA
    boxes, no im2col matrix) vs the fp32 reference conv on the image; hw=40/224: ragged row tiles."""
    g = torch.Generator().manual_seed(5)
    n, k = (3 if hw < 100 else 2), 64
    img = torch.randint(0, 256, (n, hw, hw, 3), generator=g, dtype=torch.uint8)
    if img_dtype == torch.float32:
        img = img.float()
    w = torch.randn(k, 7, 7, 3, generator=g) * 0.05
    # cell weights: [k][r'][s'][(dy*2+dx)*3+c] = w[k][2r'+dy][2s'+dx][c], zero where the index is 7
    ws = torch.zeros(k, 4, 4, 16)
    for rp in range(4):
        for sp in range(4):
            for dy in range(2):
                for dx in range(2):
                    if 2 * rp + dy < 7 and 2 * sp + dx < 7:
                        ws[:, rp, sp, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = w[:, 2 * rp + dy, 2 * sp + dx]
    scale, shift = (torch.rand(k, generator=g) + 0.5) / 255.0, torch.randn(k, generator=g) * 0.1
    img_d, w_d, ws_d = img.to(DEV), w.to(DEV).to(half), ws.reshape(k, 256).to(DEV).to(half)
    scale, shift = scale.to(DEV), shift.to(DEV)
    want = _ref_conv(img_d, w_d.float(), scale, shift, None, 2, 3, True)
    p = (hw + 6 - 7) // 2 + 1
    cells = torch.full((n, p + 3, p + 3, 16), float("nan"), device=DEV, dtype=half)
    C.vlnce_stem_s2d(ptr(img_d), lib.code(img_d), ptr(cells), lib.code(cells), n, hw, hw, lib.stream())
    torch.cuda.synchronize()
    # the cell image against a torch gather of the same definition
    pad = torch.zeros(n, 2 * (p + 3), 2 * (p + 3), 3)
    pad[:, 3:3 + hw, 3:3 + hw] = img.float()[:, :min(hw, 2 * (p + 3) - 3), :min(hw, 2 * (p + 3) - 3)]
    want_cells = pad.reshape(n, p + 3, 2, p + 3, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(n, p + 3, p + 3, 12)
    assert torch.equal(cells[..., :12].float().cpu(), want_cells)
    assert (cells[..., 12:].float() == 0).all()
    for y_dtype in (half, torch.float32):
        y = torch.full((n, p, p, k), float("nan"), device=DEV, dtype=y_dtype)
        C.vlnce_conv_stem_tma(ptr(cells), ptr(ws_d), lib.code(ws_d), ptr(scale), ptr(shift), ptr(y), lib.code(y),
                              n, p, p, k, 1, 64 | flags, lib.stream())
        torch.cuda.synchronize()
        err = (y.float() - want).abs().max().item()
        assert torch.isfinite(y.float()).all()
        assert err <= (2e-2 if y_dtype == torch.bfloat16 else 4e-3) * max(want.abs().max().item(), 1.0), err
