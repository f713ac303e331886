"""Launch plans for the two frozen conv trunks of the policy, driving the C-ABI kernels.

* RGB: torchvision ResNet-50/18 v1.5 with BatchNorm (``resnet_encoders.py:118-229``):
  eval-mode BN is folded into the conv epilogue (scale/shift + residual + ReLU); train-mode
  BN (the constructor default) runs conv -> batch statistics -> fused apply and
  updates the running statistics like ``torch.nn.BatchNorm2d``.
* Depth: DD-PPO GroupNorm ResNet-50 + 3x3 compression (``resnet_encoders.py:17-115``;
  habitat_baselines v0.1.7 ``ResNetEncoder``): conv (raw) -> GroupNorm statistics -> fused
  normalise(+residual)(+ReLU).

Activations are NHWC (the simulator's layout, so no permute) in bf16 (``precision="bf16"``,
tcgen05 tensor-core tiles) or fp32 (``precision="fp32"``, CUDA-core reference conv — the
1e-3 parity mode).  Weights are re-packed (KRSC, folded norms) whenever the parameters'
version counters change.  Python only sequences launches; all arithmetic is in the library.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import lib
from .lib import C, ptr


class BufferPool:
    """Named persistent device buffers (static addresses -> CUDA-graph friendly)."""

    def __init__(self, device):
        self.device = device
        self._bufs: Dict[str, torch.Tensor] = {}
        self._views: Dict[tuple, torch.Tensor] = {}  # (name, shape, dtype) -> the typed view handed out last time

    def get(self, name: str, shape, dtype) -> torch.Tensor:
        # the launch plans ask for the same (name, shape, dtype) every step: hand the same view object back instead of
        # re-slicing / re-viewing the byte buffer (three tensor ops, ~10 us of host time per request on eager paths)
        key = (name, tuple(shape), dtype)
        v = self._views.get(key)
        if v is not None:
            return v
        numel = 1
        for s in shape:
            numel *= int(s)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        b = self._bufs.get(name)
        if b is None or b.numel() < nbytes:
            # host memory only exists in the CPU dry-run walks of the launch plans (tests): keep those deterministic
            alloc = torch.zeros if torch.device(self.device).type == "cpu" else torch.empty
            b = alloc(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._bufs[name] = b
            for k in [k for k in self._views if k[0] == name]:  # views of the buffer this one replaces
                del self._views[k]
        v = b[:nbytes].view(dtype).view(*shape)
        self._views[key] = v
        return v

    def nbytes(self) -> int:
        return sum(b.numel() for b in self._bufs.values())


import os as _os

AUTOTUNE = _os.environ.get("VLNCE_AUTOTUNE", "1") != "0"
FUSED_STEM = _os.environ.get("VLNCE_FUSED_STEM", "1") != "0"
HALO_3X3 = _os.environ.get("VLNCE_HALO_3X3", "1") != "0"
PERSISTENT = _os.environ.get("VLNCE_PERSISTENT", "1") != "0"
S2D_STEM = _os.environ.get("VLNCE_S2D_STEM", "1") != "0"
MT2 = 1 << 21     # conv_tma flag: two pixel sub-tiles share each weight stage
NOSTAT = 1 << 22  # conv_tma flag: do not keep the weight slice stationary in shared memory
HALO = 1 << 24  # tuning-code flag: TMA-halo 3x3 kernel instead of the generic tile
TMAP = 1 << 25  # tuning-code flag: persistent TMA-fed kernel (1x1 rows / 3x3 halo, overlapped epilogue)
S2TMA = 1 << 26  # tuning-code flag: stride-2 conv (1x1 / 3x3) through the strided-box TMA mode of the persistent kernel
EXCL = 1 << 20  # launch flag understood by the conv entry points: one CTA per SM
# tile width of the tf32 mode where the filter count allows it (else the library default 64); measured on the B200 at 64
# rows (RGB trunk): 64 -> 5.41 ms, 128 -> 3.72 ms, 256 -> 4.12 ms, 64 / 128 with two pixel sub-tiles -> 4.04 / 3.88 ms
TF32_BLOCK_N = int(_os.environ.get("VLNCE_TF32_BLOCK_N", "128"))
RGB_EXCLUSIVE = _os.environ.get("VLNCE_RGB_EXCLUSIVE", "0") != "0"
RAGGED_TMA = _os.environ.get("VLNCE_RAGGED_TMA", "1") != "0"  # persistent TMA 3x3 tiles on maps that are not multiples of 16x8
# fp16 instead of fp32 raw conv output under GroupNorm (depth trunk): half the bytes of the statistics/apply passes
GN_RAW16 = _os.environ.get("VLNCE_GN_RAW16", "0") != "0"
DEPTH_CHUNKS = int(_os.environ.get("VLNCE_DEPTH_CHUNKS", "1"))  # independent batch chunks of the depth trunk
# image-stationary depth trunk: the whole GroupNorm ResNet in one launch (csrc/depth_fused.cu); 0 = per-layer launches
DEPTH_FUSED = _os.environ.get("VLNCE_DEPTH_FUSED", "1") != "0"
DEPTH_FUSED_CTAS = int(_os.environ.get("VLNCE_DEPTH_FUSED_CTAS", "0"))  # persistent grid (0 = one CTA per image)
# Tuned (tile width | ring depth << 16) per conv geometry, optionally persisted so that later
# processes on the same box (profilers, the second bench leg) skip the timing sweep.
_TUNE_CACHE_PATH = _os.environ.get("VLNCE_AUTOTUNE_CACHE", "")
_TUNE_CACHE = {}
# Committed tile-selection table for the B200 (kernel variant | tile width | ring depth per conv geometry, measured
# by the sweep below on this pool's B200s: tools/make_tile_table.py): the autotuner only runs on a MISS, so first use
# of the shipped geometries costs no timing launches and every process picks the same tiles.  VLNCE_TILE_TABLE=0
# ignores it (re-measure everything).
_TILE_TABLE_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "tile_table_b200.json")
if _os.environ.get("VLNCE_TILE_TABLE", "1") != "0" and _os.path.exists(_TILE_TABLE_PATH):
    try:
        import json as _json

        _TUNE_CACHE.update(_json.load(open(_TILE_TABLE_PATH)))
    except Exception:
        pass
if _TUNE_CACHE_PATH and _os.path.exists(_TUNE_CACHE_PATH):
    try:
        import json as _json

        _TUNE_CACHE.update(_json.load(open(_TUNE_CACHE_PATH)))
    except Exception:
        pass


def _batch_bucket(n: int) -> int:
    """Tile selection is keyed by the batch size rounded down to a power of two and capped at 64: above that every
    layer already fills the 148 SMs and the best tile no longer depends on the image count (a training loop with a
    varying number of rows - RecollectTrainer's padded T x N - must not re-tune per batch)."""
    b = 1
    while b * 2 <= min(int(n), 64):
        b *= 2
    return b


def _tune_cache_put(key: str, val: int):
    _TUNE_CACHE[key] = val
    if _TUNE_CACHE_PATH:
        import json as _json

        tmp = _TUNE_CACHE_PATH + f".{_os.getpid()}.tmp"
        with open(tmp, "w") as f:
            _json.dump(_TUNE_CACHE, f)
        _os.replace(tmp, _TUNE_CACHE_PATH)


def _version(params) -> tuple:
    return tuple((p.data_ptr(), p._version) for p in params)


class _Conv:
    """One packed convolution (+ its normalisation)."""

    def __init__(self, conv: nn.Conv2d, norm: nn.Module, precision: str, in_scale: float = 1.0,
                 stem: Optional[str] = None, wdt=torch.bfloat16, in_chan_scale=None):
        self.conv, self.norm, self.precision, self.in_scale, self.stem = conv, norm, precision, in_scale, stem
        self.in_chan_scale = in_chan_scale  # per-input-channel factor folded into the packed weights (1/std)
        self.wdt = wdt if precision == "bf16" else torch.float32
        self.K, self.Cin, self.R, self.S = conv.weight.shape
        self.stride, self.pad = conv.stride[0], conv.padding[0]
        self.w = self.scale = self.shift = None
        self._ver = None
        self.block_n = {}  # tuned tile width per problem shape (0 = library heuristic)

    def params(self):
        ps = [self.conv.weight]
        ps += [p for p in self.norm.parameters()] + [b for b in self.norm.buffers()]
        return ps

    def pack(self):
        ver = _version(self.params())
        if ver == self._ver:
            return
        w = self.conv.weight.detach()
        if self.in_chan_scale is not None:
            w = w * torch.tensor(self.in_chan_scale, device=w.device, dtype=w.dtype).view(1, -1, 1, 1)
        wdt = self.wdt
        krsc = w.permute(0, 2, 3, 1).contiguous()
        if self.precision == "bf16" and self.stem == "rgb":  # im2col operand: (r*7+s)*3+c, pad 147->192
            flat = krsc.reshape(self.K, -1)
            krsc = torch.zeros(self.K, 192, device=w.device, dtype=w.dtype)
            krsc[:, : flat.shape[1]] = flat
        elif self.precision in ("bf16", "tf32") and self.stem == "depth":  # r*7+s, pad 49->64 (tf32 mode: fp16 operand,
            # 11-bit significand >= TF32's 10: the depth stem runs as the same im2col GEMM in both tensor-core modes)
            flat = krsc.reshape(self.K, -1)
            krsc = torch.zeros(self.K, 64, device=w.device, dtype=w.dtype)
            krsc[:, : flat.shape[1]] = flat
        self.w = krsc.to(torch.float16 if (self.precision == "tf32" and self.stem == "depth") else wdt).contiguous()
        if self.precision == "tf32" and self.stem != "depth":
            # kind::tf32 reads the upper 19 bits of each fp32 operand: round the packed weights to nearest TF32 once
            # (the conv epilogue does the same for the activations it writes), instead of letting the MMA truncate
            bits = self.w.view(torch.int32)
            self.w = ((bits + 0x1000) & ~0x1FFF).view(torch.float32).contiguous()
        self.w_s2d = None
        if self.precision in ("bf16", "tf32") and self.stem == "rgb":
            # stride-1 4x4 conv over 2x2 space-to-depth cells: [k][r'][s'][(dy*2+dx)*3+c] (+4 zero channels).
            # tf32 mode: the stem runs on fp16 operands — pixel values (integers 0..255) are exact in fp16 and its 11-bit
            # significand rounds the weights no coarser than TF32's 10 bits would; fp32 accumulation and output
            w8 = torch.zeros(self.K, 8, 8, 3, device=w.device, dtype=w.dtype)
            w8[:, :7, :7] = w.permute(0, 2, 3, 1)
            w8 = w8.reshape(self.K, 4, 2, 4, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(self.K, 4, 4, 12)
            ws = torch.zeros(self.K, 4, 4, 16, device=w.device, dtype=w.dtype)
            ws[..., :12] = w8
            self.w_s2d = ws.reshape(self.K, 256).to(torch.float16 if self.precision == "tf32" else wdt).contiguous()
        if isinstance(self.norm, nn.BatchNorm2d):
            n = self.norm
            sc = n.weight.detach().float() / torch.sqrt(n.running_var.detach().float() + n.eps)
            sh = n.bias.detach().float() - n.running_mean.detach().float() * sc
            self.scale = (sc * self.in_scale).contiguous()
            self.bn_scale = sc.contiguous()  # without the input scaling (taped fp32 path of trainable trunks)
            self.shift = sh.contiguous()
            self.pre_scale = torch.full_like(sc, self.in_scale)
            self.pre_shift = torch.zeros_like(sc)
        self._ver = ver


class ConvTrunk:
    # RGB trunk only: its conv CTAs claim just over half an SM's shared memory (one CTA per SM) so the
    # depth trunk / head kernels of concurrent graph branches co-run on the other half
    exclusive = False

    # 16-bit tensor-core operand type of the "bf16" precision mode: bf16 for the BN-folded RGB
    # trunk (unbounded residual stream), fp16 for the GroupNorm depth trunk (activations are
    # renormalised to O(1) every layer; fp16 carries 3 more significand bits at the same cost).
    half_dtype = torch.bfloat16

    def __init__(self, device, precision: str = "bf16"):
        assert precision in ("bf16", "fp32", "tf32")
        self.device = device
        self.precision = precision
        self.adt = self.half_dtype if precision == "bf16" else torch.float32
        self.pool = BufferPool(device)
        self._slot = 0

    # -------------------------------------------------------------- low-level steps
    _pfx = ""  # buffer-name prefix of the batch chunk being sequenced (chunks run concurrently: own buffers)

    # ---- trainable trunks: while ``tape`` is a list the fp32 forward keeps every layer's input, raw
    # conv output and output in its own buffer and appends one record per layer; ``backward`` walks it in reverse.
    tape = None

    def _name(self, tag: str) -> str:
        return self._pfx + tag + (f"@{len(self.tape)}" if self.tape is not None else "")

    def _act(self, tag: str, shape):
        return self.pool.get(self._name(tag), shape, self.adt)

    def _raw(self, tag, shape):
        """fp32 raw conv output feeding GroupNorm / train-mode BatchNorm statistics (one bf16
        rounding per layer instead of two)."""
        return self.pool.get(self._name(tag), shape, torch.float32)

    def begin_tape(self):
        assert self.precision == "fp32", "trainable trunks run the fp32 path"
        self.tape = []

    def backward(self, dout_ptr: int, ld: int, g: Dict[str, torch.Tensor], pname: Dict[int, str]):
        """Backward of the last taped forward.  ``dout_ptr``: fp32 gradient w.r.t. the pooled positions
        (N, O*O, ld) the forward wrote (channels [0, C) of each row); ``g``: parameter name -> zeroed gradient tensor;
        ``pname``: id(parameter) -> its name in ``g``.  Layer by layer: ReLU mask -> residual branch -> norm backward on
        the raw conv output -> wgrad (KRSC, permuted into the (K,C,R,S) parameter gradient) and dgrad."""
        from .backward import _colsum

        tape, st = self.tape, lib.stream()
        assert tape, "backward needs a taped forward (begin_tape)"
        f32 = torch.float32
        grads: Dict[int, torch.Tensor] = {}
        for i in range(len(tape) - 1, -1, -1):
            rec = tape[i]
            kind = rec["kind"]
            if kind == "avgpool":
                n, h, w, c = rec["shape_in"]
                dx = self.pool.get(f"{self._pfx}d@{i}", rec["shape_in"], f32)
                C.vlnce_adaptive_avgpool_bwd(dout_ptr, ld, ptr(dx), n, h, w, c, rec["O"], st)
                grads[rec["x"].data_ptr()] = dx
                continue
            if kind == "maxpool":
                n, h, w, c = rec["shape_in"]
                dy = grads.pop(rec["y"].data_ptr())
                dx = self.pool.get(f"{self._pfx}d@{i}", rec["shape_in"], f32)
                C.vlnce_maxpool3x3s2_bwd(ptr(rec["x"]), ptr(dy), ptr(dx), n, h, w, c, st)
                grads[rec["x"].data_ptr()] = dx
                continue
            cv, x, raw, y = rec["cv"], rec["x"], rec["raw"], rec["y"]
            n, p, q, K = rec["shape"]
            rows = n * p * q
            dy = grads.pop(y.data_ptr())
            if rec["relu"]:
                C.vlnce_relu_bwd(ptr(dy), K, ptr(y), K, rows, K, st)
            res = rec["residual"]
            if res is not None:
                k = res.data_ptr()
                if k in grads:
                    C.vlnce_axpy(ptr(grads[k]), ptr(dy), 1.0, dy.numel(), st)
                else:
                    grads[k] = dy  # dy is not written again by this layer; later consumers accumulate into it
            d_raw = self.pool.get(f"{self._pfx}draw@{i}", rec["shape"], f32)
            norm = cv.norm
            if kind == "gn":
                part = self.pool.get(self._pfx + "gn_part", (n, 2 * K), f32)
                C.vlnce_groupnorm_bwd(ptr(raw), ptr(dy), ptr(norm.weight), ptr(d_raw), ptr(part), n, p * q, K,
                                      norm.num_groups, float(norm.eps), st)
                gb = self.pool.get(self._pfx + "gn_gb", (2 * K,), f32)
                _colsum(ptr(part), 2 * K, ptr(gb), n, 2 * K)
                g[pname[id(norm.weight)]].copy_(gb[:K])
                g[pname[id(norm.bias)]].copy_(gb[K:])
            else:
                C.vlnce_batchnorm_bwd(ptr(raw), ptr(dy), ptr(norm.weight), ptr(norm.running_mean),
                                      ptr(norm.running_var), ptr(d_raw), ptr(g[pname[id(norm.weight)]]),
                                      ptr(g[pname[id(norm.bias)]]), rows, K, float(norm.eps), int(rec["training"]), st)
            n, h, w, c = rec["shape_in"]
            dw = self.pool.get(self._pfx + "dw_krsc", (K, cv.R, cv.S, c), f32)
            C.vlnce_conv2d_ref_wgrad(ptr(x), lib.code(x), ptr(d_raw), ptr(dw), n, h, w, c, K, cv.R, cv.S, cv.stride, cv.pad,
                                     float(cv.in_scale), 0, st)
            gw = g[pname[id(cv.conv.weight)]]
            gw.copy_(dw.permute(0, 3, 1, 2))
            if cv.in_chan_scale is not None:  # normalize_rgb: the packed weights carry 1/std per input channel
                gw.mul_(torch.tensor(cv.in_chan_scale, device=gw.device, dtype=gw.dtype).view(1, -1, 1, 1))
            if rec.get("input"):
                continue  # the image / depth map has no gradient
            k = x.data_ptr()
            if k in grads:
                C.vlnce_conv2d_ref_dgrad(ptr(d_raw), ptr(cv.w), ptr(grads[k]), n, h, w, c, K, cv.R, cv.S, cv.stride,
                                         cv.pad, 1, st)
            else:
                dx = self.pool.get(f"{self._pfx}d@{i}", rec["shape_in"], f32)
                C.vlnce_conv2d_ref_dgrad(ptr(d_raw), ptr(cv.w), ptr(dx), n, h, w, c, K, cv.R, cv.S, cv.stride, cv.pad, 0, st)
                grads[k] = dx
        assert not grads, "activation gradients left unconsumed"

    def _conv(self, x, shape_in, cv: _Conv, y, *, scale, shift, residual=None, relu=False,
              gemm_c: Optional[int] = None, stem_rgb: bool = False, stem_s2d: bool = False):
        """x NHWC (n,h,w,c) -> y; ``gemm_c`` set = x is an explicit stem operand (n*p*q, gemm_c);
        ``stem_rgb`` = x is the raw (n,h,w,3) image and the 7x7/2 window is gathered in-kernel."""
        n, h, w, c = shape_in
        st = lib.stream()
        if self.precision == "bf16" or stem_s2d or gemm_c is not None:   # (tf32 mode: only its fp16 stems come this way)
            if stem_s2d:  # x = s2d cells, (n, h, w) = the stem's OUTPUT size
                geom = ("stem_s2d", n, h, w, cv.K)
            elif stem_rgb:
                geom = ("stem", n, h, w, cv.K)
            elif gemm_c is not None:
                geom = (n * h * w, 1, 1, gemm_c, cv.K, 1, 1, 1, 0)
            else:
                geom = (n, h, w, c, cv.K, cv.R, cv.S, cv.stride, cv.pad)

            def launch(code):
                if self.exclusive and not (code & TMAP):
                    code |= EXCL
                if stem_s2d:
                    C.vlnce_conv_stem_tma(ptr(x), ptr(cv.w_s2d), lib.code(cv.w_s2d), ptr(scale), ptr(shift), ptr(y),
                                          lib.code(y), n, h, w, cv.K, int(relu), code & 0x6fffff, st)
                elif stem_rgb:
                    C.vlnce_conv2d_stem_rgb(ptr(x), lib.code(x), ptr(cv.w), lib.code(cv.w), ptr(scale), ptr(shift),
                                            ptr(y), lib.code(y), n, h, w, cv.K, int(relu), code, st)
                elif code & S2TMA:
                    C.vlnce_conv_s2_tma(ptr(x), ptr(cv.w), lib.code(cv.w), ptr(scale), ptr(shift), ptr(y), lib.code(y),
                                        n, h, w, c, cv.K, cv.R, int(relu), code & 0x6fffff, st)
                elif code & TMAP:
                    C.vlnce_conv_tma(ptr(x), ptr(cv.w), lib.code(cv.w), ptr(scale), ptr(shift), ptr(residual), ptr(y),
                                     lib.code(y), geom[0], geom[1], geom[2], geom[3], cv.K, geom[5], int(relu),
                                     code & 0x6fffff, st)
                elif code & HALO:
                    C.vlnce_conv3x3_halo(ptr(x), ptr(cv.w), lib.code(cv.w), ptr(scale), ptr(shift), ptr(residual),
                                         ptr(y), lib.code(y), n, h, w, c, cv.K, int(relu), code & 0x1fffff, st)
                else:
                    C.vlnce_conv2d_bf16(ptr(x), ptr(cv.w), lib.code(cv.w), ptr(scale), ptr(shift), ptr(residual),
                                        ptr(y), lib.code(y), *geom, int(relu), code, st)

            c_in = geom[3] if not (stem_rgb or stem_s2d) else 0
            tma_ok = (PERSISTENT and not stem_rgb and not stem_s2d and c_in % 64 == 0 and cv.K % 64 == 0 and geom[7:] in ((1, 0), (1, 1))
                      and ((geom[5] == 1 and geom[8] == 0)
                           # 3x3: 16x8-pixel halo tiles; sizes that are not multiples of the tile (224-pixel frames: 56 / 28 /
                           # 14-pixel maps) get ragged border tiles, zero-filled / clipped by the TMA unit
                           or (geom[5] == 3 and geom[8] == 1 and (RAGGED_TMA or (geom[1] % 16 == 0 and geom[2] % 8 == 0))
                               and geom[1] >= 14 and geom[2] >= 8)))
            halo_ok = (HALO_3X3 and not stem_rgb and gemm_c is None and cv.R == 3 and cv.S == 3 and cv.stride == 1
                       and cv.pad == 1 and c % 64 == 0 and cv.K % 64 == 0 and h % 16 == 0 and w % 8 == 0)
            s2_ok = False
            if (PERSISTENT and not stem_rgb and not stem_s2d and gemm_c is None and residual is None
                    and (cv.R, cv.S, cv.pad) in ((1, 1, 0), (3, 3, 1)) and cv.stride == 2 and c % 64 == 0 and cv.K % 64 == 0
                    and h % 2 == 0 and w % 2 == 0 and (w // 2) % 8 == 0):
                bh = min(16, h // 2)
                s2_ok = (h // 2) % bh == 0 and 128 % (8 * bh) == 0 and n % (128 // (8 * bh)) == 0
            nb = _batch_bucket(n)
            if stem_s2d or stem_rgb:
                gkey = (geom[0], nb) + tuple(geom[2:])
            elif gemm_c is not None:
                gkey = ("gemm", nb, h, w) + tuple(geom[3:])
            else:
                gkey = (nb,) + tuple(geom[1:])
            key = (gkey, lib.code(y), residual is not None)
            bn = cv.block_n.get(key)
            if bn is None:
                skey = repr((key, lib.code(cv.w), scale is not None, relu, halo_ok, tma_ok, self.exclusive)
                            + ((True,) if s2_ok else ()))   # (appended only when set: older table keys stay valid)
                if skey in _TUNE_CACHE:
                    bn = int(_TUNE_CACHE[skey])
                else:
                    tune = AUTOTUNE and x.is_cuda and not torch.cuda.is_current_stream_capturing()
                    if tune:
                        bn = self._tune(launch, cv.K, halo_ok, tma_ok, stem_s2d, s2_ok)
                        _tune_cache_put(skey, bn)
                    elif stem_s2d:
                        bn = 64 | MT2
                    else:
                        bn = TMAP if tma_ok else (HALO if halo_ok else (S2TMA if s2_ok else 0))
                cv.block_n[key] = bn
            launch(bn)
        else:
            if self.precision == "tf32" and gemm_c is None and not stem_rgb and not stem_s2d and x.dtype == torch.float32:
                # fp32 storage, TF32 products on the tensor cores wherever the persistent TMA kernel covers the geometry
                geo = (cv.R, cv.S, cv.stride, cv.pad)
                chan_ok = c % 32 == 0 and (cv.K % 64 == 0 or (cv.K == 32 and scale is None and residual is None))
                small3 = (geo == (3, 3, 1, 1) and residual is None and h < 16 and w % 8 == 0 and 128 % (8 * h) == 0
                          and n % (128 // (8 * h)) == 0)   # 8x8 / 4x8 images: several per tile (tap-shifted boxes)
                bn32 = TF32_BLOCK_N if (TF32_BLOCK_N and cv.K % (TF32_BLOCK_N & 0xffff) == 0) else 0
                if chan_ok and (geo == (1, 1, 1, 0) or small3 or (geo == (3, 3, 1, 1) and h >= 14 and w >= 8 and
                                                                  (RAGGED_TMA or (h % 16 == 0 and w % 8 == 0)))):
                    C.vlnce_conv_tma(ptr(x), ptr(cv.w), lib.F32, ptr(scale), ptr(shift), ptr(residual), ptr(y), lib.F32,
                                     n, h, w, c, cv.K, cv.R, int(relu), bn32, st)
                    return
                if (chan_ok and residual is None and geo in ((1, 1, 2, 0), (3, 3, 2, 1)) and h % 2 == 0 and w % 2 == 0
                        and (w // 2) % 8 == 0):
                    bh = min(16, h // 2)
                    if (h // 2) % bh == 0 and 128 % (8 * bh) == 0 and n % (128 // (8 * bh)) == 0:
                        C.vlnce_conv_s2_tma(ptr(x), ptr(cv.w), lib.F32, ptr(scale), ptr(shift), ptr(y), lib.F32, n, h, w, c,
                                            cv.K, cv.R, int(relu), bn32, st)
                        return
                if c % 32 == 0:   # what the tiles do not cover (4x4 images, odd sizes): fp32 on CUDA cores, one warp
                    # per 8 pixels x 4 filters with the lanes across the input channels
                    C.vlnce_conv2d_warp(ptr(x), ptr(cv.w), ptr(scale), ptr(shift), ptr(residual), ptr(y), n, h, w, c,
                                        cv.K, cv.R, cv.S, cv.stride, cv.pad, int(relu), st)
                    return
            C.vlnce_conv2d_ref(ptr(x), lib.code(x), ptr(cv.w), lib.F32, ptr(scale), ptr(shift),
                               ptr(residual), lib.F32, ptr(y), lib.F32, n, h, w, c, cv.K, cv.R, cv.S,
                               cv.stride, cv.pad, int(relu), st)

    @staticmethod
    def _tune(launch, K, halo_ok=False, tma_ok=False, stem_s2d=False, s2_ok=False):
        """Pick kernel variant, output-tile width and ring depth empirically (once per layer
        shape): CUDA-event timing of the legal candidates on the real operands; the winner is
        cached on the layer."""
        widths = [bn for bn in (64, 128, 256) if K % bn == 0] or [32]
        cands = [bn | (st << 16) for bn in widths for st in (1, 2, 4)]
        if halo_ok:
            cands += [HALO | bn | (st << 16) for bn in widths for st in (3, 6)]
        if tma_ok:
            cands += [TMAP | bn | (4 << 16) | fl for bn in widths for fl in (0, NOSTAT)]
            cands += [TMAP | bn | (4 << 16) | MT2 | fl for bn in widths if bn <= 128 for fl in (0, NOSTAT)]
        if s2_ok:
            cands += [S2TMA | bn | (4 << 16) | fl for bn in widths for fl in (0, NOSTAT)]
            cands += [S2TMA | bn | (4 << 16) | MT2 | fl for bn in widths if bn <= 128 for fl in (0, NOSTAT)]
        if stem_s2d:  # only the persistent TMA kernel runs this geometry
            cands = [bn | (4 << 16) | fl for bn in widths if bn <= 128 for fl in (0, MT2, NOSTAT, MT2 | NOSTAT)]
        best, best_t = 0, float("inf")
        for _ in range(12):  # bring the clocks up before timing anything
            launch(cands[-1])
        for bn in cands:
            launch(bn)
            launch(bn)
            t = float("inf")
            for _ in range(3):  # best of three short bursts: one noisy burst must not decide a layer's kernel
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(6):
                    launch(bn)
                b.record()
                b.synchronize()
                t = min(t, a.elapsed_time(b))
            if t < best_t:
                best, best_t = bn, t
        return best

    @staticmethod
    def _out_hw(h, w, cv: _Conv):
        return ((h + 2 * cv.pad - cv.R) // cv.stride + 1, (w + 2 * cv.pad - cv.S) // cv.stride + 1)


# ======================================================================================
class RGBTrunk(ConvTrunk):
    """torchvision ResNet trunk; ``cnn`` = the nn.Sequential holding conv1,bn1,relu,maxpool,
    layer1..4 (state_dict keys ``net.rgb_encoder.cnn.{0..7}.*``)."""

    exclusive = RGB_EXCLUSIVE

    # MODEL.normalize_rgb (resnet_encoders.py:181-190): ImageNet statistics
    MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    def __init__(self, cnn: nn.Sequential, device, precision="bf16", half_dtype=torch.bfloat16, normalize=False):
        # eval-mode (folded) BatchNorm leaves the residual stream unbounded -> bf16 range;
        # train-mode BatchNorm renormalises every layer to O(1) -> fp16 precision.
        self.half_dtype = half_dtype
        super().__init__(device, precision)
        mods = list(cnn.children())
        _C = lambda c, g, p, **kw: _Conv(c, g, p, wdt=half_dtype, **kw)  # noqa: E731
        # normalize: (x/255 - mean)/std = (x - 255 mean) * (1/255) * (1/std): the centring is a pre-pass over the
        # pixels (the conv's zero padding is applied after the normalisation), 1/std goes into the weights
        self.normalize = bool(normalize)
        self.stem = _C(mods[0], mods[1], precision, in_scale=1.0 / 255.0, stem="rgb",
                       in_chan_scale=[1.0 / v for v in self.STD] if normalize else None)
        self.blocks = []
        for layer in mods[4:8]:
            for blk in layer.children():
                d = {"c1": _C(blk.conv1, blk.bn1, precision), "c2": _C(blk.conv2, blk.bn2, precision)}
                if hasattr(blk, "conv3"):
                    d["c3"] = _C(blk.conv3, blk.bn3, precision)
                if blk.downsample is not None:
                    d["ds"] = _C(blk.downsample[0], blk.downsample[1], precision)
                self.blocks.append(d)
        self.out_channels = self.blocks[-1]["c3" if "c3" in self.blocks[-1] else "c2"].K
        self._bn_ws = None
        self._s2d_ok = self.stem.K % 64 == 0

    def _all(self) -> List[_Conv]:
        out = [self.stem]
        for b in self.blocks:
            out += list(b.values())
        return out

    def _norm_conv(self, x, shape_in, cv: _Conv, tag, *, residual=None, relu=True, gemm_c=None,
                   out_hw=None, stem_rgb=False, stem_s2d=False):
        n, h, w, _ = shape_in
        p, q = out_hw if out_hw is not None else self._out_hw(h, w, cv)
        y = self._act(tag, (n, p, q, cv.K))
        bn: nn.BatchNorm2d = cv.norm
        if self.tape is not None and not bn.training:
            # trainable trunk, eval-mode norms: same raw -> affine split as the train-mode path below so that the
            # backward has the raw conv output; scale / shift are the folded running statistics
            raw = self._raw("bn_raw", (n, p, q, cv.K))
            pre = cv.in_scale != 1.0
            self._conv(x, shape_in, cv, raw, scale=cv.pre_scale if pre else None, shift=cv.pre_shift if pre else None)
            C.vlnce_affine_apply(ptr(raw), lib.code(raw), ptr(cv.bn_scale), ptr(cv.shift), 0, ptr(residual), ptr(y),
                                 lib.code(y), n, p * q, cv.K, int(relu), lib.stream())
            self.tape.append(dict(kind="bn", cv=cv, x=x, shape_in=tuple(shape_in), raw=raw, y=y, residual=residual,
                                  relu=relu, shape=(n, p, q, cv.K), training=False, input=cv.stem is not None))
            return y, (n, p, q, cv.K)
        if not bn.training:
            self._conv(x, shape_in, cv, y, scale=cv.scale, shift=cv.shift, residual=residual, relu=relu,
                       gemm_c=gemm_c, stem_rgb=stem_rgb, stem_s2d=stem_s2d)
            return y, (n, p, q, cv.K)
        # train-mode BN (Q1): raw conv (input scaling only) -> batch stats -> fused apply
        raw = self._raw("bn_raw", (n, p, q, cv.K))
        self._conv(x, shape_in, cv, raw, scale=cv.pre_scale if cv.in_scale != 1.0 else None,
                   shift=cv.pre_shift if cv.in_scale != 1.0 else None, gemm_c=gemm_c, stem_rgb=stem_rgb,
                   stem_s2d=stem_s2d)
        st = lib.stream()
        sc = self.pool.get("bn_scale", (cv.K,), torch.float32)
        sh = self.pool.get("bn_shift", (cv.K,), torch.float32)
        ws = self.pool.get("bn_ws", (2 * cv.K,), torch.float64)
        mom = 0.1 if bn.momentum is None else bn.momentum
        C.vlnce_batchnorm_train_stats(ptr(raw), lib.code(raw), ptr(bn.weight), ptr(bn.bias),
                                      ptr(bn.running_mean), ptr(bn.running_var), float(mom), ptr(sc),
                                      ptr(sh), ptr(ws), n * p * q, cv.K, float(bn.eps), st)
        bn.num_batches_tracked += 1
        cv._ver = None  # running stats changed under the folded eval-mode scale/shift
        C.vlnce_affine_apply(ptr(raw), lib.code(raw), ptr(sc), ptr(sh), 0, ptr(residual), ptr(y), lib.code(y),
                             n, p * q, cv.K, int(relu), st)
        if self.tape is not None:
            self.tape.append(dict(kind="bn", cv=cv, x=x, shape_in=tuple(shape_in), raw=raw, y=y, residual=residual,
                                  relu=relu, shape=(n, p, q, cv.K), training=True, input=cv.stem is not None))
        return y, (n, p, q, cv.K)

    def forward(self, rgb: torch.Tensor, out: torch.Tensor, ld_out: int, out_hw: int = 4):
        """rgb (R,H,W,3) u8|f32 in 0..255 -> out[(R,16,ld_out) f32][:, :, :C] = 4x4-pooled features."""
        for cv in self._all():
            cv.pack()
        n, h, w, _ = rgb.shape
        st = lib.stream()
        if self.normalize:
            centred = self.pool.get("rgb_centred", (n, h, w, 3), torch.float32)
            C.vlnce_rgb_center(ptr(rgb), lib.code(rgb), ptr(centred), *(255.0 * m for m in self.MEAN), n * h * w, st)
            rgb = centred
        cv = self.stem
        p, q = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
        x = None
        if self.precision in ("bf16", "tf32") and S2D_STEM and self._s2d_ok:
            # stem as a stride-1 4x4 conv over 2x2 space-to-depth cells, fully TMA-fed (overlapping
            # 4-cell windows read straight from the cell image); falls back to the gather stem if the
            # driver rejects the overlapping-stride tensor map
            cells = self.pool.get("stem_s2d", (n, p + 3, q + 3, 16), torch.float16 if self.precision == "tf32" else self.adt)
            try:
                C.vlnce_stem_s2d(ptr(rgb), lib.code(rgb), ptr(cells), lib.code(cells), n, h, w, st)
                x, shp = self._norm_conv(cells, (n, p, q, 16), cv, "stem_y", out_hw=(p, q), stem_s2d=True)
            except lib.VlnceError:
                if torch.cuda.is_current_stream_capturing():
                    raise
                self._s2d_ok = False
                x = None
        if x is not None:
            pass
        elif self.precision == "bf16" and FUSED_STEM:
            # im2col-free stem: the 7x7x3 windows are gathered inside the conv tile's producer warps
            x, shp = self._norm_conv(rgb, (n, h, w, 3), cv, "stem_y", out_hw=(p, q), stem_rgb=True)
        elif self.precision == "bf16":
            a = self.pool.get("stem_a", (n * p * q, 192), self.adt)
            C.vlnce_im2col_rgb_stem(ptr(rgb), lib.code(rgb), ptr(a), lib.code(a), n, h, w, st)
            x, shp = self._norm_conv(a, (n, p, q, 192), cv, "stem_y", gemm_c=192, out_hw=(p, q))
        else:
            x, shp = self._norm_conv(rgb, (n, h, w, 3), cv, "stem_y")
        n, h, w, c = shp
        p, q = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        y = self._act("pool_y", (n, p, q, c))
        C.vlnce_maxpool3x3s2(ptr(x), ptr(y), lib.code(x), n, h, w, c, st)
        if self.tape is not None:
            self.tape.append(dict(kind="maxpool", x=x, y=y, shape_in=(n, h, w, c)))
        x, shp = y, (n, p, q, c)
        slot = 0
        for b in self.blocks:
            identity = x
            if "ds" in b:
                identity, _ = self._norm_conv(x, shp, b["ds"], "ds", relu=False)
            o, s1 = self._norm_conv(x, shp, b["c1"], "t1")
            if "c3" in b:
                o, s2 = self._norm_conv(o, s1, b["c2"], "t2")
                x, shp = self._norm_conv(o, s2, b["c3"], f"blk{slot}", residual=identity, relu=True)
            else:
                x, shp = self._norm_conv(o, s1, b["c2"], f"blk{slot}", residual=identity, relu=True)
            slot ^= 1
        n, h, w, c = shp
        C.vlnce_adaptive_avgpool(ptr(x), lib.code(x), ptr(out), ld_out, n, h, w, c, out_hw, st)
        if self.tape is not None:
            self.tape.append(dict(kind="avgpool", x=x, shape_in=(n, h, w, c), O=out_hw))
        return shp


# ======================================================================================
class DepthTrunk(ConvTrunk):
    """DD-PPO GroupNorm ResNet encoder; ``enc`` has ``backbone`` (conv1, maxpool, layer1..4 of
    blocks with ``convs`` / ``downsample`` Sequentials) and ``compression``."""

    half_dtype = torch.float16

    def __init__(self, enc: nn.Module, device, precision="bf16"):
        super().__init__(device, precision)
        bb = enc.backbone
        _C = lambda c, g, p, **kw: _Conv(c, g, p, wdt=torch.float16, **kw)  # noqa: E731
        self.stem = _C(bb.conv1[0], bb.conv1[1], precision, stem="depth")
        self.blocks = []
        for layer in (bb.layer1, bb.layer2, bb.layer3, bb.layer4):
            for blk in layer.children():
                convs = list(blk.convs.children())
                pairs = [(convs[i], convs[i + 1]) for i in range(len(convs)) if isinstance(convs[i], nn.Conv2d)]
                d = {"convs": [_C(c, g, precision) for c, g in pairs]}
                if blk.downsample is not None:
                    d["ds"] = _C(blk.downsample[0], blk.downsample[1], precision)
                self.blocks.append(d)
        self.compress = _C(enc.compression[0], enc.compression[1], precision)
        self.out_channels = self.compress.K

    def _all(self):
        out = [self.stem, self.compress]
        for b in self.blocks:
            out += b["convs"] + ([b["ds"]] if "ds" in b else [])
        return out

    def _gn_conv(self, x, shape_in, cv: _Conv, tag, *, residual=None, relu=True, gemm_c=None, out_hw=None):
        n, h, w, _ = shape_in
        p, q = out_hw if out_hw is not None else self._out_hw(h, w, cv)
        gn: nn.GroupNorm = cv.norm
        c4n = cv.K // 4
        raw16 = (GN_RAW16 and self.adt == torch.float16 and c4n <= 256 and 256 % c4n == 0
                 and (cv.K // gn.num_groups) % 2 == 0 and gn.num_groups <= 64)
        raw = self._act("gn_raw16", (n, p, q, cv.K)) if raw16 else self._raw("gn_raw", (n, p, q, cv.K))
        self._conv(x, shape_in, cv, raw, scale=None, shift=None, gemm_c=gemm_c)
        st = lib.stream()
        sc = self.pool.get(self._pfx + "gn_scale", (n, cv.K), torch.float32)
        sh = self.pool.get(self._pfx + "gn_shift", (n, cv.K), torch.float32)
        ws = self.pool.get(self._pfx + "gn_ws", (1 << 20,), torch.float32)
        y = self._act(tag, (n, p, q, cv.K))
        if raw16:
            C.vlnce_groupnorm_fused16(ptr(raw), ptr(gn.weight), ptr(gn.bias), ptr(residual), ptr(y), ptr(ws),
                                      ws.numel() * 4, n, p * q, cv.K, gn.num_groups, float(gn.eps), int(relu), st)
        elif raw.dtype == torch.float32:
            C.vlnce_groupnorm_fused(ptr(raw), ptr(gn.weight), ptr(gn.bias), ptr(residual), ptr(y), lib.code(y), ptr(ws),
                                    ws.numel() * 4, ptr(sc), ptr(sh), n, p * q, cv.K, gn.num_groups, float(gn.eps),
                                    int(relu) | (2 if self.precision == "tf32" else 0), st)
        else:
            C.vlnce_groupnorm_stats(ptr(raw), lib.code(raw), ptr(gn.weight), ptr(gn.bias), ptr(sc), ptr(sh), ptr(ws),
                                    ws.numel() * 4, n, p * q, cv.K, gn.num_groups, float(gn.eps), st)
            C.vlnce_affine_apply(ptr(raw), lib.code(raw), ptr(sc), ptr(sh), cv.K, ptr(residual), ptr(y),
                                 lib.code(y), n, p * q, cv.K, int(relu), st)
        if self.tape is not None:
            self.tape.append(dict(kind="gn", cv=cv, x=x, shape_in=tuple(shape_in), raw=raw, y=y, residual=residual,
                                  relu=relu, shape=(n, p, q, cv.K), input=cv.stem is not None))
        return y, (n, p, q, cv.K)

    def forward(self, depth: torch.Tensor, out: torch.Tensor, ld_out: int, out_hw: int = 4):
        """depth (R,H,W,1) f32 -> out[(R,16,ld_out) f32][:, :, :128].

        GroupNorm normalises per sample, so the batch can be cut into independent chunks with bit-identical
        results: with ``DEPTH_CHUNKS`` > 1 every chunk runs the whole trunk on its own stream (parallel branches
        of a captured graph).  The trunk is a chain of ~160 short kernels of 16-128 CTAs each; several chains side
        by side fill the SMs that one chain leaves idle."""
        for cv in self._all():
            cv.pack()
        n = depth.shape[0]
        if (DEPTH_FUSED and self.fused and self.precision == "bf16" and self.tape is None and depth.is_cuda
                and self._forward_fused(depth, out, ld_out, out_hw)):
            return (n, out_hw, out_hw, self.out_channels)
        k = DEPTH_CHUNKS if (self.precision == "bf16" and depth.is_cuda and n >= 8 * DEPTH_CHUNKS) else 1
        if k <= 1:
            return self._forward_chunk(depth, out, ld_out, out_hw, "")
        from .engine import Branch

        per = (n + k - 1) // k
        branches, shp = [], None
        for i in range(k):
            lo, hi = i * per, min(n, (i + 1) * per)
            if lo >= hi:
                break
            with Branch(f"depth_chunk{i}") as br:
                shp = self._forward_chunk(depth[lo:hi], out[lo:hi], ld_out, out_hw, f"c{i}_")
            branches.append(br)
        for br in branches:
            br.join()
        return (n,) + tuple(shp[1:])

    # ------------------------------------------------------------------ image-stationary path
    fused = True          # instance switch (tests compare the two paths)
    _plans = None         # (geometry, buffers, weight versions) -> device plan

    def _fused_table(self, n, h, w, out, ld_out):
        """Step table of ``vlnce_depth_fused_plan`` (include/vlnce_b200.h) for a (n, h, w, 1) depth batch: the stem
        as a 1x1 conv over the padded im2col operand, max-pool, every bottleneck / basic block (downsample first),
        the compression; the last step writes fp32 straight into the caller's (R, positions, ld) feature rows."""
        import struct

        hp, wp = h // 2, w // 2
        p, q = (hp + 6 - 7) // 2 + 1, (wp + 6 - 7) // 2 + 1
        f16 = torch.float16
        rows = []

        def conv(x, cv, y, hin, win, *, residual=None, relu=True, y_f32=None, ld=0, external=False, cin=None, k=None,
                 stride=None):
            gn = cv.norm
            eps_bits = struct.unpack("<I", struct.pack("<f", float(gn.eps)))[0]
            rows.append([0, ptr(x), ptr(cv.w), ptr(gn.weight), ptr(gn.bias), ptr(residual) or 0, ptr(y) or 0,
                         ptr(y_f32) or 0, ld, cv.Cin if cin is None else cin, cv.K, cv.R if k is None else k,
                         cv.stride if stride is None else stride, hin, win, gn.num_groups, int(relu), int(external),
                         eps_bits, 0])

        a = self.pool.get("stem_a", (n * p * q, 64), f16)
        y = self.pool.get("fz_stem", (n, p, q, self.stem.K), f16)
        conv(a, self.stem, y, p, q, external=True, cin=64, k=1, stride=1)
        hh, ww, c = p // 2, q // 2, self.stem.K
        x = self.pool.get("fz_pool", (n, hh, ww, c), f16)
        rows.append([1, ptr(y), 0, 0, 0, 0, ptr(x), 0, 0, c, c, 3, 2, p, q, 1, 0, 0, 0, 0])
        slot = 0
        for b in self.blocks:
            identity = x
            if "ds" in b:
                cv = b["ds"]
                ho, wo = hh // cv.stride, ww // cv.stride
                identity = self.pool.get(f"fz_ds_{ho}x{wo}x{cv.K}", (n, ho, wo, cv.K), f16)
                conv(x, cv, identity, hh, ww, relu=False)
            o, oh, ow = x, hh, ww
            for i, cv in enumerate(b["convs"]):
                last = i == len(b["convs"]) - 1
                ho, wo = oh // cv.stride, ow // cv.stride
                tag = f"fz_blk{slot}" if last else f"fz_t{i}"
                yb = self.pool.get(f"{tag}_{ho}x{wo}x{cv.K}", (n, ho, wo, cv.K), f16)
                conv(o, cv, yb, oh, ow, residual=identity if last else None, relu=True)
                o, oh, ow = yb, ho, wo
            x, hh, ww = o, oh, ow
            slot ^= 1
        conv(x, self.compress, None, hh, ww, relu=True, y_f32=out, ld=ld_out)
        return a, torch.tensor(rows, dtype=torch.int64), (hh, ww)

    def _forward_fused(self, depth, out, ld_out, out_hw) -> bool:
        """One launch for the whole trunk; False = geometry outside what the fused kernel covers (the caller falls
        back to the per-layer launches)."""
        n, h, w, _ = depth.shape
        if self._plans is None:
            self._plans = {}
        key = (n, h, w, out.data_ptr(), ld_out, tuple(cv._ver for cv in self._all()))
        plan = self._plans.get(key)
        if plan is None:
            if torch.cuda.is_current_stream_capturing():
                raise lib.VlnceError("depth trunk: run one eager forward before capturing a CUDA graph (the fused plan "
                                     "is built on first use)")
            if h % 64 != 0 or w % 64 != 0:
                plan = False
            else:
                a, table, (hh, ww) = self._fused_table(n, h, w, out, ld_out)
                if (hh, ww) != (out_hw, out_hw):
                    plan = False  # the adaptive pool is not the identity: per-layer path
                else:
                    nbytes = int(C.vlnce_depth_fused_plan_bytes(table.shape[0]))
                    host = torch.zeros(nbytes, dtype=torch.uint8)
                    try:
                        C.vlnce_depth_fused_plan(table.data_ptr(), table.shape[0], n, host.data_ptr(), nbytes)
                        plan = (a, host.to(depth.device), table)
                    except lib.VlnceError:
                        plan = False
            if len(self._plans) > 8:
                self._plans.clear()
            self._plans[key] = plan
        if plan is False:
            return False
        a, dev_plan, _ = plan
        st = lib.stream()
        C.vlnce_im2col_depth_stem(ptr(depth), ptr(a), lib.code(a), n, h, w, st)
        C.vlnce_depth_fused_run(ptr(dev_plan), n, DEPTH_FUSED_CTAS, st)
        return True

    def _forward_chunk(self, depth: torch.Tensor, out: torch.Tensor, ld_out: int, out_hw: int, pfx: str):
        self._pfx = pfx
        n, h, w, _ = depth.shape
        st = lib.stream()
        hp, wp = h // 2, w // 2
        p, q = (hp + 6 - 7) // 2 + 1, (wp + 6 - 7) // 2 + 1
        if self.precision in ("bf16", "tf32"):
            a = self.pool.get(self._pfx + "stem_a", (n * p * q, 64), torch.float16 if self.precision == "tf32" else self.adt)
            C.vlnce_im2col_depth_stem(ptr(depth), ptr(a), lib.code(a), n, h, w, st)
            x, shp = self._gn_conv(a, (n, p, q, 64), self.stem, "stem_y", gemm_c=64, out_hw=(p, q))
        else:
            pooled = self.pool.get("stem_a", (n, hp, wp, 1), torch.float32)
            C.vlnce_avgpool2x2(ptr(depth), ptr(pooled), n, h, w, 1, st)
            x, shp = self._gn_conv(pooled, (n, hp, wp, 1), self.stem, "stem_y")
        n, h, w, c = shp
        p, q = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        y = self._act("pool_y", (n, p, q, c))
        C.vlnce_maxpool3x3s2(ptr(x), ptr(y), lib.code(x), n, h, w, c, st)
        if self.tape is not None:
            self.tape.append(dict(kind="maxpool", x=x, y=y, shape_in=(n, h, w, c)))
        x, shp = y, (n, p, q, c)
        slot = 0
        for b in self.blocks:
            identity = x
            if "ds" in b:
                identity, _ = self._gn_conv(x, shp, b["ds"], "ds", relu=False)
            o, s = x, shp
            for i, cv in enumerate(b["convs"]):
                last = i == len(b["convs"]) - 1
                o, s = self._gn_conv(o, s, cv, f"blk{slot}" if last else f"t{i}",
                                     residual=identity if last else None, relu=True)
            x, shp = o, s
            slot ^= 1
        x, shp = self._gn_conv(x, shp, self.compress, "cmp", relu=True)
        n, h, w, c = shp
        C.vlnce_adaptive_avgpool(ptr(x), lib.code(x), ptr(out), ld_out, n, h, w, c, out_hw, st)
        if self.tape is not None:
            self.tape.append(dict(kind="avgpool", x=x, shape_in=(n, h, w, c), O=out_hw))
        return shp
