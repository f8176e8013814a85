"""Synthetic inputs fixed by SURVEY.md §8(d) so that every consumer (numpy tests, the C
host drivers, the on-device generator in csrc/synth.cu) produces the same bits.

The reference ships neither its image (``Cat_Image0.ppm``, ``MobileNet.c:215``) nor its
weights (``weights_c.txt``, ``MobileNet.c:37``), so the bench and the parity tests run on
seeded data:

* images: u8 uniform, ``[N,224,224,3]`` interleaved like the PPM payload; byte ``j`` of the
  whole stream is byte ``j % 8`` (little endian) of ``splitmix64(seed, j // 8)`` — image ``n``
  therefore does not depend on the batch it is generated in (shards reproduce).
* weights: approximately normal (Irwin–Hall sum of twelve 16-bit uniforms — integer exact,
  no libm, so C / CUDA / numpy agree bit for bit) scaled by ``sqrt(2 / fan_in)``.
* BatchNorm: gamma~U(.5,1.5) beta~N(0,.1) mean~N(0,.1) var~U(.5,1.5) eps=1e-3, folded to
  ``scale = gamma / sqrt(var + eps)``, ``shift = beta - mean * scale``.
"""
from __future__ import annotations

import numpy as np

from .layers import LAYERS, FC, POOL, STEM, DEPTHWISE, TOTAL_WEIGHTS, TOTAL_CHANNELS, IMG

IMAGE_SEED = 0x5EED0001
WEIGHT_SEED = 0x5EED1000
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, counter: np.ndarray) -> np.ndarray:
    """Counter-based splitmix64: word ``k`` of stream ``seed``."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (counter.astype(np.uint64) + np.uint64(1)) * _GOLDEN).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def images(n: int, first: int = 0, seed: int = IMAGE_SEED) -> np.ndarray:
    """u8 ``[n,224,224,3]`` — images ``first .. first+n-1`` of the global stream."""
    per = IMG * IMG * 3  # 150528, a multiple of 8
    words = splitmix64(seed, np.arange(first * per // 8, (first + n) * per // 8, dtype=np.uint64))
    return words.view(np.uint8).reshape(n, IMG, IMG, 3).copy()


def _uniform(seed: int, count: int) -> np.ndarray:
    """double in [0,1): top 53 bits."""
    w = splitmix64(seed, np.arange(count, dtype=np.uint64))
    return (w >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _normal(seed: int, count: int) -> np.ndarray:
    """Irwin–Hall(12) of 16-bit uniforms: mean 0, variance 1 (to 1e-10), support ±6."""
    w = splitmix64(seed, np.arange(3 * count, dtype=np.uint64)).reshape(count, 3)
    total = np.zeros(count, dtype=np.int64)
    for sh in (0, 16, 32, 48):
        total += ((w >> np.uint64(sh)) & np.uint64(0xFFFF)).astype(np.int64).sum(axis=1)
    return (total.astype(np.float64) - 12 * 32767.5) / 65536.0


def weights(seed: int = WEIGHT_SEED) -> np.ndarray:
    """Flat fp32 filter values in the reference's file order (4 209 088 of them)."""
    out = np.empty(TOTAL_WEIGHTS, dtype=np.float32)
    for L in LAYERS:
        if L.kind == POOL:
            continue
        fan_in = 27 if L.kind == STEM else 9 if L.kind == DEPTHWISE else L.cin
        std = np.sqrt(2.0 / fan_in) if L.kind != FC else np.sqrt(1.0 / fan_in)
        out[L.w_off:L.w_off + L.w_cnt] = (_normal(seed + L.index, L.w_cnt) * std).astype(np.float32)
    return out


def batchnorm(seed: int = WEIGHT_SEED) -> tuple[np.ndarray, np.ndarray]:
    """Folded per-channel (scale, shift), ``TOTAL_CHANNELS`` long; the last 1000 are the FC
    layer's (scale 1, shift = bias ~ N(0, 0.01))."""
    scale = np.ones(TOTAL_CHANNELS, dtype=np.float32)
    shift = np.zeros(TOTAL_CHANNELS, dtype=np.float32)
    for L in LAYERS:
        if L.kind == POOL:
            continue
        s = seed + L.index
        if L.kind == FC:
            shift[L.c_off:L.c_off + L.cout] = (_normal(s + 0x100, L.cout) * 0.01).astype(np.float32)
            continue
        gamma = 0.5 + _uniform(s + 0x100, L.cout)
        beta = 0.1 * _normal(s + 0x200, L.cout)
        mean = 0.1 * _normal(s + 0x300, L.cout)
        var = 0.5 + _uniform(s + 0x400, L.cout)
        sc = gamma / np.sqrt(var + 1e-3)
        scale[L.c_off:L.c_off + L.cout] = sc.astype(np.float32)
        shift[L.c_off:L.c_off + L.cout] = (beta - mean * sc).astype(np.float32)
    return scale, shift


def kat_ints(seed: int, count: int, lo: int, hi: int) -> np.ndarray:
    """Small integers in [lo, hi] for the bit-exact known-answer tests."""
    w = splitmix64(seed, np.arange(count, dtype=np.uint64))
    return lo + ((w >> np.uint64(33)) % np.uint64(hi - lo + 1)).astype(np.int64)


def _round_bf16(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def bf16_storage_weights(w: np.ndarray, in_scale: float = 1.0 / 127.5) -> np.ndarray:
    """The filter values a bf16 context actually multiplies with (csrc/api.cu, stem_tc.cu):
    pointwise and FC filters are stored in bf16; the stem's taps are folded with the input scale
    and stored in fp16 (returned here divided by the scale again); depthwise taps stay fp32."""
    from .layers import LAYERS, POINTWISE, FC, STEM
    out = np.array(w, dtype=np.float32, copy=True)
    for L in LAYERS:
        sl = slice(L.w_off, L.w_off + L.w_cnt)
        if L.kind in (POINTWISE, FC):
            out[sl] = _round_bf16(out[sl])
        elif L.kind == STEM:
            s = np.float32(in_scale)
            out[sl] = (out[sl] * s).astype(np.float16).astype(np.float32) / s
    return out
