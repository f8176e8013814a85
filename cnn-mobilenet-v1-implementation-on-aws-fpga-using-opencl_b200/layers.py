"""Layer schedule of the reference host program.

One row per block of ``MobileNet.c:207-2763``: kernel kind,
channels, spatial size, stride, and where the layer's filter sits in the flat weight
file that ``readSquezeNetKernel`` (``MobileNet.c:31-47``) is meant to walk
(counts at ``MobileNet.c:241,337,419,...,2696``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

STEM, DEPTHWISE, POINTWISE, POOL, FC = 0, 1, 2, 3, 4
KIND_NAMES = {STEM: "convolute", DEPTHWISE: "depthwise", POINTWISE: "pointwise", POOL: "pool", FC: "pointwise(fc)"}

IMG = 224
NUM_CLASSES = 1000

_DW_C = [32, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024]
_DW_S = [1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1]  # L26 is stride 1 (SURVEY App. B note)
_PW_O = [64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024, 1024]


@dataclass(frozen=True)
class Layer:
    index: int  # 1-based, as in the reference's "Layer k" printouts
    kind: int
    cin: int
    cout: int
    hin: int
    hout: int
    stride: int
    w_off: int  # offset in the flat weight file
    w_cnt: int  # number of filter values (MobileNet.c nW column of App. A)
    c_off: int  # offset in the per-channel scale/shift arrays

    @property
    def macs(self) -> int:
        if self.kind == STEM:
            return self.hout * self.hout * self.cout * 27
        if self.kind == DEPTHWISE:
            return self.hout * self.hout * self.cout * 9
        if self.kind in (POINTWISE, FC):
            return self.hout * self.hout * self.cout * self.cin
        return self.hin * self.hin * self.cout  # pool: adds

    @property
    def in_elems(self) -> int:
        return self.cin * self.hin * self.hin

    @property
    def out_elems(self) -> int:
        return self.cout * self.hout * self.hout


def build_layers() -> List[Layer]:
    layers: List[Layer] = []
    w = c = 0
    layers.append(Layer(1, STEM, 3, 32, 224, 112, 2, w, 864, c))
    w += 864
    c += 32
    h = 112
    k = 2
    for b in range(13):
        ch, s, co = _DW_C[b], _DW_S[b], _PW_O[b]
        ho = h // s
        layers.append(Layer(k, DEPTHWISE, ch, ch, h, ho, s, w, ch * 9, c))
        w += ch * 9
        c += ch
        h = ho
        k += 1
        layers.append(Layer(k, POINTWISE, ch, co, h, h, 1, w, ch * co, c))
        w += ch * co
        c += co
        k += 1
    layers.append(Layer(28, POOL, 1024, 1024, 7, 1, 1, w, 0, c))
    layers.append(Layer(29, FC, 1024, 1000, 1, 1, 1, w, 1024 * 1000, c))
    return layers


LAYERS = build_layers()
TOTAL_WEIGHTS = LAYERS[-1].w_off + LAYERS[-1].w_cnt  # 4 209 088
BN_CHANNELS = LAYERS[-1].c_off  # 10 944
TOTAL_CHANNELS = BN_CHANNELS + NUM_CLASSES  # + fc bias
assert TOTAL_WEIGHTS == 4209088 and BN_CHANNELS == 10944
