"""Keras `mobilenet_1_0_224_tf.h5` -> the weight file `mnv1_load_weights` reads.

The reference's exporter (`keras.py:1-8`) opens the HDF5 container as raw bytes, reinterprets them as
float64 and writes them out: it never reads a tensor.  This module is what it was meant to be, with no
dependency the image lacks (no h5py / libhdf5 here): a reader for the subset of HDF5 that h5py writes
for Keras weight files —

  * superblock version 0 / 1 (optionally behind a user block), 8-byte offsets and lengths,
  * "old style" groups: symbol-table message -> v1 B-tree (TREE) -> symbol-table nodes (SNOD) with names
    in a local heap (HEAP),
  * version-1 object headers with continuation blocks,
  * datasets with a simple dataspace, little-endian IEEE float / integer types and CONTIGUOUS or
    COMPACT layout (h5py writes un-chunked datasets for `create_dataset(data=...)`, which is what
    Keras' `save_weights` does); chunked / filtered datasets raise `NotImplementedError`,

— and the transform of SURVEY App. D: Keras layer names -> the reference's layer order, HWIO ->
the reference's flat filter orders (`MobileNet.c` App. A), BatchNorm folded to per-channel
scale / shift, FC bias in the last 1000 shift entries.

Pinning: the HDF5 reader is checked against a file written by libhdf5 itself
(`tests/golden/libhdf5_sample.mat`, a MATLAB v7.3 file from scipy's test data: same superblock, group and
object-header structures), and the whole exporter against a synthetic Keras-layout file produced by
`tests/h5_writer.py`.  The real `mobilenet_1_0_224_tf.h5` is absent from the mount
(`.MISSING_LARGE_BLOBS`), so the name mapping follows App. D and is unverified against the blob.
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class H5File:
    """Read-only view of an HDF5 file (the subset described in the module docstring)."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.base = self._find_superblock()
        self.root_header = self._parse_superblock()

    # ---------------------------------------------------------------- low level
    def _u(self, off: int, size: int) -> int:
        return int.from_bytes(self.buf[off:off + size], "little")

    def _addr(self, rel: int) -> int:
        if rel == UNDEF:
            raise H5Error("undefined address")
        return self.base + rel

    def _find_superblock(self) -> int:
        off = 0
        while off + 8 <= len(self.buf):          # the signature sits at 0, 512, 1024, ... (user block)
            if self.buf[off:off + 8] == SIGNATURE:
                return off
            off = 512 if off == 0 else off * 2
        raise H5Error("not an HDF5 file: signature not found")

    def _parse_superblock(self) -> int:
        o = self.base + 8
        version = self.buf[o]
        if version > 1:
            raise H5Error(f"superblock version {version} is not supported (h5py's default writes 0)")
        if self.buf[o + 5] != 8 or self.buf[o + 6] != 8:
            raise H5Error("only 8-byte offsets / lengths are supported")
        o += 8 + 2 + 2 + 4                        # versions + sizes, group leaf / internal K, consistency flags
        if version == 1:
            o += 4                                # indexed-storage K + reserved
        base_rel = self._u(o, 8)
        if base_rel not in (0, self.base):        # libhdf5 stores the user-block size here
            raise H5Error("unexpected base address")
        o += 8 * 4                                # base, free-space info, end of file, driver info
        return self._u(o + 8, 8)                  # root symbol-table entry: link name offset, object header address

    # ---------------------------------------------------------------- object headers
    def _messages(self, header_rel: int):
        """Yield (type, flags, payload offset, payload size) of a version-1 object header."""
        o = self._addr(header_rel)
        if self.buf[o] != 1:
            raise H5Error(f"object header version {self.buf[o]} is not supported (needs libver='earliest')")
        nmsg = self._u(o + 2, 2)
        blocks = [(o + 16, self._u(o + 8, 4))]    # 12 bytes of prefix padded to 16
        seen = 0
        while blocks and seen < nmsg:
            start, size = blocks.pop(0)
            p, end = start, start + size
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                body = p + 8
                seen += 1
                if mtype == 0x10:                 # continuation: offset, length
                    blocks.append((self._addr(self._u(body, 8)), self._u(body + 8, 8)))
                else:
                    yield mtype, mflags, body, msize
                p = body + msize

    def _heap_name(self, heap_rel: int, name_off: int) -> str:
        h = self._addr(heap_rel)
        if self.buf[h:h + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        data = self._addr(self._u(h + 24, 8))
        end = self.buf.index(b"\0", data + name_off)
        return self.buf[data + name_off:end].decode("utf-8")

    def _group_entries(self, btree_rel: int, heap_rel: int):
        """Yield (name, object header address) for every link under a v1 group B-tree."""
        n = self._addr(btree_rel)
        sig = self.buf[n:n + 4]
        if sig == b"SNOD":
            count = self._u(n + 6, 2)
            for i in range(count):
                e = n + 8 + 40 * i
                yield self._heap_name(heap_rel, self._u(e, 8)), self._u(e + 8, 8)
            return
        if sig != b"TREE" or self.buf[n + 4] != 0:
            raise H5Error("group B-tree node expected")
        used = self._u(n + 6, 2)
        p = n + 24 + 8                            # header, then key 0
        for _ in range(used):
            child = self._u(p, 8)
            yield from self._group_entries(child, heap_rel)
            p += 16                               # child address + next key

    def _children(self, header_rel: int):
        for mtype, _f, body, _s in self._messages(header_rel):
            if mtype == 0x11:                     # symbol table: B-tree, local heap
                return list(self._group_entries(self._u(body, 8), self._u(body + 8, 8)))
        return None                               # not a group

    # ---------------------------------------------------------------- datasets
    @staticmethod
    def _dtype(cls_ver: int, bits0: int, size: int) -> np.dtype:
        cls = cls_ver & 0x0F
        if bits0 & 1:
            raise H5Error("big-endian data is not supported")
        if cls == 1 and size in (2, 4, 8):
            return np.dtype(f"<f{size}")
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(f"<{'i' if bits0 & 8 else 'u'}{size}")
        raise H5Error(f"datatype class {cls} of size {size} is not supported")

    def _dataset(self, header_rel: int) -> np.ndarray | None:
        shape = dtype = None
        data = None
        for mtype, _f, body, size in self._messages(header_rel):
            if mtype == 0x01:                     # dataspace
                ver, rank = self.buf[body], self.buf[body + 1]
                dims = body + (8 if ver == 1 else 4)
                shape = tuple(self._u(dims + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:                   # datatype
                dtype = self._dtype(self.buf[body], self.buf[body + 1], self._u(body + 4, 4))
            elif mtype == 0x08:                   # data layout
                ver = self.buf[body]
                if ver == 3:
                    cls = self.buf[body + 1]
                    if cls == 1:
                        data = ("contiguous", self._u(body + 2, 8), self._u(body + 10, 8))
                    elif cls == 0:
                        data = ("compact", body + 4, self._u(body + 2, 2))
                    else:
                        raise NotImplementedError("chunked datasets are not supported (Keras weight files are contiguous)")
                elif ver in (1, 2):
                    rank, cls = self.buf[body + 1], self.buf[body + 2]
                    if cls != 1:
                        raise NotImplementedError("only contiguous datasets are supported for layout versions 1 / 2")
                    data = ("contiguous", self._u(body + 8, 8), None)
                else:
                    raise H5Error(f"data layout version {ver}")
        if shape is None or dtype is None or data is None:
            return None
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = count * dtype.itemsize
        if data[0] == "compact":
            off = data[1]
        else:
            if data[1] == UNDEF:                  # never written: fill value (zeros)
                return np.zeros(shape, dtype)
            off = self._addr(data[1])
        if off + nbytes > len(self.buf):
            raise H5Error("dataset extends past the end of the file")
        return np.frombuffer(self.buf, dtype, count, off).reshape(shape).copy()

    # ---------------------------------------------------------------- public
    def datasets(self) -> dict[str, np.ndarray]:
        """Every dataset of the file, keyed by its full path without the leading '/'."""
        out: dict[str, np.ndarray] = {}
        stack = [("", self.root_header)]
        visited = set()
        while stack:
            prefix, hdr = stack.pop()
            if hdr in visited:
                continue
            visited.add(hdr)
            kids = self._children(hdr)
            if kids is None:
                arr = self._dataset(hdr)
                if arr is not None:
                    out[prefix] = arr
                continue
            for name, child in kids:
                stack.append((f"{prefix}/{name}" if prefix else name, child))
        return out


# -------------------------------------------------------------------- Keras MobileNet -> reference order
BN_EPS = 1e-3
# (cin, cout) of the 13 depthwise-separable blocks, SURVEY App. B
_PLAN = [(32, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512)] + [(512, 512)] * 5 + [(512, 1024), (1024, 1024)]


def _find(ds: dict[str, np.ndarray], layer: str, var: str) -> np.ndarray:
    """Keras stores `<layer>/<layer>/<var>:0` (model.save_weights) or `model_weights/<layer>/<layer>/<var>:0`."""
    suffix = f"{layer}/{var}:0"
    hits = [k for k in ds if k == suffix or k.endswith("/" + suffix)]
    if not hits:
        raise H5Error(f"dataset {suffix} not found")
    return ds[min(hits, key=len)]


def _fold(ds, layer: str, c: int):
    g, b = _find(ds, layer, "gamma"), _find(ds, layer, "beta")
    m, v = _find(ds, layer, "moving_mean"), _find(ds, layer, "moving_variance")
    for a in (g, b, m, v):
        if a.shape != (c,):
            raise H5Error(f"{layer}: BatchNorm vector of shape {a.shape}, expected ({c},)")
    scale = (g.astype(np.float64) / np.sqrt(v.astype(np.float64) + BN_EPS))
    shift = b.astype(np.float64) - m.astype(np.float64) * scale
    return scale.astype(np.float32), shift.astype(np.float32)


def mobilenet_from_datasets(ds: dict[str, np.ndarray]):
    """(weights[4 209 088], scale[11 944], shift[11 944]) in the order `mnv1_set_weights` takes
    (include/mnv1.h; filters per layer in the reference's `readSquezeNetKernel` order, App. A)."""
    w, sc, sh = [], [], []

    def take(kernel: np.ndarray, want: tuple, perm: tuple, bn: str, c: int):
        if kernel.shape != want:
            raise H5Error(f"{bn}: kernel of shape {kernel.shape}, expected {want}")
        w.append(np.ascontiguousarray(kernel.transpose(perm), dtype=np.float32).reshape(-1))
        s, t = _fold(ds, bn, c)
        sc.append(s); sh.append(t)

    take(_find(ds, "conv1", "kernel"), (3, 3, 3, 32), (3, 2, 0, 1), "conv1_bn", 32)            # HWIO -> [O][I][H][W]
    for i, (cin, cout) in enumerate(_PLAN, start=1):
        take(_find(ds, f"conv_dw_{i}", "depthwise_kernel"), (3, 3, cin, 1), (2, 3, 0, 1), f"conv_dw_{i}_bn", cin)   # -> [C][1][3][3]
        take(_find(ds, f"conv_pw_{i}", "kernel"), (1, 1, cin, cout), (3, 2, 0, 1), f"conv_pw_{i}_bn", cout)        # -> [O][I]
    fc = _find(ds, "conv_preds", "kernel")
    if fc.shape != (1, 1, 1024, 1000):
        raise H5Error(f"conv_preds: kernel of shape {fc.shape}")
    w.append(np.ascontiguousarray(fc.transpose(3, 2, 0, 1), dtype=np.float32).reshape(-1))       # -> [1000][1024]
    bias = _find(ds, "conv_preds", "bias").astype(np.float32)
    sc.append(np.ones(1000, np.float32)); sh.append(bias)
    weights, scale, shift = np.concatenate(w), np.concatenate(sc), np.concatenate(sh)
    assert weights.size == 4209088 and scale.size == 10944 + 1000 == shift.size
    return weights, scale, shift


def save_weights_bin(path: str, weights: np.ndarray, scale: np.ndarray, shift: np.ndarray) -> None:
    """The "MNV1WTS1" container of csrc/weights_io.cpp: magic | u64 n_weights | u64 n_channels | f32 arrays."""
    with open(path, "wb") as f:
        f.write(b"MNV1WTS1" + struct.pack("<QQ", weights.size, scale.size))
        f.write(weights.astype("<f4").tobytes() + scale.astype("<f4").tobytes() + shift.astype("<f4").tobytes())


def export(h5_path: str, out_path: str) -> None:
    """`keras.py`'s job: mobilenet_1_0_224_tf.h5 -> a file `mnv1_load_weights` (TF-SAME padding, ReLU6) reads."""
    save_weights_bin(out_path, *mobilenet_from_datasets(H5File(h5_path).datasets()))
