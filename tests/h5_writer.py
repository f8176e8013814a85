"""Test helper: a minimal HDF5 WRITER for the structures h5py emits for Keras weight files
(superblock v0, symbol-table groups: TREE -> several SNOD + local HEAP, version-1 object headers,
contiguous little-endian float32 datasets).  Used only to build synthetic Keras-layout fixtures for
`mnv1_b200.keras_h5`; the reader itself is pinned on a file libhdf5 wrote (tests/golden/libhdf5_sample.mat).
Layout follows the HDF5 File Format Specification v1.1 (superblock 0, "old style" groups).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4          # libhdf5 default: at most 2K = 8 symbols per SNOD
INTERNAL_K = 16     # at most 2K = 32 children per TREE node


class _Out:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, data: bytes, align: int = 8) -> int:
        while len(self.b) % align:
            self.b.append(0)
        off = len(self.b)
        self.b += data
        return off


def _msg(mtype: int, payload: bytes) -> bytes:
    payload += b"\0" * (-len(payload) % 8)
    return struct.pack("<HHB3x", mtype, len(payload), 0) + payload


def _object_header(messages: list[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _dataset(out: _Out, arr: np.ndarray, split_header: bool) -> int:
    arr = np.ascontiguousarray(arr, dtype="<f4")
    data = out.alloc(arr.tobytes())
    space = _msg(0x01, struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape))
    dtype = _msg(0x03, struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127))
    layout = _msg(0x08, struct.pack("<BBQQ", 3, 1, data, arr.nbytes))
    if not split_header:
        return out.alloc(_object_header([space, dtype, layout]))
    # layout message in a continuation block, as libhdf5 does when a header outgrows its first allocation
    cont = out.alloc(layout)
    hdr = _object_header([space, dtype, _msg(0x10, struct.pack("<QQ", cont, len(layout)))])
    # the continuation message counts as a message; the layout message is one more
    hdr = bytearray(hdr)
    struct.pack_into("<H", hdr, 2, 4)
    return out.alloc(bytes(hdr))


def _group(out: _Out, children: dict[str, int]) -> tuple[int, int, int]:
    """Write heap, SNODs and the B-tree of a group; returns (object header, btree, heap) addresses."""
    names = sorted(children)
    seg = bytearray(b"\0" * 8)                       # offset 0: the empty name
    name_off = {}
    for n in names:
        name_off[n] = len(seg)
        seg += n.encode() + b"\0"
        seg += b"\0" * (-len(seg) % 8)
    seg_addr = out.alloc(bytes(seg))
    heap = out.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, seg_addr))
    snods, keys = [], [0]
    for i in range(0, len(names), 2 * LEAF_K):
        part = names[i:i + 2 * LEAF_K]
        ent = b"".join(struct.pack("<QQII16x", name_off[n], children[n], 0, 0) for n in part)
        snods.append(out.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)) + ent))
        keys.append(name_off[part[-1]])
    if len(snods) > 2 * INTERNAL_K:
        raise ValueError("too many children for a single-level B-tree")
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
    for i, s in enumerate(snods):
        node += struct.pack("<QQ", keys[i], s)
    node += struct.pack("<Q", keys[len(snods)] if snods else 0)
    btree = out.alloc(node)
    hdr = out.alloc(_object_header([_msg(0x11, struct.pack("<QQ", btree, heap))]))
    return hdr, btree, heap


def write_h5(path: str, tree: dict, user_block: int = 0) -> None:
    """tree: nested dict; leaves are numpy arrays (datasets), dicts are groups."""
    out = _Out()
    out.alloc(b"\0" * 96)                             # superblock placeholder
    counter = [0]

    def emit(node) -> int:
        if isinstance(node, dict):
            return _group(out, {k: emit(v) for k, v in node.items()})[0]
        counter[0] += 1
        return _dataset(out, node, split_header=counter[0] % 5 == 0)

    kids = {k: emit(v) for k, v in tree.items()}
    root_hdr, btree, heap = _group(out, kids)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", user_block, UNDEF, len(out.b), UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    out.b[0:96] = sb
    with open(path, "wb") as f:
        f.write(b"\0" * user_block + bytes(out.b))
