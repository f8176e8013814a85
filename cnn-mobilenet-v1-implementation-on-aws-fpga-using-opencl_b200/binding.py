"""ctypes binding of ``libmnv1.so`` (include/mnv1.h) — the same calls a C host makes.

There is no fallback: if the shared library is missing or no sm_100 GPU is visible the
constructors raise.  Tests and bench go through this module, i.e. through the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from .layers import LAYERS, NUM_CLASSES, TOTAL_CHANNELS, TOTAL_WEIGHTS

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MNV1_LIB", os.path.join(PKG_DIR, "libmnv1.so"))  # MNV1_LIB: kernel experiments only
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "mnv1.h")

F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
PAD_REF, PAD_TFSAME = 0, 1
CONVOLUTE, DEPTHWISE, POINTWISE, POOL, FC = 0, 1, 2, 3, 4


class Mnv1Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mnv1 error {code}: {msg}")
        self.code = code


def build_library(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libmnv1.so (in-tree)."""
    r = subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libmnv1.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with __graft_entry__.build() or `make -C {PKG_DIR}/csrc` "
                "(there is no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        L.mnv1_last_error.restype = C.c_char_p
        L.mnv1_last_error.argtypes = [C.c_void_p]
        L.mnv1_version.restype = C.c_char_p
        L.mnv1_last_kernel_name.restype = C.c_char_p
        L.mnv1_last_kernel_name.argtypes = [C.c_void_p]
        L.mnv1_launch_count.restype = C.c_long
        L.mnv1_launch_count.argtypes = [C.c_void_p]
        L.mnv1_buf_device_ptr.restype = C.c_void_p
        L.mnv1_buf_device_ptr.argtypes = [C.c_void_p]
        L.mnv1_dp_last_error.restype = C.c_char_p
        L.mnv1_dp_last_error.argtypes = [C.c_void_p]
        L.mnv1_dp_ctx.restype = C.c_void_p
        L.mnv1_dp_ctx.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _vp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> Optional[np.ndarray]:
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Buffer:
    def __init__(self, ctx: "Context", handle, shape=None, nbytes=0):
        self.ctx, self.h, self.shape, self.nbytes = ctx, handle, shape, nbytes

    @property
    def device_ptr(self) -> int:
        return lib().mnv1_buf_device_ptr(self.h) or 0

    def free(self):
        if self.h:
            lib().mnv1_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Filter:
    def __init__(self, ctx: "Context", handle, kind, cin, cout):
        self.ctx, self.h, self.kind, self.cin, self.cout = ctx, handle, kind, cin, cout

    def free(self):
        if self.h:
            lib().mnv1_filter_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One per GPU (replaces cl_context + cl_command_queue, MobileNet.c:147-205)."""

    def __init__(self, device: int = 0, dtype: int = F32):
        self.h = C.c_void_p()
        self.dtype = dtype
        rc = lib().mnv1_ctx_create(int(device), int(dtype), C.byref(self.h))
        if rc:
            raise Mnv1Error(rc, (lib().mnv1_last_error(None) or b"").decode())

    def _ck(self, rc: int):
        if rc:
            raise Mnv1Error(rc, (lib().mnv1_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            lib().mnv1_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- settings
    def set_stream(self, cuda_stream: int):
        self._ck(lib().mnv1_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_pad_mode(self, pad: int):
        self._ck(lib().mnv1_ctx_set_pad_mode(self.h, int(pad)))

    def set_input_transform(self, scale: float, bias: float):
        self._ck(lib().mnv1_ctx_set_input_transform(self.h, C.c_float(scale), C.c_float(bias)))

    def use_graph(self, on: bool):
        self._ck(lib().mnv1_ctx_use_graph(self.h, int(on)))

    def sync(self):
        self._ck(lib().mnv1_sync(self.h))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._ck(lib().mnv1_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    @property
    def last_kernel_name(self) -> str:
        return (lib().mnv1_last_kernel_name(self.h) or b"").decode()

    @property
    def launch_count(self) -> int:
        return lib().mnv1_launch_count(self.h)

    # ---- buffers
    def malloc(self, n, c, h, w) -> Buffer:
        b = C.c_void_p()
        self._ck(lib().mnv1_malloc(self.h, n, c, h, w, C.byref(b)))
        return Buffer(self, b, (n, c, h, w))

    def malloc_u8(self, nbytes: int) -> Buffer:
        b = C.c_void_p()
        self._ck(lib().mnv1_malloc_u8(self.h, C.c_size_t(nbytes), C.byref(b)))
        return Buffer(self, b, None, nbytes)

    def upload_u8(self, arr: np.ndarray) -> Buffer:
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        b = self.malloc_u8(arr.nbytes)
        self._ck(lib().mnv1_upload_u8(self.h, b.h, _vp(arr), C.c_size_t(arr.nbytes)))
        return b

    def upload_planar(self, arr: np.ndarray) -> Buffer:
        arr = _f32(arr)
        n, c, h, w = arr.shape
        b = self.malloc(n, c, h, w)
        self._ck(lib().mnv1_upload_planar(self.h, b.h, _vp(arr)))
        return b

    def upload_planar_u8(self, arr: np.ndarray) -> Buffer:
        """integer contexts: planar u8 host array, the reference's own layout (MobileNet.c:116-143)"""
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        n, c, h, w = arr.shape
        b = self.malloc(n, c, h, w)
        self._ck(lib().mnv1_upload_planar_u8(self.h, b.h, _vp(arr)))
        return b

    def download_planar_u8(self, b: Buffer) -> np.ndarray:
        out = np.empty(b.shape, dtype=np.uint8)
        self._ck(lib().mnv1_download_planar_u8(self.h, b.h, _vp(out)))
        return out

    def set_u8_store(self, wrap: bool):
        self._ck(lib().mnv1_ctx_set_u8_store(self.h, int(wrap)))

    def download_planar(self, b: Buffer) -> np.ndarray:
        out = np.empty(b.shape, dtype=np.float32)
        self._ck(lib().mnv1_download_planar(self.h, b.h, _vp(out)))
        return out

    # ---- filters
    def filter(self, kind, w, cin, cout, scale=None, shift=None, act=ACT_RELU) -> Filter:
        w, scale, shift = _f32(w), _f32(scale), _f32(shift)
        f = C.c_void_p()
        self._ck(lib().mnv1_filter_create(self.h, int(kind), _vp(w), cin, cout, _vp(scale), _vp(shift), int(act),
                                          C.byref(f)))
        return Filter(self, f, kind, cin, cout)

    # ---- the four kernels, kernel.cl argument order
    def convolute(self, out: Buffer, r: Buffer, g: Buffer, b: Buffer, f: Filter, rows, cols, filtersize, stride,
                  op_size):
        self._ck(lib().mnv1_convolute(self.h, out.h, r.h, g.h, b.h, f.h, rows, cols, filtersize, stride, op_size))

    def convolute_rgb(self, out: Buffer, rgb: Buffer, f: Filter, rows, cols, filtersize, stride, op_size):
        self._ck(lib().mnv1_convolute_rgb(self.h, out.h, rgb.h, f.h, rows, cols, filtersize, stride, op_size))

    def depthwise(self, out: Buffer, inp: Buffer, f: Filter, rows, cols, filtersize, stride, op_size):
        self._ck(lib().mnv1_depthwise(self.h, out.h, inp.h, f.h, rows, cols, filtersize, stride, op_size))

    def pointwise(self, out: Buffer, inp: Buffer, f: Filter, rows, cols, filtersize, op_size, simt=False):
        fn = lib().mnv1_pointwise_simt if simt else lib().mnv1_pointwise
        self._ck(fn(self.h, out.h, inp.h, f.h, rows, cols, filtersize, op_size))

    def dw_pw_block(self, out: Buffer, inp: Buffer, dw: Filter, pw: Filter, rows, cols, stride):
        self._ck(lib().mnv1_dw_pw_block(self.h, out.h, inp.h, dw.h, pw.h, rows, cols, stride))

    def use_fused_blocks(self, on: bool):
        self._ck(lib().mnv1_ctx_use_fused_blocks(self.h, int(on)))

    def fused_layers(self) -> np.ndarray:
        """fused[i] = 1: mnv1_forward runs depthwise layer i+1 and the pointwise after it as one kernel."""
        f = np.zeros(29, dtype=np.int32)
        self._ck(lib().mnv1_fused_layers(self.h, _vp(f)))
        return f

    def pool(self, out: Buffer, inp: Buffer, rows, cols, filtersize, op_size):
        self._ck(lib().mnv1_pool(self.h, out.h, inp.h, rows, cols, filtersize, op_size))

    def softmax(self, logits: Buffer, classes: int):
        n = logits.shape[0]
        prob = np.empty((n, classes), dtype=np.float32)
        top1 = np.empty(n, dtype=np.int32)
        p1 = np.empty(n, dtype=np.float32)
        self._ck(lib().mnv1_softmax(self.h, logits.h, classes, _vp(prob), _vp(top1), _vp(p1)))
        return prob, top1, p1

    # ---- whole network
    def set_weights(self, weights, scale=None, shift=None, act=ACT_RELU6):
        weights, scale, shift = _f32(weights), _f32(scale), _f32(shift)
        assert weights.size == TOTAL_WEIGHTS
        assert scale is None or scale.size == TOTAL_CHANNELS
        assert shift is None or shift.size == TOTAL_CHANNELS
        self._ck(lib().mnv1_set_weights(self.h, _vp(weights), _vp(scale), _vp(shift), int(act)))

    def load_weights(self, path: str, act=ACT_RELU6):
        self._ck(lib().mnv1_load_weights(self.h, path.encode(), int(act)))

    def plan(self, max_batch: int):
        self._ck(lib().mnv1_plan(self.h, int(max_batch)))
        self.planned_batch = max(int(max_batch), getattr(self, "planned_batch", 0))

    def forward(self, images_u8: np.ndarray, want_logits=True):
        """Host in / host out (H2D + 28 kernels + D2H inside)."""
        images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
        n = images_u8.shape[0]
        logits = np.empty((n, NUM_CLASSES), dtype=np.float32) if want_logits else None
        top1 = np.empty(n, dtype=np.int32)
        p1 = np.empty(n, dtype=np.float32)
        self._ck(lib().mnv1_forward(self.h, _vp(images_u8), n, _vp(logits), _vp(top1), _vp(p1)))
        return logits, top1, p1

    def forward_raw(self, images_ptr: int, n: int, logits_ptr: int, top1_ptr: int, prob_ptr: int):
        """mnv1_forward with caller-owned (ideally pinned) host pointers."""
        self._ck(lib().mnv1_forward(self.h, C.c_void_p(images_ptr), n, C.c_void_p(logits_ptr), C.c_void_p(top1_ptr),
                                    C.c_void_p(prob_ptr)))

    def forward_submit(self, images_ptr: int, n: int, logits_ptr: int, top1_ptr: int, prob_ptr: int) -> int:
        """Pipelined mnv1_forward: returns a ticket; up to three batches in flight."""
        t = C.c_long(-1)
        self._ck(lib().mnv1_forward_submit(self.h, C.c_void_p(images_ptr), n, C.c_void_p(logits_ptr or None),
                                           C.c_void_p(top1_ptr or None), C.c_void_p(prob_ptr or None), C.byref(t)))
        return t.value

    def forward_wait(self, ticket: int):
        self._ck(lib().mnv1_forward_wait(self.h, C.c_long(ticket)))

    def forward_device(self, d_images: int, n: int, d_logits: int, d_top1: int = 0, d_prob: int = 0):
        self._ck(lib().mnv1_forward_device(self.h, C.c_void_p(d_images), n, C.c_void_p(d_logits),
                                           C.c_void_p(d_top1 or None), C.c_void_p(d_prob or None)))

    def forward_upto(self, images_u8: np.ndarray, last_layer: int) -> np.ndarray:
        images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
        n = images_u8.shape[0]
        L = LAYERS[last_layer - 1]
        out = np.empty((n, L.cout, L.hout, L.hout), dtype=np.float32)
        self._ck(lib().mnv1_forward_upto(self.h, _vp(images_u8), n, last_layer, _vp(out)))
        return out.reshape(n, L.cout) if L.hout == 1 else out

    def profile_layers(self, d_images: int, n: int, iters: int = 5) -> np.ndarray:
        t = np.zeros(29, dtype=np.float32)
        self._ck(lib().mnv1_profile_layers(self.h, C.c_void_p(d_images), n, iters, _vp(t)))
        return t

    # ---- logits gather of the data-parallel mode (peer stores from the head kernel)
    def gather_create(self, world: int, rank: int, rows_per_rank: int):
        self._ck(lib().mnv1_gather_create(self.h, world, rank, rows_per_rank))
        self._gather = True

    def gather_set_rows(self, first_row: int, max_rows: int):
        self._ck(lib().mnv1_gather_set_rows(self.h, C.c_long(first_row), int(max_rows)))

    def gather_active(self) -> bool:
        return getattr(self, "_gather", False)

    def gather_attach(self, peer: "Context"):
        self._ck(lib().mnv1_gather_attach(self.h, peer.h))

    def gather_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(lib().mnv1_gather_export(self.h, buf))
        return buf.raw

    def gather_import(self, peer_rank: int, handle: bytes):
        self._ck(lib().mnv1_gather_import(self.h, int(peer_rank), C.c_char_p(handle)))

    def gather_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(lib().mnv1_gather_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def gather_destroy(self):
        self._ck(lib().mnv1_gather_destroy(self.h))
        self._gather = False

    def h2d_probe_open(self, nbytes: int):
        p = C.c_void_p()
        self._ck(lib().mnv1_h2d_probe_open(self.h, C.c_size_t(nbytes), C.byref(p)))
        return p

    def h2d_probe_run(self, probe, reps: int) -> float:
        g = C.c_float()
        self._ck(lib().mnv1_h2d_probe_run(probe, int(reps), C.byref(g)))
        return g.value

    def h2d_probe_close(self, probe):
        lib().mnv1_h2d_probe_close(probe)

    def h2d_probe(self, nbytes: int, reps: int = 20) -> float:
        g = C.c_float()
        self._ck(lib().mnv1_h2d_probe(self.h, C.c_size_t(nbytes), reps, C.byref(g)))
        return g.value

    def forward_prefix_device(self, d_images: int, n: int, last_layer: int):
        self._ck(lib().mnv1_forward_prefix_device(self.h, C.c_void_p(d_images), n, last_layer))

    def profile_prefixes(self, d_images: int, n: int, iters: int = 21) -> np.ndarray:
        """cum_ms[k-1]: median replay time of the graph of layers 1..k (-1 where no launch ends at layer k)."""
        t = np.zeros(29, dtype=np.float32)
        self._ck(lib().mnv1_profile_prefixes(self.h, C.c_void_p(d_images), n, iters, _vp(t)))
        return t

    def synth_images_device(self, d_images: int, n: int, first: int, seed: int):
        self._ck(lib().mnv1_synth_images_device(self.h, C.c_void_p(d_images), n, C.c_long(first), C.c_uint64(seed)))


class DataParallel:
    """mnv1_dp_*: one process driving several GPUs (a context + worker thread per device, batch cut into
    contiguous shards, logits gathered by peer stores of the head kernel)."""

    def __init__(self, devices: Sequence[int], dtype: int = BF16, max_batch_per_gpu: int = 256):
        self.h = C.c_void_p()
        self.devices = list(devices)
        self.rows = int(max_batch_per_gpu)
        arr = (C.c_int * len(self.devices))(*self.devices)
        rc = lib().mnv1_dp_create(arr, len(self.devices), int(dtype), self.rows, C.byref(self.h))
        if rc:
            raise Mnv1Error(rc, (lib().mnv1_dp_last_error(None) or b"").decode())

    def _ck(self, rc: int):
        if rc:
            raise Mnv1Error(rc, (lib().mnv1_dp_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            lib().mnv1_dp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self) -> int:
        return lib().mnv1_dp_size(self.h)

    def ctx_handle(self, rank: int) -> int:
        return lib().mnv1_dp_ctx(self.h, rank)

    def set_pad_mode(self, pad: int):
        self._ck(lib().mnv1_dp_set_pad_mode(self.h, int(pad)))

    def set_input_transform(self, scale: float, bias: float):
        self._ck(lib().mnv1_dp_set_input_transform(self.h, C.c_float(scale), C.c_float(bias)))

    def set_weights(self, weights, scale=None, shift=None, act=ACT_RELU6):
        weights, scale, shift = _f32(weights), _f32(scale), _f32(shift)
        self._ck(lib().mnv1_dp_set_weights(self.h, _vp(weights), _vp(scale), _vp(shift), int(act)))

    def forward(self, images_u8: np.ndarray):
        images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
        n = images_u8.shape[0]
        logits = np.empty((n, NUM_CLASSES), dtype=np.float32)
        top1 = np.empty(n, dtype=np.int32)
        p1 = np.empty(n, dtype=np.float32)
        self._ck(lib().mnv1_dp_forward(self.h, _vp(images_u8), n, _vp(logits), _vp(top1), _vp(p1)))
        return logits, top1, p1

    def forward_submit(self, images_ptr: int, n: int, logits_ptr: int, top1_ptr: int, prob_ptr: int) -> int:
        t = C.c_long(-1)
        self._ck(lib().mnv1_dp_forward_submit(self.h, C.c_void_p(images_ptr), n, C.c_void_p(logits_ptr or None),
                                              C.c_void_p(top1_ptr or None), C.c_void_p(prob_ptr or None), C.byref(t)))
        return t.value

    def forward_wait(self, ticket: int):
        self._ck(lib().mnv1_dp_forward_wait(self.h, C.c_long(ticket)))

    def set_shard_weights(self, weights=None):
        if weights is None:
            self._ck(lib().mnv1_dp_set_shard_weights(self.h, None))
        else:
            arr = (C.c_float * len(weights))(*[float(w) for w in weights])
            self._ck(lib().mnv1_dp_set_shard_weights(self.h, arr))

    def calibrate(self):
        """Shards in proportion to each GPU's pinned H2D rate with all of them copying at once; returns the rates (GB/s)."""
        arr = (C.c_float * self.size)()
        self._ck(lib().mnv1_dp_calibrate(self.h, arr))
        return list(arr)

    def forward_device(self, d_images: Sequence[int], n_per_gpu: int):
        arr = (C.c_void_p * len(d_images))(*d_images)
        self._ck(lib().mnv1_dp_forward_device(self.h, arr, int(n_per_gpu)))

    def gather_ptrs(self, rank: int):
        """(logits, top1, top1_prob) device pointers of rank's gather block: [size * rows] rows each."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = lib().mnv1_gather_ptrs(C.c_void_p(self.ctx_handle(rank)), C.byref(a), C.byref(b), C.byref(c))
        if rc:
            raise Mnv1Error(rc, "gather_ptrs")
        return a.value, b.value, c.value


def dp_shard(n: int, rank: int, world: int):
    first, count = C.c_int(), C.c_int()
    rc = lib().mnv1_dp_shard(n, rank, world, C.byref(first), C.byref(count))
    if rc:
        raise Mnv1Error(rc, "bad shard arguments")
    return first.value, count.value


def dp_shard_weighted(n: int, rank: int, world: int, weights=None):
    first, count = C.c_int(), C.c_int()
    arr = None if weights is None else (C.c_float * len(weights))(*[float(w) for w in weights])
    rc = lib().mnv1_dp_shard_weighted(n, rank, world, arr, C.byref(first), C.byref(count))
    if rc:
        raise Mnv1Error(rc, "bad shard arguments")
    return first.value, count.value


def declared_symbols() -> Sequence[str]:
    """Every function include/mnv1.h declares (used by the symbol-export test)."""
    import re
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnv1_[a-z0-9_]+)\s*\(", src)))
