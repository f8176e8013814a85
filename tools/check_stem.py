import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mnv1_b200
from mnv1_b200 import binding as mn, synth
import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rng = np.random.default_rng(30)
img = synth.images(n)
w = (rng.standard_normal((32, 3, 3, 3)) * np.sqrt(2.0 / 27)).astype(np.float32)
wo = ((w * np.float32(1 / 127.5)).astype(np.float16).astype(np.float32) / np.float32(1 / 127.5))
want = oracle.convolute(img, img.reshape(-1)[1:], img.reshape(-1)[2:], wo, n, 224, 224, 2, 32, pad_mode=1, in_scale=1 / 127.5,
                        in_bias=-1.0, act=oracle.ACT_RELU6, rbf16=True, pix_stride=3, img_stride=224 * 224 * 3)
ctx = mn.Context(0, mn.BF16)
ctx.set_pad_mode(1); ctx.set_input_transform(1 / 127.5, -1.0)
f = ctx.filter(mn.CONVOLUTE, w, 3, 32, None, None, mn.ACT_RELU6)
out = ctx.malloc(n, 32, 112, 112)
rgb = ctx.upload_u8(img)
ctx.convolute_rgb(out, rgb, f, 224, 224, 3, 2, 32)
got = ctx.download_planar(out)
err = np.abs(got - want) / np.maximum(1, np.abs(want))
bad = err > 2 ** -7
print(ctx.last_kernel_name, "bad fraction", bad.mean())
if bad.any():
    idx = np.argwhere(bad.any(axis=1))  # (img, y, x)
    print("first bad pixels", idx[:10].tolist(), "last", idx[-3:].tolist(), "count", len(idx))
    m = idx[:, 0] * 12544 + idx[:, 1] * 112 + idx[:, 2]
    tiles = np.unique(m // 128)
    print("bad tiles", tiles[:20], len(tiles), "of", n * 98, "tile mod 592:", np.unique(tiles % 592)[:10], "tile//592", np.unique(tiles // 592))
