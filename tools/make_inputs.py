#!/usr/bin/env python
"""Writes the two files the host programs read — the reference ships neither
(Cat_Image0.ppm, MobileNet.c:215; weights_c.txt, MobileNet.c:37): the seeded synthetic image as a
P6 PPM and the seeded synthetic network as an MNV1WTS1 binary weight file.
    python tools/make_inputs.py [outdir]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mnv1_b200  # noqa: E402
from mnv1_b200 import binding as mn, synth  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "."
os.makedirs(out, exist_ok=True)
img = synth.images(1)[0]
with open(os.path.join(out, "Cat_Image0.ppm"), "wb") as f:
    f.write(b"P6\n224 224\n255\n" + img.tobytes())
w = synth.weights()
sc, sh = synth.batchnorm()
rc = mn.lib().mnv1_save_weights_bin(os.path.join(out, "weights_c.txt").encode(), w.ctypes.data_as(C.c_void_p),
                                    sc.ctypes.data_as(C.c_void_p), sh.ctypes.data_as(C.c_void_p))
assert rc == 0
print("wrote", out + "/Cat_Image0.ppm", out + "/weights_c.txt")
