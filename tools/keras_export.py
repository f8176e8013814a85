#!/usr/bin/env python
"""mobilenet_1_0_224_tf.h5 -> weight file for mnv1_load_weights / the host programs (what the reference's
keras.py was meant to do).  usage: keras_export.py <model.h5> <weights.bin> [--list]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mnv1_b200  # noqa: E402,F401
from mnv1_b200 import keras_h5  # noqa: E402


def main(argv):
    if len(argv) < 3:
        raise SystemExit(__doc__)
    ds = keras_h5.H5File(argv[1]).datasets()
    if "--list" in argv:
        for k in sorted(ds):
            print(f"{k:60s} {ds[k].shape} {ds[k].dtype}")
    w, sc, sh = keras_h5.mobilenet_from_datasets(ds)
    keras_h5.save_weights_bin(argv[2], w, sc, sh)
    print(f"{argv[2]}: {w.size} filter values, {sc.size} scale / shift pairs (BatchNorm folded, eps {keras_h5.BN_EPS}); "
          "load with pad mode TF-SAME, input transform x/127.5 - 1, ReLU6")


if __name__ == "__main__":
    main(sys.argv)
