#!/usr/bin/env python
"""Compare two stdout logs of the MobileNet host programs.

The reference prints one `Kernel Execution time for Layer k: <seconds>` line per layer
(MobileNet.c:315 ... :2763) and a final `Highest Probability of the element is present at location L
and it's value is P.` (MobileNet.c:2792); host/MobileNet*.c on the C-ABI print the same lines.
    python tools/compare_logs.py reference.log ours.log
prints a per-layer table (seconds, speed-up) and whether the predicted class and probability agree.
Exit status: 0 same class (or no final line in either log), 1 different class, 2 unusable logs.
"""
import re
import sys

LAYER = re.compile(r"Kernel Execution time for Layer\s+(\d+)\s*:\s*([0-9.eE+-]+)")
FINAL = re.compile(r"Highest Probability of the element is present at location\s+(-?\d+)\s+and it's value is\s+([0-9.eE+-]+?)\.?\s*$", re.M)


def parse(text):
    """-> ({layer: seconds}, (location, probability) or None); a layer printed twice keeps the last value."""
    layers = {int(k): float(v) for k, v in LAYER.findall(text)}
    m = FINAL.search(text)
    return layers, ((int(m.group(1)), float(m.group(2))) if m else None)


def compare(ref_text, our_text):
    ref, ref_final = parse(ref_text)
    ours, our_final = parse(our_text)
    rows = []
    for k in sorted(set(ref) | set(ours)):
        a, b = ref.get(k), ours.get(k)
        rows.append((k, a, b, (a / b) if a is not None and b not in (None, 0.0) else None))
    same_class = None if ref_final is None or our_final is None else ref_final[0] == our_final[0]
    return {"rows": rows, "ref_total": sum(ref.values()), "our_total": sum(ours.values()), "ref_final": ref_final,
            "our_final": our_final, "same_class": same_class}


def main(argv):
    if len(argv) != 3:
        print(__doc__)
        return 2
    res = compare(open(argv[1]).read(), open(argv[2]).read())
    if not res["rows"]:
        print("no `Kernel Execution time for Layer` lines found")
        return 2
    print(f"{'layer':>5s} {'reference s':>14s} {'ours s':>14s} {'speed-up':>10s}")
    for k, a, b, r in res["rows"]:
        f = lambda v: f"{v:14.6f}" if v is not None else f"{'-':>14s}"
        print(f"{k:5d} {f(a)} {f(b)} {(f'{r:10.1f}' if r is not None else f'{chr(45):>10s}')}")
    print(f"{'sum':>5s} {res['ref_total']:14.6f} {res['our_total']:14.6f} "
          f"{(res['ref_total'] / res['our_total'] if res['our_total'] else float('nan')):10.1f}")
    for name, fin in (("reference", res["ref_final"]), ("ours", res["our_final"])):
        print(f"{name}: " + (f"location {fin[0]}, probability {fin[1]:.6f}" if fin else "no final line"))
    if res["same_class"] is False:
        print("DIFFERENT predicted class")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
