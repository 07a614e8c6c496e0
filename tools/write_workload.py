"""Writes a synthetic workload file for tools/bench_capi.c:  python tools/write_workload.py OUT.bin PAIRS LENGTH [trim expansion [pairs-per-generator-batch]]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from cpecan_b200 import synth  # noqa: E402

out, n, length = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
trim, expansion = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (14, 20)
kw = {"batch": int(sys.argv[6])} if len(sys.argv) > 6 else {}
packed = synth.evolved_pairs(n, length, seed=0xC0FFEE, trim=trim, expansion=expansion, **kw)
with open(out, "wb") as f:
    np.asarray([n], dtype=np.int64).tofile(f)
    for k in ("xOff", "yOff", "aOff"):
        np.ascontiguousarray(packed[k], dtype=np.int64).tofile(f)
    np.ascontiguousarray(packed["seqX"][: int(packed["xOff"][-1])], dtype=np.uint8).tofile(f)
    np.ascontiguousarray(packed["seqY"][: int(packed["yOff"][-1])], dtype=np.uint8).tofile(f)
    np.ascontiguousarray(packed["anchors"][: 3 * int(packed["aOff"][-1])], dtype=np.int64).tofile(f)
