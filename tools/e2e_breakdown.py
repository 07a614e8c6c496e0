"""Where the end-to-end time of one batch goes: create (H2D), run, fetch (D2H), close."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import cpecan_b200 as cp
from cpecan_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = cp.pairwiseAlignmentBandingParameters_construct()
packed = synth.evolved_pairs(n, 1000, seed=0xC0FFEE, trim=int(p.constraintDiagonalTrim), expansion=int(p.diagonalExpansion))
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = cp.Context(0, stream=stream.cuda_stream)
model = cp.stateMachine5_construct(cp.fiveState)
pinned = {}
keep = []
for k in ("seqX", "xOff", "seqY", "yOff", "anchors", "aOff"):
    t = torch.from_numpy(packed[k]).pin_memory()
    keep.append(t)
    pinned[k] = t.numpy()
out = None
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = cp.Batch(ctx, None, None, packed=pinned)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    b.run(model, p, cp.MODE_ALIGNED_PAIRS)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if out is None:
        out = torch.empty((b.result_count(0) + 1024, 3), dtype=torch.int32).pin_memory()
    off, tri = b.fetch_pairs(0, out=out.numpy())
    torch.cuda.synchronize(); t3 = time.perf_counter()
    b.close()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    st = None
    print("rep %d: create %.0f ms, run %.0f ms, fetch %.0f ms, close %.0f ms, total %.0f ms, triples %d" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0), off[-1]))
