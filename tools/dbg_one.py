import sys, os
sys.path.insert(0, '/root/repo')
import cpecan_b200 as cp
from cpecan_b200 import synth
ctx = cp.Context(0)
pc = cp.pairwiseAlignmentBandingParameters_construct()
if len(sys.argv) > 1 and sys.argv[1] == 'narrow':
    pc.constraintDiagonalTrim, pc.diagonalExpansion, pc.splitMatrixBiggerThanThis = 0, 4, 1 << 40
    packed = synth.evolved_pairs(1, 1000, seed=1, trim=0, expansion=4)
else:
    packed = synth.evolved_pairs(1, 1000, seed=1, trim=14, expansion=20)
b = cp.Batch(ctx, None, None, packed=packed)
for i in range(3):
    b.run(cp.stateMachine5_construct(), pc, cp.MODE_ALIGNED_PAIRS)
print(b.stats().msForward, b.stats().nRegions)
