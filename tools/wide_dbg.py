import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, helpers, faulthandler
faulthandler.enable()
from cpecan_b200 import synth
import cpecan_b200 as cp
rng = np.random.default_rng(6)
p = cp.pairwiseAlignmentBandingParameters_construct()
p.splitMatrixBiggerThanThis = 1 << 40
sX = synth.random_sequence(rng, 2600, acgt_only=True)
sY = sX[:1300] + synth.random_sequence(rng, 7, acgt_only=True) + sX[1300:]
e=np.zeros((0,3),dtype=np.int64)
for name,O in (("port",helpers.port_oracle()),("ref",helpers.ref_oracle())):
    r=O.aligned_pairs(helpers.ModelSpec(0).orc(), helpers.orc_params_from(p), sX, sY, e); print(name,"before gpu",len(r), flush=True)
ctx=cp.Context(0)
b=cp.Batch(ctx,[sX],[sY],[e],[0],[0]); b.run(cp.stateMachine5_construct(),p,cp.MODE_ALIGNED_PAIRS)
off,tri=b.fetch_pairs(0); print("gpu",len(tri), b.stats().cells, b.stats().maxWidth, flush=True)
for name,O in (("port",helpers.port_oracle()),("ref",helpers.ref_oracle())):
    r=O.aligned_pairs(helpers.ModelSpec(0).orc(), helpers.orc_params_from(p), sX, sY, e); print(name,"after gpu",len(r), flush=True)
    g=helpers.sort_triples(tri); w=helpers.sort_triples(r); print(name, g.shape, w.shape, np.array_equal(g[:,1:],w[:,1:]) if g.shape==w.shape else None, flush=True)
