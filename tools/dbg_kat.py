import json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import cpecan_b200 as cp, helpers
g = json.load(open('tests/golden/reference_cases.json'))
ctx = cp.Context(0)
for c in g['cases'][:1]:
    spec = helpers.ModelSpec(c["type"], c["transitions"], c["emissions"])
    p = cp.pairwiseAlignmentBandingParameters_construct()
    for k, v in c["params"].items():
        setattr(p, k, v)
    a = np.asarray(c["anchors"], dtype=np.int64).reshape(-1, 3)
    m = spec.cpb()
    res = cp.getAlignedPairsWithIndelsUsingAnchors(m, c["sX"], c["sY"], a, p, c["raggedLeft"], c["raggedRight"], ctx=ctx)
    for got, key in zip(res, ("alignedPairs", "gapXPairs", "gapYPairs")):
        print(key, 'got', helpers.sort_triples(got).tolist(), 'want', helpers.sort_triples(c[key]).tolist())
    only = cp.getAlignedPairsUsingAnchors(m, c["sX"], c["sY"], a, p, c["raggedLeft"], c["raggedRight"], ctx=ctx)
    print('only', helpers.sort_triples(only).tolist())
