"""Test-side helpers: ctypes loaders for the oracle libraries (oracle/oracle_api.h) and case generators.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "libcpecan_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libcpecan_ref.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("threshold", C.c_double),
        ("minDiagsBetweenTraceBack", C.c_int64),
        ("traceBackDiagonals", C.c_int64),
        ("diagonalExpansion", C.c_int64),
        ("constraintDiagonalTrim", C.c_int64),
        ("splitMatrixBiggerThanThis", C.c_int64),
        ("dynamicAnchorExpansion", C.c_int64),
    ]


class OrcModel(C.Structure):
    _fields_ = [("type", C.c_int64), ("fromHmm", C.c_int64), ("transitions", C.c_double * 25), ("emissions", C.c_double * 80)]


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One of the two implementations of oracle_api.h."""

    def __init__(self, path):
        self.path = path
        L = C.CDLL(path)
        L.orc_identity.restype = C.c_char_p
        L.orc_logadd.restype = C.c_double
        L.orc_logadd.argtypes = [C.c_double, C.c_double]
        for f in ("orc_band", "orc_split_points", "orc_model_dump", "orc_aligned_pairs"):
            getattr(L, f).restype = C.c_int64
        L.orc_forward_prob.restype = C.c_double
        self.L = L
        self.identity = L.orc_identity().decode()

    def default_params(self):
        p = OrcParams()
        self.L.orc_default_params(C.byref(p))
        return p

    def logadd(self, x, y):
        return self.L.orc_logadd(x, y)

    def band(self, anchors, lX, lY, expansion, dynamic=False):
        a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1)
        out = np.zeros(3 * (lX + lY + 1), dtype=np.int64)
        self.L.orc_band(_i64p(a) if a.size else None, C.c_int64(a.size // 3), C.c_int64(lX), C.c_int64(lY), C.c_int64(expansion),
                        int(dynamic), _i64p(out))
        return out.reshape(-1, 3)

    def split_points(self, anchors, lX, lY, split, rl, rr):
        a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1)
        cap = a.size // 3 + 2
        out = np.zeros(4 * cap, dtype=np.int64)
        n = self.L.orc_split_points(_i64p(a) if a.size else None, C.c_int64(a.size // 3), C.c_int64(lX), C.c_int64(lY), C.c_int64(split),
                                    int(rl), int(rr), _i64p(out), C.c_int64(cap))
        return [tuple(int(v) for v in out[4 * i:4 * i + 4]) for i in range(n)]

    def model_dump(self, m):
        out = np.zeros(8192, dtype=np.float64)
        n = self.L.orc_model_dump(C.byref(m), _f64p(out), C.c_int64(out.size))
        assert n <= out.size
        return out[:n].copy()

    def aligned_pairs(self, m, p, sX, sY, anchors, rl=False, rr=False):
        a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1)
        cap = 1 << 16
        while True:
            out = np.zeros(3 * cap, dtype=np.int64)
            n = self.L.orc_aligned_pairs(C.byref(m), C.byref(p), _b(sX), _b(sY), _i64p(a) if a.size else None, C.c_int64(a.size // 3),
                                         int(rl), int(rr), _i64p(out), C.c_int64(cap))
            if n <= cap:
                return out[:3 * n].reshape(-1, 3).copy()
            cap = int(n)

    def aligned_pairs_with_indels(self, m, p, sX, sY, anchors, rl=False, rr=False):
        a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1)
        cap = 1 << 16
        while True:
            outs = [np.zeros(3 * cap, dtype=np.int64) for _ in range(3)]
            counts = np.zeros(3, dtype=np.int64)
            self.L.orc_aligned_pairs_with_indels(C.byref(m), C.byref(p), _b(sX), _b(sY), _i64p(a) if a.size else None,
                                                 C.c_int64(a.size // 3), int(rl), int(rr), _i64p(outs[0]), _i64p(outs[1]), _i64p(outs[2]),
                                                 C.c_int64(cap), _i64p(counts))
            if counts.max() <= cap:
                return [o[:3 * int(c)].reshape(-1, 3).copy() for o, c in zip(outs, counts)]
            cap = int(counts.max())

    def expectations(self, m, p, sX, sY, anchors, rl=False, rr=False):
        S = 5 if m.type < 2 else 3
        a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1)
        hmm = np.zeros(S * S + S * 16 + 1, dtype=np.float64)
        self.L.orc_expectations(C.byref(m), C.byref(p), _b(sX), _b(sY), _i64p(a) if a.size else None, C.c_int64(a.size // 3), int(rl),
                                int(rr), _f64p(hmm))
        return hmm

    def forward_prob(self, m, p, sX, sY, anchors, rl=False, rr=False):
        a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1)
        return self.L.orc_forward_prob(C.byref(m), C.byref(p), _b(sX), _b(sY), _i64p(a) if a.size else None, C.c_int64(a.size // 3),
                                       int(rl), int(rr))

    def batch(self, m, p, packed, mode=0, threads=1, rl=None, rr=None):
        n = len(packed["xOff"]) - 1
        S = 5 if m.type < 2 else 3
        counts = np.zeros(n, dtype=np.int64)
        checks = np.zeros(n, dtype=np.int64)
        hmm = np.zeros(S * S + S * 16 + 1, dtype=np.float64)
        u8 = C.POINTER(C.c_uint8)
        self.L.orc_batch(C.byref(m), C.byref(p), C.c_int64(n), C.c_void_p(packed["seqX"].ctypes.data), _i64p(packed["xOff"]),
                         C.c_void_p(packed["seqY"].ctypes.data), _i64p(packed["yOff"]), _i64p(packed["anchors"]), _i64p(packed["aOff"]),
                         rl.ctypes.data_as(u8) if rl is not None else None, rr.ctypes.data_as(u8) if rr is not None else None,
                         int(mode), int(threads), _i64p(counts), _i64p(checks), _f64p(hmm))
        return counts, checks, hmm


def _b(s):
    return s.encode() if isinstance(s, str) else bytes(s)


def build_port():
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "pairhmm_oracle.c")):
        subprocess.check_call(["make", "-s", "-f", os.path.join(ORACLE_DIR, "Makefile"), "port"])
    return PORT_SO


_cache = {}


def port_oracle():
    if "port" not in _cache:
        _cache["port"] = Oracle(build_port())
    return _cache["port"]


def ref_oracle():
    """The reference's own code (oracle/_ref), or None when it has not been built / did not travel."""
    if "ref" not in _cache:
        if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/impl"):
            subprocess.call(["make", "-s", "-f", os.path.join(ORACLE_DIR, "Makefile"), "ref"])
        _cache["ref"] = Oracle(REF_SO) if os.path.exists(REF_SO) else None
    return _cache["ref"]


def best_oracle():
    return ref_oracle() or port_oracle()


# ---- model / parameter specs usable on both sides ----
class ModelSpec:
    def __init__(self, type_, transitions=None, emissions=None):
        self.type = type_
        self.S = 5 if type_ < 2 else 3
        self.transitions = None if transitions is None else np.ascontiguousarray(transitions, dtype=np.float64).reshape(self.S, self.S)
        self.emissions = None if emissions is None else np.ascontiguousarray(emissions, dtype=np.float64).reshape(self.S, 16)

    @staticmethod
    def random(rng, type_):
        S = 5 if type_ < 2 else 3
        t = rng.random((S, S)) + 0.01
        t /= t.sum(1, keepdims=True)
        e = rng.random((S, 16)) + 0.01
        e /= e.sum(1, keepdims=True)
        return ModelSpec(type_, t, e)

    def orc(self):
        m = OrcModel()
        m.type = self.type
        m.fromHmm = 0 if self.transitions is None else 1
        if self.transitions is not None:
            for i, v in enumerate(self.transitions.ravel()):
                m.transitions[i] = v
            for i, v in enumerate(self.emissions.ravel()):
                m.emissions[i] = v
        return m

    def cpb(self):
        import cpecan_b200 as cp

        if self.transitions is None:
            return cp.stateMachine5_construct(self.type) if self.S == 5 else cp.stateMachine3_construct(self.type)
        return cp.hmm_getStateMachine(self.type, self.transitions, self.emissions)


def orc_params_from(p):
    """CpbParams -> OrcParams"""
    o = OrcParams()
    o.threshold = p.threshold
    o.minDiagsBetweenTraceBack = p.minDiagsBetweenTraceBack
    o.traceBackDiagonals = p.traceBackDiagonals
    o.diagonalExpansion = p.diagonalExpansion
    o.constraintDiagonalTrim = p.constraintDiagonalTrim
    o.splitMatrixBiggerThanThis = p.splitMatrixBiggerThanThis
    o.dynamicAnchorExpansion = p.dynamicAnchorExpansion
    return o


def sort_triples(t):
    """canonical order for set comparison: by (x, y)"""
    t = np.asarray(t, dtype=np.int64).reshape(-1, 3)
    if t.shape[0] == 0:
        return t
    order = np.lexsort((t[:, 2], t[:, 1]))
    return t[order]
