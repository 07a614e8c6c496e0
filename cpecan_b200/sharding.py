"""Multi-GPU host logic: one process per GPU, pairs partitioned over ranks, no data-path collective.

Independent pairs share nothing but the read-only model, so configs 2, 3 and 5 shard with
no traffic at all.  The only collective of the whole path is the EM reduction: every rank's expected
transition / emission counts (CPB_HMM_LEN(S) doubles: 106 for the five-state, 58 for the three-state machine) are
summed with one all-reduce per EM iteration -- the NCCL replacement of cPecanEm.py:182-188, which sums the
per-job expectation files.
"""
import numpy as np


def contiguous_shards(costs, world_size):
    """Split items 0..n-1 into world_size contiguous ranges of roughly equal total cost (band cells).
    Contiguous, so results concatenate back in input order.  Returns world_size+1 boundaries."""
    costs = np.asarray(costs, dtype=np.float64)
    n = costs.size
    if n == 0:
        return [0] * (world_size + 1)
    prefix = np.concatenate([[0.0], np.cumsum(costs)])
    total = prefix[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(prefix, target, side="left"))
        # choose the boundary closest to the target, never moving backwards
        if k > 0 and abs(prefix[k - 1] - target) <= abs(prefix[min(k, n)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return bounds


def shard_packed(packed, lo, hi):
    """The pairs [lo, hi) of a packed batch (see synth.evolved_pairs) as a packed batch of their own."""
    xo, yo, ao = packed["xOff"], packed["yOff"], packed["aOff"]
    anchors = packed["anchors"][3 * ao[lo]:3 * ao[hi]]
    if anchors.size == 0:
        anchors = np.zeros(3, dtype=np.int64)
    out = dict(
        seqX=np.ascontiguousarray(np.concatenate([packed["seqX"][xo[lo]:xo[hi]], np.zeros(1, dtype=np.uint8)])),
        xOff=np.ascontiguousarray(xo[lo:hi + 1] - xo[lo]),
        seqY=np.ascontiguousarray(np.concatenate([packed["seqY"][yo[lo]:yo[hi]], np.zeros(1, dtype=np.uint8)])),
        yOff=np.ascontiguousarray(yo[lo:hi + 1] - yo[lo]),
        anchors=np.ascontiguousarray(anchors),
        aOff=np.ascontiguousarray(ao[lo:hi + 1] - ao[lo]),
    )
    for k in ("rl", "rr"):
        if packed.get(k) is not None:
            out[k] = np.ascontiguousarray(packed[k][lo:hi])
    return out


def estimate_cost(packed, expansion):
    """Cheap per-pair cost proxy for partitioning before the band is built: diagonals x (anchor-gap width + expansion)."""
    xo, yo, ao = packed["xOff"], packed["yOff"], packed["aOff"]
    n = xo.size - 1
    lx, ly = np.diff(xo), np.diff(yo)
    na = np.diff(ao)
    gap = (lx + ly) / (2.0 * (na + 1))
    return (lx + ly) * (np.minimum(gap, np.minimum(lx, ly) + 1) + expansion + 1) if n else np.zeros(0)


def allreduce_expectations(local, group=None, device_tensor=None):
    """Sum the per-rank Hmm expectation vectors over all ranks (torch.distributed: NCCL on GPUs, gloo on CPU).

    local: float64 numpy vector (CPB_HMM_LEN(S)); device_tensor: optionally a CUDA tensor that already holds the
    rank's totals in HBM (then nothing is staged through the host).  Returns the summed vector as numpy."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local if device_tensor is None else device_tensor.cpu().numpy(), dtype=np.float64).copy()
    if device_tensor is not None:
        t = device_tensor
    else:
        t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).clone()
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class _DeviceDoubles:
    """A raw device pointer to n doubles as something torch.as_tensor understands (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_vector(ptr, n):
    """torch float64 CUDA tensor aliasing n doubles at the device address ptr (e.g. Batch.device_expectation_total_ptr())."""
    import torch

    return torch.as_tensor(_DeviceDoubles(ptr, n), device="cuda")


def normalise_hmm(vec, S):
    """hmm_normalise (impl/stateMachine.c:88-112) on the flat vector: rows of the transition matrix and each state's
    emission matrix sum to one; the likelihood entry is left alone."""
    v = np.array(vec, dtype=np.float64)
    t = v[:S * S].reshape(S, S)
    t /= t.sum(axis=1, keepdims=True)
    e = v[S * S:S * S + S * 16].reshape(S, 16)
    e /= e.sum(axis=1, keepdims=True)
    return v
