"""
Coil-sharded multi-GPU execution: one process per GPU,
each rank builds the same SENSE tree with its slice of the coil maps, so
A^H A = sum_g A_g^H A_g needs exactly one exchange per apply -- an all-reduce
(sum) of the N-element image -- carried by NCCL over NVLink/NVSwitch through
torch.distributed.  Image-domain vectors (x, r, p, Ap) are replicated, so the
CG scalars every rank computes are bit-identical (deterministic reductions,
blas1.cu) and need no collective; `replicated_vectors` tells
Backend.pdot/pnorm2 not to sum them (the reference's hook,
backend.py:469-479, would otherwise multiply them by the world size).

The reference ships no communicator; its only hook is the duck-typed
`team.allreduce(scalar)` -- kept here for partial (non-replicated) scalars.
"""
import numpy as np


def coil_slice(ncoils, rank, world):
    """Contiguous block of coils owned by `rank` (first `ncoils % world` ranks get one extra)."""
    base, extra = divmod(int(ncoils), int(world))
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


class CoilTeam(object):
    def __init__(self, group=None, replicated_vectors=True):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("CoilTeam needs an initialised torch.distributed process group")
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.replicated_vectors = replicated_vectors

    def coils(self, ncoils):
        return coil_slice(ncoils, self.rank, self.world)

    # reference hook: sum a host scalar over ranks (backend.py:472,478)
    def allreduce(self, value):
        import torch
        dev = 'cuda' if self._dist.get_backend(self.group) == 'nccl' else 'cpu'
        t = torch.tensor([float(np.real(value))], dtype=torch.float64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def allreduce_tensor(self, t):
        """In-place sum of a torch tensor over ranks on the current stream."""
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_array(self, d_arr):
        """In-place sum of a contiguous complex64 device array (as 2N floats)."""
        self.allreduce_tensor(as_torch(d_arr))
        return d_arr

    def barrier(self):
        self._dist.barrier(group=self.group)


def as_torch(d_arr):
    """float32 torch view (no copy) of a contiguous B200 device array."""
    import torch
    assert d_arr.contiguous or d_arr.shape[1] == 1, "all-reduce needs a contiguous array"
    base = d_arr._arr._keep
    if base is None:
        raise RuntimeError("array has no torch owner")
    off = d_arr._arr.value - base.data_ptr()
    nbytes = int(d_arr.size) * np.dtype(d_arr.dtype).itemsize
    return base[off:off + nbytes].view(torch.float32)
