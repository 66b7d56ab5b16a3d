"""
Coil-sharded multi-GPU logic on CPU: world_size-2 `gloo` process group.

The device kernels cannot run here, so the ranks evaluate their coil slice of the SENSE normal
operator with the numpy oracle; what is under test is the host-side plumbing the B200 path uses
unchanged under NCCL: the coil partition, CoilTeam's all-reduce of the image (the only data-path
collective), the scalar hook of the reference (backend.py:469-479) and the replicated-vector rule
that keeps CG scalars out of the collectives.
"""
import os
import socket

import numpy as np
import pytest

from indigo_b200 import synth
from indigo_b200.team import coil_slice


def test_coil_slice_partitions_every_coil_once():
    for ncoils in (1, 2, 7, 8, 16, 32, 48):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s = coil_slice(ncoils, r, world)
                seen += list(range(ncoils))[s]
            assert seen == list(range(ncoils))
            sizes = [len(range(ncoils)[coil_slice(ncoils, r, world)]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from indigo_b200.team import CoilTeam
    from oracle import sense as osense
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        team = CoilTeam()
        assert (team.rank, team.world) == (rank, world) and team.replicated_vectors
        rs = np.random.RandomState(7)                        # same seed on every rank: replicated inputs
        N, C = (8, 6, 4), 4
        coord = synth.random_3d(rs, 150)
        maps = synth.unit_rss_maps(rs, N, C)
        x = synth.rand64c(rs, int(np.prod(N)), 1)
        mine = team.coils(C)
        part = osense.SenseOperator(N, coord, np.asfortranarray(maps[..., mine]), 2.0).normal(x)
        t = torch.from_numpy(np.ascontiguousarray(part).view(np.float32).copy())
        team.allreduce_tensor(t)                             # the one exchange of an apply
        total = t.numpy().view(np.complex64).reshape(part.shape)
        full = osense.SenseOperator(N, coord, maps, 2.0).normal(x)
        err = np.linalg.norm(total - full) / np.linalg.norm(full)
        # reference hook: partial scalars are summed ...
        s = team.allreduce(float(rank + 1))
        # ... but scalars of replicated vectors must not be (every rank already holds the full value)
        from indigo_b200.standalone import StandaloneBase

        class Probe(StandaloneBase):
            def dot(self, a, b): return 3.0
            def norm2(self, a): return 5.0
            pdot = __import__("indigo_b200.backend", fromlist=["B200Backend"]).B200Backend.pdot
            pnorm2 = __import__("indigo_b200.backend", fromlist=["B200Backend"]).B200Backend.pnorm2
        p = Probe()
        team.replicated_vectors = True
        rep = (p.pdot(None, None, team), p.pnorm2(None, team))
        team.replicated_vectors = False
        summed = (p.pdot(None, None, team), p.pnorm2(None, team))
        out.put((rank, float(err), s, rep, summed))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_image_allreduce_and_scalar_rules():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, s, rep, summed in res:
        assert err < 2e-6, err                               # sum over ranks of A_g^H A_g x == A^H A x
        assert s == 3.0                                      # 1 + 2
        assert rep == (3.0, 5.0)                             # replicated vectors: no collective
        assert summed == (6.0, 10.0)                         # partial vectors: summed over the 2 ranks


# ---------------------------------------------------------------------------------------------------------
# B200Backend.cg with a team whose vectors are NOT replicated (each rank holds a slice), or with a
# reference-style team that only offers allreduce(scalar): every scalar must go through pdot / pnorm2
# (backend.py:469-479, 661-677).  The solver logic is exercised here with numpy primitives under the
# standalone base; on the GPU the same method runs on the CUDA primitives.
def _np_probe_backend():
    from indigo_b200.backend import B200Backend
    from indigo_b200.standalone import StandaloneBase, DeviceArray

    class NpArray(DeviceArray):
        def _malloc(self, shape, dtype): return np.zeros(int(np.prod(shape)), dtype=dtype)
        def _free(self): pass
        def _zero(self): self._arr[:] = 0
        def _copy_from(self, arr): self._arr[:] = arr.ravel(order='F')
        def _copy_to(self, arr): arr[...] = self._arr.reshape(arr.shape, order='F')
        def _copy(self, other): self._arr[:] = other._arr

    class Probe(StandaloneBase):
        dndarray = NpArray
        def axpby(self, beta, y, alpha, x): y._arr[:] = beta * y._arr + alpha * x._arr
        def scale(self, x, alpha): x._arr[:] = alpha * x._arr
        def dot(self, x, y): return float(np.vdot(x._arr, y._arr).real)
        def norm2(self, x): return float(np.vdot(x._arr, x._arr).real)
        pdot, pnorm2, cg, _cg_host_scalars = (B200Backend.pdot, B200Backend.pnorm2, B200Backend.cg,
                                              B200Backend._cg_host_scalars)
    return Probe()


def _cg_worker(rank, world, port, out):
    import torch.distributed as dist
    from indigo_b200.team import CoilTeam
    from indigo_b200 import linop
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 40
        rs = np.random.RandomState(3)
        d = (1.0 + rs.rand(n)).astype(np.complex64)              # SPD diagonal operator, same on every rank
        b = (rs.rand(n) + 1j * rs.rand(n)).astype(np.complex64)
        lo, hi = rank * n // world, (rank + 1) * n // world       # this rank's slice of every vector
        P = _np_probe_backend()

        class Diag(linop.Operator):
            def __init__(self, backend, diag):
                linop.Operator.__init__(self, backend, name='diag')
                self.diag = diag
            shape = property(lambda self: (self.diag.size, self.diag.size))
            def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
                y._arr[:] = alpha * self.diag * x._arr + (beta * y._arr if beta != 0 else 0)

        team = CoilTeam(replicated_vectors=False)
        x = np.zeros((hi - lo, 1), dtype=np.complex64, order='F')
        its = []
        P.cg(Diag(P, d[lo:hi]), np.asfortranarray(b[lo:hi].reshape(-1, 1)), x, lamda=0.1, tol=0.0, maxiter=6,
             team=team, iterates=its)

        class ScalarOnlyTeam(object):                             # what the reference documents: allreduce(scalar)
            def allreduce(self, v): return team.allreduce(v)
        x2 = np.zeros((hi - lo, 1), dtype=np.complex64, order='F')
        P.cg(Diag(P, d[lo:hi]), np.asfortranarray(b[lo:hi].reshape(-1, 1)), x2, lamda=0.1, tol=0.0, maxiter=6,
             team=ScalarOnlyTeam())
        # serial solve of the whole system on every rank for comparison
        S = _np_probe_backend()
        xs = np.zeros((n, 1), dtype=np.complex64, order='F')
        S._cg_host_scalars(Diag(S, d), np.asfortranarray(b.reshape(-1, 1)), xs, 0.1, 0.0, 6, None)
        e1 = np.linalg.norm(x.ravel() - xs.ravel()[lo:hi]) / np.linalg.norm(xs.ravel()[lo:hi])
        e2 = np.linalg.norm(x2.ravel() - xs.ravel()[lo:hi]) / np.linalg.norm(xs.ravel()[lo:hi])
        out.put((rank, float(e1), float(e2), len(its)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_cg_with_partial_vectors_sums_its_scalars():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, nits in res:
        assert e1 < 1e-5 and e2 < 1e-5, (rank, e1, e2)
        assert nits == 6
