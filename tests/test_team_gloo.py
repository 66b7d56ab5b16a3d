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
        from indigo_b200.host.hostbackend import HostBackend

        class Probe(HostBackend):
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
