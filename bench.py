#!/usr/bin/env python
"""
bench.py -- BASELINE.json's headline metric on its headline configuration.

    metric   SENSE-NUFFT A^H A applies/sec (whole job), plus ccsrmm / fftn as % of HBM roofline
    workload cfg3 = BASELINE.json configs[2]: 3-D radial (kooshball) SENSE-NUFFT, image 208^3,
             2x oversampled grid 416^3, 16 coils, 16384 spokes x 416 samples (M = 6 815 744),
             the -O3 tree of examples/pics.py: one apply = six Backend calls
             ccsrmm(P^H,adj) -> fftn -> ccsrmm(G') -> ccsrmm(G',adj) -> ifftn -> ccsrmm(P^H)
    step     one A^H A apply (AHA.eval(y, x)) on synthetic, seeded data; with N > 1 GPUs the 16
             coils are sharded over the ranks (strong scaling) and each apply ends with one NCCL
             all-reduce of the 208^3 image.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg1|tiny]
    torchrun ... bench.py --gpus N ...            (one rank per GPU; rank 0 prints ONE JSON line)

`--impl reference` times the reference's CPU implementation of the same apply on this box's
host cores (numpy/scipy calls of indigo/backends/np.py through the oracle port, SpMM rows
through the reference's own OpenMP kernel oracle/_ref/_customcpu when it is present), each step
a bounded sample of the workload extrapolated to a full apply (see `cpu_baseline.sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

C64 = np.dtype("complex64")

WORKLOADS = {
    # name: (image N, coils, trajectory, oversamp)
    "cfg3": dict(N=(208, 208, 208), C=16, traj=("kooshball", 16384, 416), oversamp=2.0,
                 desc="cfg3: 3-D radial SENSE-NUFFT A^H A apply, image 208^3, grid 416^3 (2x oversampled), 16 coils, "
                      "16384 spokes x 416 samples (M=6815744), -O3 tree, 6 backend calls per apply"),
    "cfg1": dict(N=(256, 256, 1), C=8, traj=("radial2d", 402, 512), oversamp=2.0,
                 desc="cfg1: 2-D radial SENSE-NUFFT A^H A apply, image 256x256, grid 512x512x2, 8 coils, 402 spokes x 512"),
    "tiny": dict(N=(32, 32, 32), C=4, traj=("kooshball", 256, 64), oversamp=2.0,
                 desc="tiny: development smoke size (not a reportable workload)"),
}


def make_traj(spec):
    from indigo_b200 import synth
    kind, a, b = spec
    return synth.kooshball_3d(a, b) if kind == "kooshball" else synth.radial_2d(a, b)


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, world, call):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant step, from the committed
    `ncu --set full` capture of the same build and workload (profiles/traffic.json); None when no capture
    matches (other workloads, other GPU counts)."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    table = json.load(open(path))
    entry = table.get("%s:%d" % (workload, world), {})
    for prefix, rec in entry.items():
        if call.startswith(prefix):
            return rec["bytes"], rec["source"]
    return None, None


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- ours
def algorithmic_bytes(call):
    """each operand counted once.  The fused steps are charged the
    compulsory traffic of the fused formulation (pruned passes), not of the calls they replace."""
    if call["op"] == "ccsrmm":
        b = call["nnz"] * 12 + (call["m"] + 1) * 4 + 8 * call["ncols"] * (call["k"] + call["m"])
        return b + (8 * call["ncols"] * (call["k"] if call["adjoint"] else call["m"]) if call["beta_nz"] else 0)
    if call["op"] == "fused_fft":
        N, oN, C = call["N"], call["oN"], call["C"]
        nvox = N[0] * N[1] * N[2]
        px = oN[0] * N[1] * N[2]                 # points after the x pass
        py = oN[0] * oN[1] * N[2]                # after the y pass
        pz = oN[0] * oN[1] * oN[2]
        # image + pf, then (read + write) of each of the three pruned passes
        return 8 * nvox + 8 * nvox * C + 8 * C * (px + (px + py) + (py + pz))
    return 16 * call["points"]


class CallTimer(object):
    """Brackets every Backend call (or fused step) of the apply with CUDA events on the launching stream."""

    def __init__(self, B, torch, dev=None):
        self.B, self.torch, self.records, self.on = B, torch, [], False
        for name in ("ccsrmm", "ccsrmm_packed", "fftn", "ifftn"):
            setattr(B, name, self._wrap(name, getattr(B, name)))
        if dev is not None:
            fft = dict(op="fused_fft", N=dev.N, oN=dev.oN, C=dev.C)
            gfw = dict(op="ccsrmm", m=dev.M, k=dev.on, nnz=dev.nnz, ncols=dev.C, adjoint=False, beta_nz=False)
            gad = dict(op="ccsrmm", m=dev.on, k=dev.M, nnz=dev.nnz, ncols=dev.C, adjoint=False, beta_nz=False)
            shp = "x".join(str(v) for v in dev.oN)
            for name, key, info in (
                    ("expand_fft", "expand_fft[pf.*x -> zpad -> FFT3 %s x%d coils, pruned]" % (shp, dev.C), fft),
                    ("grid_to_samples", "ccsrmm_il[G' %dx%d nnz/row=%.0f ncols=%d]" % (dev.M, dev.on, dev.nnz / dev.M, dev.C), gfw),
                    ("samples_to_grid", "ccsrmm_il[G'^H stored %dx%d nnz/row=%.0f ncols=%d]" % (dev.on, dev.M, dev.nnz / dev.on, dev.C), gad),
                    ("ifft_combine", "ifft_combine[IFFT3 %s x%d coils -> crop -> sum_c conj(pf), pruned]" % (shp, dev.C), fft)):
                setattr(dev, name, self._wrap_fused(key, info, getattr(dev, name)))

    def _wrap_fused(self, key, info, fn):
        def timed(*a, **k):
            if not self.on:
                return fn(*a, **k)
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            n0 = self.B._lib.launch_count()
            e0.record(); out = fn(*a, **k); e1.record()
            self.records.append((key, info, e0, e1, self.B._lib.launch_count() - n0))
            return out
        return timed

    def _wrap(self, name, fn):
        def timed(*a, **k):
            if not self.on:
                return fn(*a, **k)
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            n0 = self.B._lib.launch_count()
            e0.record(); out = fn(*a, **k); e1.record()
            if name == "ccsrmm_packed":
                y, shp, nnz, _, _, x = a[:6]
                beta = k.get("beta", a[7] if len(a) > 7 else 0)
                info = dict(op="ccsrmm", m=int(shp[0]), k=int(shp[1]), nnz=int(nnz), ncols=int(x.shape[1]),
                            adjoint=False, beta_nz=bool(beta != 0))
                key = "ccsrmm[packed real A %dx%d nnz/row=%.0f ncols=%d]" % (shp[0], shp[1], nnz / max(1, shp[0]), x.shape[1])
            elif name == "ccsrmm":
                y, shp, ind, ptr, vals, x = a[:6]
                adj = k.get("adjoint", a[8] if len(a) > 8 else False)
                beta = k.get("beta", a[7] if len(a) > 7 else 0)
                info = dict(op="ccsrmm", m=int(shp[0]), k=int(shp[1]), nnz=int(vals.size), ncols=int(x.shape[1]),
                            adjoint=bool(adj), beta_nz=bool(beta != 0))
                key = "ccsrmm[%s %dx%d nnz/row=%.0f ncols=%d]" % ("A^H" if adj else "A", shp[0], shp[1],
                                                                   vals.size / max(1, shp[0]), x.shape[1])
            else:
                x = a[1]
                info = dict(op=name, points=int(np.prod(x.shape)))
                key = "%s%s" % (name, tuple(int(s) for s in x.shape))
            self.records.append((key, info, e0, e1, self.B._lib.launch_count() - n0))
            return out
        return timed

    def summary(self, peak):
        agg = {}
        for key, info, e0, e1, nl in self.records:
            d = agg.setdefault(key, dict(info=info, ms=[], launches=nl))
            d["ms"].append(e0.elapsed_time(e1))
        out = []
        for key, d in agg.items():
            ms = float(np.mean(d["ms"]))
            nbytes = algorithmic_bytes(d["info"])
            out.append(dict(call=key, ms=ms, launches_per_call=d["launches"], algorithmic_bytes=nbytes,
                            gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / peak))
        tot = sum(o["ms"] for o in out)
        for o in out:
            o["share"] = o["ms"] / tot if tot else 0.0
        return sorted(out, key=lambda o: -o["ms"])


def run_ours(args):
    import torch
    import torch.distributed as dist
    from indigo_b200 import B200Backend, synth
    from indigo_b200.sense import sense_operator_device, normal_operator
    from indigo_b200.fused import sense_operator_fused
    from indigo_b200.team import CoilTeam, coil_slice

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    N, C = wl["N"], wl["C"]
    if getattr(args, "coils", 0):
        C = args.coils
        wl = dict(wl, desc=wl["desc"] + " [DEVELOPMENT RUN with %d coils, not the named configuration]" % C)
    if C % world:
        raise SystemExit("coil count %d not divisible by %d ranks" % (C, world))
    B = B200Backend(local)
    team = CoilTeam() if world > 1 else None
    rs = np.random.RandomState(2024)
    coord = make_traj(wl["traj"])
    maps = synth.unit_rss_maps(rs, N, C)
    mine = coil_slice(C, rank, world)
    t0 = time.time()
    my_maps = np.asfortranarray(maps[..., mine])
    tree = args.tree
    A = None
    if tree == "fused":
        try:
            A = sense_operator_fused(B, N, coord, my_maps, wl["oversamp"])
        except RuntimeError as e:                       # grid without specialised passes (e.g. cfg1's 512x512x2)
            print("fused path unavailable (%s); using the six-call tree" % e, file=sys.stderr)
            tree = "o3"
    if A is None:
        A = sense_operator_device(B, N, coord, my_maps, wl["oversamp"])
    AHA = normal_operator(A)
    B.barrier()
    setup_s = time.time() - t0
    nvox = int(np.prod(N))
    x_h = B.pinned_array((nvox, 1)); x_h[...] = synth.rand64c(rs, nvox, 1)
    y_h = B.pinned_array((nvox, 1))
    x_d = B.copy_array(np.asarray(x_h)); y_d = B.zero_array((nvox, 1), C64)
    timer = CallTimer(B, torch, getattr(A, '_dev', None))
    lib = B._lib

    def apply():
        AHA.eval(y_d, x_d)
        if team is not None:
            team.allreduce_array(y_d)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        apply()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- timed region: inputs resident in HBM -----------------------------------------
    timer.on = True
    lib.launch_count_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for a, b in ev:
        a.record(); apply(); b.record()
    sync_all()
    timer.on = False
    launches = lib.launch_count()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---- end to end: host buffers, H2D + apply + D2H every step ------------------------
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for a, b in e2e_ev:
        a.record()
        x_d.copy_from(x_h)                      # pinned -> device, async on the stream
        apply()
        y_d.copy_to(y_h)                        # device -> pinned, synchronises
        b.record()
    sync_all()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_ev)
    e2e_path = "pinned host -> cudaMemcpyAsync -> AHA.eval -> cudaMemcpyAsync -> pinned host"
    # ---- end to end, zero-copy: the operator is evaluated on device-mapped views of the same pinned host
    # buffers (B.mapped_array): the image crosses PCIe inside the first fused pass' load and the result inside
    # the last pass' store, every step, instead of through separate copies
    e2e_alt = None
    try:
        x_m = B.mapped_array(x_h)
        y_m = B.mapped_array(y_h) if team is None else None

        def apply_mapped():
            if team is None:
                AHA.eval(y_m, x_m)
            else:
                AHA.eval(y_d, x_m)
                team.allreduce_array(y_d)
                y_d.copy_to(y_h)

        y_ref = np.array(y_h)                                   # result of the copy path, same input
        apply_mapped(); sync_all()
        err = float(np.linalg.norm(np.asarray(y_h) - y_ref) / max(np.linalg.norm(y_ref), 1e-30))
        if err > 1e-6:
            raise RuntimeError("mapped path deviates from the copy path: %.3e" % err)
        m_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        sync_all()
        for a, b in m_ev:
            a.record(); apply_mapped(); b.record()
            if team is None:
                b.synchronize()                                 # the result is in host memory before the next step starts
        sync_all()
        m_ms = sum(a.elapsed_time(b) for a, b in m_ev)
        e2e_alt = {"copy_path_ms_per_step": e2e_ms / args.steps, "mapped_path_ms_per_step": m_ms / args.steps}
        if m_ms < e2e_ms:
            e2e_ms = m_ms
            e2e_path = ("AHA.eval on device-mapped views of the pinned host buffers (B.mapped_array): H2D inside the first "
                        "pass' load, D2H inside the last pass' store" + ("" if team is None else "; result via all-reduce + copy"))
    except Exception as exc:                                    # keep the copy path's number
        e2e_alt = {"mapped_path_error": str(exc)[:200]}
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    peak, peak_src = peaks()
    calls = timer.summary(peak)
    if rank == 0:
        dom = calls[0]
        traffic, traffic_src = measured_traffic(args.workload, world, dom["call"])
        line = {
            "metric": "SENSE-NUFFT A^H A applies/sec", "value": args.steps / (total_ms * 1e-3), "unit": "applies/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex64 (fp32 accumulate)",
            "data": "synthetic (seeded kooshball trajectory, rand64c image and unit-RSS coil maps)",
            "config": {"workload": wl["desc"],
                       "tree": ("fused B200 recipe: expand+FFT (pruned, coil-interleaved, support windows) -> separable G' gather -> "
                                "x-run G'^H gather -> IFFT+combine; 4 fused steps replace the six calls of the -O3 tree" if tree == "fused" else
                                "-O3 (examples/pics.py recipe), device-built CSR operands, six Backend calls"),
                       "parallelism": "coil-sharded x%d, NCCL all-reduce of the image" % world if world > 1 else "single GPU",
                       "l2": "no explicit flush: every call streams operands far larger than L2 (grid %.1f GB)" %
                             (8.0 * np.prod([int(n * wl["oversamp"]) for n in N]) * C / world / 1e9)},
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "applies/s",
                    "h2d_bytes_per_step": int(x_h.nbytes), "d2h_bytes_per_step": int(y_h.nbytes),
                    "path": e2e_path, "paths_timed": e2e_alt},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom["call"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "share_of_step": dom["share"], "launches_per_call": dom["launches_per_call"]},
            "calls": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in c.items()} for c in calls],
            "clocks": clk, "setup_s": round(setup_s, 1),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, steps=1)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_baseline(workload, steps=1, warmup=0):
    """The reference's CPU path on a bounded sample of the workload, extrapolated to one apply.

    Sample: ONE coil on the full oversampled grid (FFT + IFFT of one coil volume, P expand/combine
    of one coil) and the first 1/64 of the spokes for the gridding SpMM pair with that coil.
    One apply = C coils x (2 FFTs + P pair) + C coils x 64 x (G pair on the sample): every term is
    linear in the coil count and in the number of samples."""
    from indigo_b200 import synth
    from oracle import np_oracle as K
    from oracle import sense as osense
    import scipy.sparse as spp

    wl = WORKLOADS[workload]
    N, C = wl["N"], wl["C"]
    kind, nsp, nread = wl["traj"]
    frac = 64 if workload == "cfg3" else 1
    coord = make_traj(wl["traj"])[:, :, ::frac]            # every frac-th spoke: same angular coverage
    rs = np.random.RandomState(2024)
    maps1 = synth.unit_rss_maps(rs, N, 1)
    op = osense.SenseOperator(N, coord, maps1, wl["oversamp"])
    ref = K.load_ref_customcpu()
    threads = os.cpu_count() or 1
    nvox, on = int(np.prod(N)), int(np.prod(op.oN))
    x = synth.rand64c(rs, nvox, 1)

    def spmm(y, A, xin, adjoint):
        if ref is not None and xin.shape[1] > 1:
            ref.csrmm(adjoint, A.shape[0], xin.shape[1], A.shape[1], 1.0 + 0j, A.data, A.indices, A.indptr,
                      xin, xin.shape[0], 0j, y, y.shape[0], False)
        else:
            K.ccsrmm(y, A.shape, A.indices, A.indptr, A.data, xin, 1, 0, adjoint=adjoint)

    def one():
        t = {}
        g = np.zeros((on, 1), dtype=C64, order="F")
        t0 = time.perf_counter(); spmm(g, op.PH, x, True); t["P"] = time.perf_counter() - t0
        G3 = g.reshape(op.oN + (1,), order="F"); F3 = np.zeros_like(G3, order="F")
        t0 = time.perf_counter(); K.fftn(F3, G3); t["fft"] = time.perf_counter() - t0
        f = F3.reshape((on, 1), order="F")
        k = np.zeros((op.M, 1), dtype=C64, order="F")
        t0 = time.perf_counter(); spmm(k, op.G, f, False); t["G"] = time.perf_counter() - t0
        t0 = time.perf_counter(); spmm(f, op.G, k, True); t["GH"] = time.perf_counter() - t0
        t0 = time.perf_counter(); K.ifftn(G3, F3); t["ifft"] = time.perf_counter() - t0
        yv = np.zeros((nvox, 1), dtype=C64, order="F")
        t0 = time.perf_counter(); spmm(yv, op.PH, G3.reshape((on, 1), order="F"), False); t["PH"] = time.perf_counter() - t0
        return t

    for _ in range(warmup):
        one()
    ts = [one() for _ in range(max(1, steps))]
    t = {k: float(np.median([d[k] for d in ts])) for k in ts[0]}
    full = C * (t["P"] + t["fft"] + t["ifft"] + t["PH"]) + C * frac * (t["G"] + t["GH"])
    return {"value": 1.0 / full, "unit": "applies/s", "cores": 1, "kind": "port",
            "threads_available": threads,
            "sample": "1 of %d coils on the full %s grid (fft %.2fs, ifft %.2fs, P pair %.2fs) + gridding pair on 1/%d of "
                      "the spokes (G %.3fs, G^H %.3fs); numpy pocketfft + scipy csr_matvecs as indigo/backends/np.py calls "
                      "them (single-threaded, like the reference NumpyBackend); extrapolated linearly to %d coils x all "
                      "samples = %.1f s per apply" % (C, "x".join(str(v) for v in op.oN), t["fft"], t["ifft"],
                                                      t["P"] + t["PH"], frac, t["G"], t["GH"], C, full),
            "seconds_per_apply": full, "parts": t}


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    t0 = time.time()
    base = cpu_baseline(args.workload, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "SENSE-NUFFT A^H A applies/sec", "value": base["value"], "unit": "applies/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * base["seconds_per_apply"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex64",
            "data": "synthetic (same seeded generators as the B200 arm)",
            "config": {"workload": wl["desc"], "tree": "-O3", "parallelism": "host CPU"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--coils", type=int, default=0,
                    help="development only: override the workload's coil count (e.g. 2 = the per-GPU shard of cfg3 at 8 GPUs)")
    ap.add_argument("--tree", default="fused", choices=["fused", "o3"],
                    help="fused: backend-specific fused recipe (default); o3: the reference's six-call -O3 tree")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
