#!/usr/bin/env python
"""At full cfg3 size: <G x, y> vs <x, G^H y> in float64 for the complex and the real-packed interleaved
kernels (development check; GPU only)."""
import ctypes, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from indigo_b200 import B200Backend
from indigo_b200.sense import gridding_matrix_device

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
C = 16
B = B200Backend(0)
lib, s = B._lib, B._stream
G, oN, _, _ = gridding_matrix_device(B, wl["N"], bench.make_traj(wl["traj"]), wl["oversamp"])
m, k = G.shape; nnz = int(G.values.size)
torch.manual_seed(0)
x = torch.randn(k * C * 2, device="cuda"); y = torch.randn(m * C * 2, device="cuda")
Gx = torch.empty_like(y); GHy = torch.empty_like(x)
cplx = lambda t: torch.view_as_complex(t.view(-1, 2))
def dot(a, b):
    a, b = cplx(a), cplx(b)
    tot = 0j
    step = 1 << 26
    for i in range(0, a.numel(), step):
        tot += complex(torch.vdot(a[i:i+step].to(torch.complex128), b[i:i+step].to(torch.complex128)).item())
    return tot
grid3 = (ctypes.c_int64 * 3)(*oN); tile3 = (ctypes.c_int64 * 3)(4, 4, 4); padded = ctypes.c_int64()
lib.grid_tile_rank(s, grid3, tile3, None, None, ctypes.byref(padded)); kp = padded.value
colrank = torch.empty(k, dtype=torch.int32, device="cuda"); rowmap = torch.empty(kp, dtype=torch.int32, device="cuda")
lib.grid_tile_rank(s, grid3, tile3, colrank.data_ptr(), rowmap.data_ptr(), ctypes.byref(padded))
t_ptr = torch.empty(kp + 1, dtype=torch.int32, device="cuda"); t_ind = torch.empty(nnz, dtype=torch.int32, device="cuda")
t_val = torch.empty(nnz * 2, dtype=torch.float32, device="cuda"); work = torch.empty(kp + 1, dtype=torch.int32, device="cuda")
lib.csr_transpose_conj(s, m, kp, nnz, G.values.ptr, G.colInds.ptr, G.rowPtrs.ptr, t_val.data_ptr(), t_ind.data_ptr(),
                       t_ptr.data_ptr(), work.data_ptr(), colrank.data_ptr())
hmax = (ctypes.c_float * 2)()
g_pk = torch.empty(nnz, dtype=torch.int64, device="cuda"); t_pk = torch.empty(nnz, dtype=torch.int64, device="cuda")
lib.csr_pack_real(s, nnz, G.values.ptr, G.colInds.ptr, g_pk.data_ptr(), hmax)
lib.csr_pack_real(s, nnz, t_val.data_ptr(), t_ind.data_ptr(), t_pk.data_ptr(), hmax)
out = {}
lib.ccsrmm_il(s, m, k, C, nnz, 1.0, 0.0, G.values.ptr, G.colInds.ptr, G.rowPtrs.ptr, x.data_ptr(), C, Gx.data_ptr(), C, None, 0)
lhs_c = dot(Gx, y)
Gx2 = torch.empty_like(y)
lib.ccsrmm_ilr(s, m, k, C, nnz, 1.0, 0.0, g_pk.data_ptr(), G.rowPtrs.ptr, x.data_ptr(), C, Gx2.data_ptr(), C, None, 0)
lhs_r = dot(Gx2, y)
lib.ccsrmm_il(s, kp, m, C, nnz, 1.0, 0.0, t_val.data_ptr(), t_ind.data_ptr(), t_ptr.data_ptr(), y.data_ptr(), C, GHy.data_ptr(), C, rowmap.data_ptr(), 1)
rhs_c = dot(x, GHy)
GHy2 = torch.empty_like(x)
lib.ccsrmm_ilr(s, kp, m, C, nnz, 1.0, 0.0, t_pk.data_ptr(), t_ptr.data_ptr(), y.data_ptr(), C, GHy2.data_ptr(), C, rowmap.data_ptr(), 1)
rhs_r = dot(x, GHy2)
out["lhs_complex"] = [lhs_c.real, lhs_c.imag]; out["lhs_packed"] = [lhs_r.real, lhs_r.imag]
out["rhs_complex"] = [rhs_c.real, rhs_c.imag]; out["rhs_packed"] = [rhs_r.real, rhs_r.imag]
out["rel_complex"] = abs(lhs_c - rhs_c) / abs(lhs_c); out["rel_packed"] = abs(lhs_r - rhs_r) / abs(lhs_r)
out["fwd_diff"] = float((Gx - Gx2).norm() / Gx.norm()); out["adj_diff"] = float((GHy - GHy2).norm() / GHy.norm())
d = (cplx(GHy) - cplx(GHy2)).abs().view(-1, C).amax(dim=1)
w = torch.argmax(d).item()
out["worst_row"] = w; out["worst_row_xyz"] = [w % oN[0], (w // oN[0]) % oN[1], w // (oN[0] * oN[1])]
out["worst_abs"] = float(d[w]); out["row_norm"] = float(cplx(GHy).view(-1, C)[w].abs().max())
out["n_bad_rows"] = int((d > 1e-3 * cplx(GHy).abs().max()).sum().item())
print(json.dumps(out))
