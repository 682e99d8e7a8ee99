"""Multi-GPU check of the slab building block (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Compares the slab matvec / Jacobi-PCG (NCCL halo exchange + all-reduces) with a single-GPU handle on the whole mesh,
then times the distributed matvec on a large mesh.  Prints one JSON line on rank 0; exit code 0 = all checks passed."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
from solidspy_opt import _capi                                   # noqa: E402
from solidspy_opt.Utils.solver import kloc_simp                   # noqa: E402
from solidspy_opt.distributed import SlabSolver, TopoBackend, slab_rows   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ke = kloc_simp(206.8e9, 0.28)
    out = {"world": world}
    # ---- correctness on a mesh one GPU can also hold ----
    nx, ny = 300, 257
    rng = np.random.default_rng(11)
    scale = 1e-9 + rng.uniform(0, 1, nx*ny)**3
    cons = np.zeros(((nx + 1)*(ny + 1), 2), dtype=np.int8)
    cons[::nx + 1] = -1
    u = rng.standard_normal((ny + 1, nx + 1, 2))
    u[:, 0, :] = 0.0
    loads = np.array([[(nx + 1)*(ny//4) + nx, 0.0, -1.0]])
    with _capi.Topo(nx, ny, ke, device=local) as full:
        full.set_cons(cons)
        full.set(_capi.F_SCALE, scale)
        full.set_loads(loads)
        y_ref = full.matvec(u.reshape(-1)).reshape(ny + 1, nx + 1, 2)
        full.solve(_capi.PRECOND_GMG, 1e-12, 5000, warm=False)
        u_ref = full.get(_capi.F_U).reshape(ny + 1, nx + 1, 2)
        f_full = full.get(_capi.F_LOAD).reshape(ny + 1, nx + 1, 2)
    e0, e1 = slab_rows(ny, world, rank)
    be = TopoBackend(nx, e1 - e0, ke, cons.reshape(ny + 1, nx + 1, 2)[e0:e1 + 1].reshape(-1, 2),
                     scale.reshape(ny, nx)[e0:e1].reshape(-1), local)
    S = SlabSolver(be, rank, world, dist)
    ul, yl = be.from_numpy(u[e0:e1 + 1]), be.empty()
    S.matvec(ul, yl)
    torch.cuda.synchronize()
    err_mv = float(np.abs(yl.cpu().numpy() - y_ref[e0:e1 + 1]).max()/np.abs(y_ref).max())
    us, it, rel = S.solve_jacobi_pcg(be.from_numpy(f_full[e0:e1 + 1]), rtol=1e-10, maxit=200000)
    err_u = float(np.abs(us.cpu().numpy() - u_ref[e0:e1 + 1]).max()/np.abs(u_ref).max())
    errs = torch.tensor([err_mv, err_u], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    out.update(matvec_rel_err=float(errs[0]), pcg_rel_err=float(errs[1]), pcg_iters=it, pcg_relres=rel)
    be.topo.close()
    # ---- distributed matvec timing on a mesh split over the ranks (strong scaling of stage 1) ----
    nx, ny = (8192, 4096) if os.environ.get("DIST_BIG", "1") == "1" else (2048, 1024)
    e0, e1 = slab_rows(ny, world, rank)
    nyl = e1 - e0
    cons_l = np.zeros(((nx + 1)*(nyl + 1), 2), dtype=np.int8)
    cons_l[::nx + 1] = -1
    be = TopoBackend(nx, nyl, ke, cons_l, np.full(nx*nyl, 0.125), local)
    S = SlabSolver(be, rank, world, dist)
    a, b = be.zeros(), be.empty()
    a.normal_()
    a[:, 0, :] = 0
    for _ in range(5):
        S.matvec(a, b)
    torch.cuda.synchronize()
    dist.barrier()
    e_0, e_1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e_0.record()
    for _ in range(reps):
        S.matvec(a, b)
    e_1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e_0.elapsed_time(e_1)/reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nn, ne = (nx + 1)*(ny + 1), nx*ny
    out.update(mesh=[nx, ny], matvec_ms=float(ms), matvec_algorithmic_gbs=(32*nn + 8*ne)/(float(ms)*1e-3)/1e9,
               halo_bytes_per_rank=S.halo_bytes_per_exchange)
    be.topo.close()
    ok = out["matvec_rel_err"] < 1e-13 and out["pcg_rel_err"] < 1e-6
    if rank == 0:
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
