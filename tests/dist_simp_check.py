"""Multi-GPU check of the row-slab SIMP iteration (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dist_simp_check.py
Every rank owns one slab (SlabSimp over NCCL); rank 0 also runs the whole mesh on a single-GPU handle and the
per-iteration compliance / density / iteration counts are compared.  With DIST_BIG=NXxNY a large mesh is timed as
well (graph-captured PCG, then eager).  Prints one JSON line on rank 0; exit code 0 = all checks passed."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
from solidspy_opt import _capi                                   # noqa: E402
from solidspy_opt.distributed import SlabSimp                     # noqa: E402
from solidspy_opt.optimize import kloc_simp                       # noqa: E402


def model(nx, ny):
    """Cantilever of beam(n=1) without building the element table: column 0 clamped, unit load at mid-height of the
    free end."""
    cons = np.zeros((ny + 1, nx + 1, 2), dtype=np.int8)
    cons[:, 0, :] = -1
    loads = np.array([[(nx + 1)*(ny//2) + nx, 0.0, -1.0]])
    return cons.reshape(-1, 2), loads, kloc_simp(206.8e9, 0.28)


def gather_rho(sim, world):
    mine = torch.as_tensor(sim.local_rho()).cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts).cpu().numpy().reshape(-1)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "ok": True}
    # ---- correctness against the single-GPU path -----------------------------------------------------------------
    steps = 4
    # (p2p with and without the one-launch exchanges, graph-captured and eager; NCCL graph-captured and eager)
    # fused: 2 = one-launch exchanges + in-kernel interface rows on the slab-local levels (default), 1 = one-launch exchanges
    # only, 0 = two-phase push / finish kernels
    # fused = 3 adds the persistent kernels for the small slab-local levels: they need a hierarchy with >= 3 slab-local levels,
    # so the second mesh (K = 4 on two slabs) is there for them
    cases = [(384, 64*world, t, g_, f) for t, g_, f in (("p2p", True, 2), ("p2p", False, 2), ("p2p", True, 1), ("p2p", True, 0),
                                                         ("nccl", True, 0), ("nccl", False, 0))]
    cases += [(1536, 384*world, t, g_, f) for t, g_, f in (("p2p", True, 3), ("p2p", False, 3), ("p2p", True, 2))]
    for nx, ny, transport, use_graph, fused in cases:
        cons, loads, ke = model(nx, ny)
        sim = SlabSimp(nx, ny, ke, cons, loads, rmin=2.5, dist=dist, device=local, rtol=1e-10, use_graph=use_graph,
                       transport=transport, fused=fused)
        g, hist = 0.0, []
        for _ in range(steps):
            g, res = sim.step(g)
            hist.append((res["compliance"], res["cg_iters"], res["oc_steps"], g))
        rho = gather_rho(sim, world)
        key = "%dx%d_" % (nx, ny) + transport + ("_fused%d" % fused if fused else "") + ("_graph" if use_graph else "_eager")
        out[key + "_used"] = {"graph": bool(sim.use_graph), "p2p": bool(sim.p2p), "fused": bool(sim.fused), "inkernel": bool(sim.inkernel),
                              "mid": bool(sim.mid), "K": sim.K}
        out[key + "_error"] = [sim.graph_error, sim.p2p_error]
        out["K"] = sim.K
        sim.close()
        if rank == 0:
            with _capi.Topo(nx, ny, ke, device=local) as full:
                full.set_cons(cons)
                full.set_loads(loads)
                full.fill(_capi.F_RHO, 0.5)
                params = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=2.5, dx=1.0, dy=1.0, move=0.2, rtol=1e-10,
                                          precond=_capi.PRECOND_GMG, maxit=2000, warm=1)
                g1, worst_c, its = 0.0, 0.0, []
                for k in range(steps):
                    g1, r = full.simp_step(params, g1)
                    worst_c = max(worst_c, abs(hist[k][0] - r.compliance)/abs(r.compliance))
                    its.append((r.cg_iters, hist[k][1]))
                    if hist[k][2] != r.oc_steps:
                        out["ok"] = False
                drho = float(np.abs(rho - full.get(_capi.F_RHO)).max())
            out[key] = {"compliance_rel": worst_c, "drho_max": drho, "cg_iters_single_vs_slabs": its}
            if worst_c > 1e-8 or drho > 1e-7:
                out["ok"] = False
    # ---- the reference's own trajectory (golden fixture of config 1, 60x60) through SIMP(dist=...) ------------------------
    gpath = os.path.join(ROOT, "tests", "golden", "simp_c1.npz")
    if world == 2 and os.path.exists(gpath):
        from solidspy_opt.optimize import SIMP
        gld = np.load(gpath, allow_pickle=True)
        hist = {}
        rho_mine = SIMP(float(gld["length"]), float(gld["height"]), int(gld["nx"]), int(gld["ny"]), gld["dirs"], gld["positions"],
                        int(gld["niter"]), int(gld["penal"]), history=hist, verbose=False, dist=dist, device=local)
        mine = torch.as_tensor(rho_mine).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        rho_all = torch.cat(parts).cpu().numpy()
        n = len(gld["compliance"])
        c_err = float(np.max(np.abs(np.array(hist["compliance"][:n]) - gld["compliance"])/np.abs(gld["compliance"])))
        r_err = float(np.abs(rho_all - gld["rho_final"]).max())
        out["golden_c1"] = {"iterations": len(hist["compliance"]), "expected": n, "compliance_rel": c_err, "rho_final_abs": r_err}
        if len(hist["compliance"]) != n or c_err > 1e-6 or r_err > 1e-6:
            out["ok"] = False
    # ---- timing on a large mesh ------------------------------------------------------------------------------------
    big = os.environ.get("DIST_BIG", "")
    if big:
        nx, ny = (int(v) for v in big.lower().split("x"))
        cons, loads, ke = model(nx, ny)
        for transport, use_graph, fused in (("p2p", True, 3), ("p2p", True, 2), ("p2p", True, 1), ("p2p", True, 0), ("nccl", True, 0)):
            sim = SlabSimp(nx, ny, ke, cons, loads, rmin=4.0, dist=dist, device=local, rtol=1e-9, use_graph=use_graph,
                           transport=transport, fused=fused)
            g = 0.0
            for _ in range(3):
                g, res = sim.step(g)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n, its = 5, []
            for _ in range(n):
                g, res = sim.step(g)
                its.append(res["cg_iters"])
            torch.cuda.synchronize()
            dist.barrier()
            dt = (time.perf_counter() - t0)/n
            key = "big_" + transport + ("_fused%d" % fused if fused else "")
            out[key] = {"mesh": "%dx%d" % (nx, ny), "ms_per_step": dt*1e3, "it_per_s": 1.0/dt, "cg_iters": its,
                        "graph_used": bool(sim.use_graph), "graph_error": sim.graph_error, "p2p": bool(sim.p2p),
                        "p2p_error": sim.p2p_error, "K": sim.K, "fused": bool(sim.fused),
                        "stage_ms": {k: round(v, 3) for k, v in res["stage_ms"].items()}}
            sim.close()
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
