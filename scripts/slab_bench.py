"""Strong-scaling measurement of the SIMP design iteration on one mesh (run under torchrun with N ranks, or plainly
with one process for the single-GPU reference):
    python scripts/slab_bench.py 8192 4096 [steps]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 scripts/slab_bench.py 8192 4096
Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
from solidspy_opt import _capi                                   # noqa: E402
from solidspy_opt.distributed import SlabSimp                     # noqa: E402
from solidspy_opt.optimize import kloc_simp                       # noqa: E402


def model(nx, ny):
    cons = np.zeros((ny + 1, nx + 1, 2), dtype=np.int8)
    cons[:, 0, :] = -1
    loads = np.array([[(nx + 1)*(ny//2) + nx, 0.0, -1.0]])
    return cons.reshape(-1, 2), loads, kloc_simp(206.8e9, 0.28)


def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    warm = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cons, loads, ke = model(nx, ny)
    out = {"mesh": "%dx%d" % (nx, ny), "world": world, "steps": steps}
    if world == 1:
        with _capi.Topo(nx, ny, ke) as topo:
            topo.set_cons(cons)
            topo.set_loads(loads)
            topo.fill(_capi.F_RHO, 0.5)
            params = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-9,
                                      precond=_capi.PRECOND_GMG, maxit=2000, warm=1)
            g, its, stage = 0.0, [], {}
            for k in range(warm + steps):
                if k == warm:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                g, r = topo.simp_step(params, g)
                if k >= warm:
                    its.append(r.cg_iters)
                    t = topo.timing()
                    for n in ("setup_ms", "solve_ms", "sens_ms", "filter_ms", "oc_ms"):
                        stage[n[:-3]] = stage.get(n[:-3], 0.0) + getattr(t, n)/steps
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0)/steps
        out.update(ms_per_step=dt*1e3, it_per_s=1.0/dt, cg_iters=its, stage_ms=stage, path="single handle")
        print(json.dumps(out))
        return
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sim = SlabSimp(nx, ny, ke, cons, loads, rmin=4.0, dist=dist, device=local, rtol=1e-9,
                   use_graph=os.environ.get("TOPO_DIST_GRAPH", "1") != "0",
                   transport=os.environ.get("TOPO_DIST_TRANSPORT", "p2p"))
    g, its, stage = 0.0, [], {}
    for k in range(warm + steps):
        if k == warm:
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c0 = sim.collectives
        g, r = sim.step(g)
        if k >= warm:
            its.append(r["cg_iters"])
            for n, v in r["stage_ms"].items():
                stage[n] = stage.get(n, 0.0) + v/steps
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0)/steps
    out.update(ms_per_step=dt*1e3, it_per_s=1.0/dt, cg_iters=its, stage_ms=stage, K=sim.K, graph=bool(sim.use_graph),
               graph_error=sim.graph_error, p2p=bool(sim.p2p), p2p_error=sim.p2p_error,
               path="row slabs, %s transport" % ("peer-memory" if sim.p2p else "NCCL"), host_collective_calls_per_step=(sim.collectives - c0)/steps)
    if rank == 0:
        print(json.dumps(out))
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
