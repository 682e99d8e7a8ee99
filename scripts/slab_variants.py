"""Row-slab SIMP on one mesh under several variants (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 scripts/slab_variants.py 8192 4096 "p2p,1,0 p2p,0,0 p2p,1,4"
A variant is transport,fused,K (K = 0: default first global level).  Device time (CUDA events on the slab stream), max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
from solidspy_opt import _capi                                   # noqa: E402
from solidspy_opt.distributed import SlabSimp                     # noqa: E402
from solidspy_opt.optimize import kloc_simp                       # noqa: E402


def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    variants = sys.argv[3].split() if len(sys.argv) > 3 else ["p2p,1,0", "p2p,0,0"]
    steps, warm = 5, 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cons = np.zeros((ny + 1, nx + 1, 2), dtype=np.int8)
    cons[:, 0, :] = -1
    loads = np.array([[(nx + 1)*(ny//4) - 1, 0.0, -1.0]])
    ke = kloc_simp(206.8e9, 0.28)
    for v in variants:
        transport, fused, K = v.split(",")
        try:
            sim = SlabSimp(nx, ny, ke, cons, loads, rmin=4.0, dist=dist, device=local, rtol=1e-9, maxit=100000, transport=transport,
                           fused=int(fused), K=int(K))
            g, its, stage, comp = 0.0, [], {}, []
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for k in range(warm + steps):
                if k == warm:
                    dist.barrier()
                    torch.cuda.synchronize()
                    with torch.cuda.stream(sim.stream):
                        ev0.record()
                g, r = sim.step(g)
                comp.append(r["compliance"])
                if k >= warm:
                    its.append(r["cg_iters"])
                    for n, x in r["stage_ms"].items():
                        stage[n] = stage.get(n, 0.0) + x/steps
            with torch.cuda.stream(sim.stream):
                ev1.record()
            torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1)/steps], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out = {"variant": v, "world": world, "ms_per_step": round(float(t.item()), 3), "cg_iters": its, "K": sim.K, "fused": sim.fused, "inkernel": sim.inkernel,
                   "p2p": sim.p2p, "graph": sim.use_graph, "graph_error": sim.graph_error,
                   "solve_us_per_pcg_iteration": round(1e3*stage["solve"]*steps/(sum(its) + steps), 1),
                   "stage_ms": {k: round(x, 3) for k, x in stage.items()}, "compliance_last": comp[-1]}
            if sim.fused:
                # cost of ONE exchange at this world size: 400 back-to-back one-launch exchanges of each kind (the data is junk afterwards)
                c, t0 = _capi, sim.slabs[0]
                for name, args in (("rows_fine", (c.X_AP, 0, 0, 0, 0)), ("rows_fine+2scalars", (c.X_AP, 0, c.S_RR_RZ, 2, 0)),
                                   ("rows_level%d" % (sim.K - 1), (c.X_LEVEL_B, sim.K - 1, 0, 0, 0)), ("gather+tail", None)):
                    def burst(n):
                        for _ in range(n):
                            if args is None:
                                t0.dist_call("p2p_gather_tail")
                            else:
                                t0.dist_call("p2p_fused", *args)
                    dist.barrier()
                    torch.cuda.synchronize()
                    gr = torch.cuda.CUDAGraph()                       # 50 exchanges per graph: no host launch cost in the figure
                    with torch.cuda.graph(gr, stream=sim.stream, capture_error_mode="relaxed"):
                        burst(50)
                    with torch.cuda.stream(sim.stream):
                        gr.replay()
                        ev0.record()
                        for _ in range(8):
                            gr.replay()
                        ev1.record()
                    torch.cuda.synchronize()
                    out.setdefault("xchg_us_in_graph", {})[name] = round(ev0.elapsed_time(ev1)*1e3/400, 2)
                    del gr
            sim.close()
        except Exception as exc:
            out = {"variant": v, "error": repr(exc)}
        if rank == 0:
            print(json.dumps(out), flush=True)
        dist.barrier()
    if os.environ.get("SLAB_PROFILE"):
        # where does a PCG iteration go on THIS world size: eager launches, every launch bracketed by CUDA events (kernels
        # that wait for a neighbour include the wait)
        transport, fused, K = variants[0].split(",")
        sim = SlabSimp(nx, ny, ke, cons, loads, rmin=4.0, dist=dist, device=local, rtol=1e-9, maxit=100000, transport=transport,
                       fused=int(fused), K=int(K), use_graph=False)
        g, its = 0.0, 0
        t0 = sim.slabs[0]
        for k in range(warm + 2):
            if k == warm:
                t0.profile_matvec(True)
            g, r = sim.step(g)
            if k >= warm:
                its += r["cg_iters"] + 1
        rep = sorted(t0.profile_report(), key=lambda x: -x[2])
        t0.profile_matvec(False)
        tot = sum(x[2] for x in rep)
        if rank == 0:
            print("# rank 0, eager + event brackets, %d PCG iterations (incl. initial passes): %.1f us of kernel time per iteration" % (its, tot*1e3/its))
            for nm, n, ms in rep[:28]:
                print("#   %-28s n %5d  avg %8.2f us  per PCG it %8.2f us" % (nm, n, ms*1e3/n, ms*1e3/its), flush=True)
        sim.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
