import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gmpnp_b200 import meshio, params, solver1d, sweep
from gmpnp_b200._lib import NewtonOpts
x = meshio.load_mesh("1D_variable_50um_mesh_5990").x[:, 0]
prm = params.params_1d()
for parts in (2, 4, 8):
    s = solver1d.Solver1D(x, batch=1)
    s.set_params([prm])
    o = NewtonOpts.reference_1d(); o.partitions = parts
    for rep in range(2):
        u = torch.zeros(1, s.n, 7, dtype=torch.float64, device="cuda:0"); un = solver1d.bulk_state(1, s.n, "cuda:0")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.march(u, un, 100, o); e1.record(); torch.cuda.synchronize()
    print("config1 partitions", parts, "ms", e0.elapsed_time(e1), "its", int(out["iters"].sum()), flush=True)
    s.close()
pts = sweep.config2_points(256)
rules = {"all 2": 2, "all 8": 8, "long 8, rest 2": lambda n, b: 8 if n >= 4000 else 2,
         "long 8, rest 4": lambda n, b: 8 if n >= 4000 else 4, "long 8, mid 4, short 2": lambda n, b: 8 if n >= 4000 else (4 if n >= 1900 else 2),
         "auto": None}
for world in (8, 4, 2):
    mine = sweep.shard(pts, 0, world)
    for name, rule in rules.items():
        sw = sweep.Sweep1D(mine, device=0, dv_max=0.75, xtol_path=1.0, partitions=rule)
        sw.upload()
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); outs = sw.solve_resident(); e1.record(); torch.cuda.synchronize()
        summ = sw.summary(outs)
        print("shard 1/%d (%d points) %-24s: %.1f ms converged %d" % (world, len(mine), name, e0.elapsed_time(e1), summ["converged"]), flush=True)
        sw.close()
