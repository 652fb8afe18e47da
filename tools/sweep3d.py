#!/usr/bin/env python
"""BASELINE config 4: all present pore meshes x wall voltages, sharded by (mesh, voltage) point across the GPUs of a box.

    python tools/sweep3d.py --voltages 16 --vmax -2.5                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/sweep3d.py --voltages 256 --vmax -12.5

Every rank solves its shard (batched per mesh, ramped pseudo-time march to steady state, failed points parked and
reported) and the per-point summaries are gathered once at the end.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import sweep, sweep3d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voltages", type=int, default=16)
ap.add_argument("--vmax", type=float, default=-2.5)
ap.add_argument("--dv", type=float, default=0.5)
ap.add_argument("--max-steps", type=int, default=40)
ap.add_argument("--meshes", nargs="*", default=None, help="subset of mesh stems (default: all 11)")
ap.add_argument("--vlimit", type=float, default=0.0,
                help="> 0: only the grid points with |V| <= vlimit are solved (the solvable sub-grid: beyond ~3 V_T the "
                     "under-resolved discrete 3D problem itself breaks down, DESIGN.md); the others are listed as not attempted")
ap.add_argument("--out", default=None, help="write the full per-point table (JSON) here")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
meshes = [m for m in sweep3d.CONFIG4_MESHES if a.meshes is None or m[0] in a.meshes]
grid = sweep3d.config4_points(a.voltages, a.vmax, meshes)
pts = [p for p in grid if a.vlimit <= 0 or abs(p.V) <= a.vlimit + 1e-12]
for i, p in enumerate(pts):
    p.index = i
mine = sweep3d.shard(pts, rank, world)
torch.cuda.synchronize()
t0 = time.time()
sw = sweep3d.Sweep3D(mine, device=local, dv_max=a.dv, max_steps=a.max_steps)
res = sw.solve()
torch.cuda.synchronize()
wall = time.time() - t0
idx = torch.tensor([p.index for p in mine], dtype=torch.int64, device=dev)
table = sweep.gather_results(torch.as_tensor(res, device=dev), idx, len(pts), world).cpu().numpy()
t = torch.tensor([wall], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ok = table[:, 0] == 0
    per_mesh = {}
    for m in meshes:
        sel = np.array([p.index for p in pts if p.mesh == m[0]], dtype=np.int64)
        okm = table[sel, 0] == 0
        per_mesh[m[0]] = {"points": int(len(sel)), "converged": int(okm.sum()),
                          "parked": [{"V": pts[i].V, "status": int(table[i, 0])} for i in sel[~okm]],
                          "newton_iterations_min_mean_max": [float(table[sel[okm], 2].min()), float(table[sel[okm], 2].mean()),
                                                             float(table[sel[okm], 2].max())] if okm.any() else None,
                          "pseudo_time_steps_max": float(table[sel[okm], 1].max()) if okm.any() else None}
    if a.out:
        json.dump({"columns": ["status", "pseudo_time_steps", "newton_iterations", "median_OH", "median_HCO3",
                               "median_CO32", "median_cation", "co2_entry_scaled", "max_cation"],
                   "points": [{"mesh": p.mesh, "V": p.V} for p in pts], "table": table.tolist()}, open(a.out, "w"))
    print(json.dumps({"workload": f"config4: {len(meshes)} pore meshes x {a.voltages}-point voltage grid down to {a.vmax} V_T"
                                  + (f", solved for |V| <= {a.vlimit} V_T ({len(pts)} of {len(grid)} grid points)" if a.vlimit > 0 else ""),
                      "n_gpus": world, "per_mesh": per_mesh,
                      "points": len(pts), "converged": int(ok.sum()), "failed_parked": int((~ok).sum()),
                      "wall_s_max_over_ranks": float(t[0]), "steady_solves_per_s": float(ok.sum() / float(t[0])),
                      "newton_iterations_mean": float(table[ok, 2].mean()) if ok.any() else None,
                      "largest_converged_abs_V_per_mesh": {m[0]: float(max([abs(p.V) for p in pts if p.mesh == m[0] and table[p.index, 0] == 0], default=0.0)) for m in meshes}}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
