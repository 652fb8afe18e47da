import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import meshio, params, solver1d
from gmpnp_b200._lib import NewtonOpts
m = meshio.load_mesh("1D_variable_1um_mesh_1090")
x = m.x[:, 0][:21].copy(); x = x / x[-1]
n = len(x)
Vs = [-0.5, -1.0]
plist = [params.params_1d(L_n=1e-6, voltage_multiplier=V) for V in Vs]
s = solver1d.Solver1D(x, batch=len(Vs))
s.set_params(plist)
o = NewtonOpts.reference_1d(); o.maxit = int(sys.argv[2])
u = torch.zeros(len(Vs), n, 7, dtype=torch.float64, device='cuda')
un = solver1d.bulk_state(len(Vs), n, 'cuda')
out = s.newton(u, un, o)
torch.cuda.synchronize()
print(out['iters'].tolist(), out['status'].tolist(), out['r0'].tolist(), out['r'].tolist(), float(u.abs().sum()))
np.save(sys.argv[1], u.cpu().numpy())
