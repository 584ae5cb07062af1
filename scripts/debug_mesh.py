import numpy as np, torch, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import mc
from supernormal_b200 import mesh
from test_gpu_mesh import _model
cuda = torch.device("cuda:0")
m = _model(cuda, n_active=0)
res = 96
u = mesh.extract_fields(m, [-1, -1, -1], [1, 1, 1], res)
print("field finite", torch.isfinite(u).all().item(), u.min().item(), u.max().item())
v, t, n_main = mesh.marching_cubes(u, 0.0)
vo, to, nmo, _ = mc.marching_cubes(u.cpu().numpy(), 0.0)
print(v.shape, vo.shape, t.shape, to.shape, n_main, nmo)
print("tri equal", np.array_equal(t.cpu().numpy(), to), "vert equal", np.array_equal(v.cpu().numpy(), vo))
print("oracle checks", mc.mesh_checks(vo, to), "cuda checks", mc.mesh_checks(v.cpu().numpy(), t.cpu().numpy()))
tt = to.astype(np.int64)
e = np.concatenate([tt[:, [0, 1]], tt[:, [1, 2]], tt[:, [2, 0]]]); und = np.sort(e, 1)
key = und[:, 0] * (tt.max() + 1) + und[:, 1]
_, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
be = und[cnt[inv] == 1]
print("boundary verts sample", vo[be[:8, 0]], vo[be[:8, 1]])
