"""Generates tests/golden/*.npz.  Run ON THE GPU BOX (the reference nerfacc kernels are CUDA-only):

    python tests/golden/make_golden.py gpurun_out/golden      # then copy the .npz files into tests/golden/

ref_*.npz hold outputs of the UNMODIFIED reference nerfacc CUDA extension (oracle/_ref/nerfacc_ref_C.so, compiled from
/root/reference by oracle/build_ref.py) on small seeded inputs: ray marching (CS/ray_marching.cu:194-289), patch weights
forward / backward (CS/render_weight.cu:502-543, :583-626) and CUB transmittance (CS/render_transmittance_cub.cu:111-134).
The CPU suite (tests/test_golden_fixtures.py) pins the oracle's C restatement to them bit for bit, so the oracle is
checked against the real reference even where no GPU is present.
oracle_*.npz hold oracle outputs for the paths whose reference is absent from the tree (tiny-cuda-nn hash grid, PyMCubes):
they pin the CUDA kernels to the oracle across toolchains, not the oracle to the reference (PARITY UNPINNED).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_march_case  # noqa: E402
import oracle  # noqa: E402
from oracle import build_ref, mc  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    ref = build_ref.load()
    assert ref is not None, "oracle/_ref/nerfacc_ref_C.so missing"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # ---- ray marching: small grids so the fixture stays small (grid stored bit-packed) ----
    for name, kw in (("shell", dict(seed=11, n_rays=192, res=(32, 32, 32), step=0.02, grid_kind="shell")),
                     ("random", dict(seed=12, n_rays=160, res=(24, 40, 16), step=0.007, grid_kind="random")),
                     ("fine", dict(seed=13, n_rays=96, res=(32, 32, 32), step=0.002, grid_kind="shell"))):
        c = make_march_case(**kw)
        args = [t(c[k]) for k in ("rays_o", "rays_d", "t_min", "t_max", "roi", "grid")]
        outs = {}
        for cone in (0.0, 0.004):
            p, i, t0, t1 = ref.ray_marching(*args, ref.ContractionType.AABB, float(c["step"]), cone)
            tag = "cone0" if cone == 0.0 else "cone4e-3"
            outs.update({f"packed_{tag}": p.cpu().numpy(), f"t0_{tag}": t0.cpu().numpy()[:, 0], f"t1_{tag}": t1.cpu().numpy()[:, 0]})
        np.savez_compressed(os.path.join(out_dir, f"ref_march_{name}.npz"), rays_o=c["rays_o"], rays_d=c["rays_d"], t_min=c["t_min"],
                            t_max=c["t_max"], roi=c["roi"], grid_bits=np.packbits(c["grid"].reshape(-1)), res=np.array(c["grid"].shape),
                            step=c["step"], **outs)
    # ---- patch weights fwd/bwd + CUB transmittance ----
    rng = np.random.RandomState(5)
    for P in (1, 9):
        counts = rng.randint(0, 24, size=96)
        counts[::7] = 0
        idx = np.repeat(np.arange(96), counts)
        S = idx.size
        a = rng.uniform(0, 1, (S, P, 1)).astype(np.float32)
        a[rng.randint(0, S, 20)] = 1.0
        a[rng.randint(0, S, 20)] = 0.0
        g = rng.randn(S, P, 1).astype(np.float32)
        num = np.bincount(idx, minlength=96).astype(np.int32)
        packed = np.stack([np.cumsum(num) - num, num], -1).astype(np.int32)
        if P == 1:
            w = ref.weight_from_alpha_forward_naive(t(packed), t(a.reshape(S, 1)))
            ga = ref.weight_from_alpha_backward_naive(w, t(g.reshape(S, 1)), t(packed), t(a.reshape(S, 1)))
            T = ref.transmittance_from_alpha_forward_cub(t(idx.astype(np.int64)), t(a.reshape(S, 1)))
            extra = dict(transmittance_cub=T.cpu().numpy())
        else:
            w = ref.weight_from_alpha_patch_based_forward_naive(t(packed), t(a))
            ga = ref.weight_from_alpha_patch_based_backward_naive(w, t(g), t(packed), t(a))
            extra = {}
        np.savez_compressed(os.path.join(out_dir, f"ref_weights_P{P}.npz"), packed_info=packed, alphas=a, grad_weights=g,
                            weights=w.cpu().numpy().reshape(S, P, 1), grad_alphas=ga.cpu().numpy().reshape(S, P, 1), **extra)
    # ---- oracle-generated (reference absent): hash grid + marching cubes ----
    spec = oracle.hashgrid_spec(n_levels=6, log2_hashmap_size=12, base_resolution=8, per_level_scale=1.5)
    rs = np.random.RandomState(7)
    x = rs.uniform(-1, 1, (257, 3)).astype(np.float32)
    table = (rs.uniform(-1, 1, spec.n_params) * 0.1).astype(np.float16)
    feats = oracle.hashgrid_fwd(spec, x, table)
    np.savez_compressed(os.path.join(out_dir, "oracle_hashgrid_small.npz"), x=x, table=table, feats=feats, n_levels=6, log2_hashmap_size=12,
                        base_resolution=8, per_level_scale=1.5)
    g = np.linspace(-1, 1, 20).astype(np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    u = (0.6 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2) + 0.05 * rs.randn(20, 20, 20)).astype(np.float32)
    v, tr, n_main, _ = mc.marching_cubes(u)
    np.savez_compressed(os.path.join(out_dir, "oracle_mc_bumpy20.npz"), u=u, vertices=v, triangles=tr, n_main=n_main)
    print("wrote", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
