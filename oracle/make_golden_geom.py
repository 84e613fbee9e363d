"""Generate the fixtures for the data-side neighbours of the block by running the
UNMODIFIED reference here (build container only; needs /root/reference):

    python oracle/make_golden_geom.py

* ``tests/golden/op_cache/<sha1>_0.npz`` -- a cache entry WRITTEN BY the reference's ``get_operators``
  (geometry.py:526-568) for ``torus(12,16)``, k_eig=16.  (potpourri3d's cotan_laplacian / vertex_areas are the
  numpy restatements in ``ref_import.py``: they only feed the mesh's operator VALUES, not the file format.)
* ``tests/golden/geom_small.npz`` -- what the reference's cache-HIT branch (geometry.py:494-519) returns for that
  entry (at k_eig=16 and truncated to 12), ``compute_hks_autoscale`` of it in fp32/fp64, and the fp64 output of a
  2-block reference net fed with those HKS features (the cache -> HKS -> net pipeline of the experiments,
  e.g. human_segmentation_original.py:111-126).

Separate from make_golden.py so the ARPACK-seeded fixtures that file wrote are not regenerated.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_import import import_reference  # noqa: E402
import diffusion_net_b200.synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def coo(prefix, A):
    A = A.coalesce()
    return {prefix + "_rows": A.indices()[0].numpy().astype(np.int32),
            prefix + "_cols": A.indices()[1].numpy().astype(np.int32), prefix + "_vals": A.values().numpy()}


def main():
    dn = import_reference()
    K, S = 16, 16
    verts, faces = syn.torus_mesh(12, 16, seed=4)
    verts = dn.geometry.normalize_positions(verts)
    cache_out = os.path.join(OUT, "op_cache")
    shutil.rmtree(cache_out, ignore_errors=True)
    os.makedirs(cache_out)
    with tempfile.TemporaryDirectory() as tmp:
        dn.geometry.get_operators(verts, faces, k_eig=K, op_cache_dir=tmp)            # miss: computes + writes
        files = sorted(os.listdir(tmp))
        assert len(files) == 1, files
        hit = dn.geometry.get_operators(verts, faces, k_eig=K, op_cache_dir=tmp)      # hit: reads the file back
        hit12 = dn.geometry.get_operators(verts, faces, k_eig=12, op_cache_dir=tmp)
        shutil.copy(os.path.join(tmp, files[0]), os.path.join(cache_out, files[0]))
    frames, mass, L, evals, evecs, gradX, gradY = hit
    fx = {"verts": verts.numpy(), "faces": faces.numpy(), "cache_file": np.array(files[0]),
          "frames": frames.numpy(), "mass": mass.numpy(), "evals": evals.numpy(), "evecs": evecs.numpy(),
          "evals12": hit12[3].numpy(), "evecs12": hit12[4].numpy()}
    fx.update(coo("L", L))
    fx.update(coo("gradX", gradX))
    fx.update(coo("gradY", gradY))
    hks32 = dn.geometry.compute_hks_autoscale(evals, evecs, S)
    hks64 = dn.geometry.compute_hks_autoscale(evals.double(), evecs.double(), S)
    fx["hks_f32"], fx["hks_f64"] = hks32.numpy(), hks64.numpy()
    sc = torch.tensor([0.05, 0.5, 2.0])
    fx["hks3_scales"] = sc.numpy()
    fx["hks3_f64"] = dn.geometry.compute_hks(evals.double(), evecs.double(), sc.double()).numpy()

    torch.manual_seed(5)
    net = dn.layers.DiffusionNet(C_in=S, C_out=6, C_width=32, N_block=2, dropout=False)
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for name, prm in net.named_parameters():
            if name.endswith("diffusion_time"):
                prm.copy_(1e-3 + 0.3 * torch.rand(prm.shape, generator=g))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    fx.update({"p:" + k: v.numpy() for k, v in sd.items()})
    n64 = dn.layers.DiffusionNet(C_in=S, C_out=6, C_width=32, N_block=2, dropout=False).double()
    n64.load_state_dict({k: v.double() for k, v in sd.items()})
    n64.eval()
    with torch.no_grad():
        d = lambda t: t.double()
        fx["net_out_f64"] = n64(hks64, d(mass), L=d(L), evals=d(evals), evecs=d(evecs), gradX=d(gradX),
                                gradY=d(gradY)).numpy()
    np.savez_compressed(os.path.join(OUT, "geom_small.npz"), **fx)
    for f in ("geom_small.npz", os.path.join("op_cache", files[0])):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
