"""Import the UNMODIFIED reference (``/root/reference/src/diffusion_net``) in the
build container.

TEST INFRASTRUCTURE ONLY (see oracle/dn_oracle.py header).  ``/root/reference``
exists only in the build container, never on the GPU box; callers must check
``reference_available()`` first.  Nothing here copies reference source: the
package is imported from where it lies.

``diffusion_net/geometry.py:17-18`` imports ``potpourri3d`` and
``robust_laplacian`` at module scope; neither is installed here and neither is
touched by the DiffusionNetBlock hot path.  We register stub modules for them.
To let the reference's own ``get_operators()`` run on synthetic triangle meshes
(for *test-input generation only*, never as a parity oracle) the ``potpourri3d``
stub carries numpy restatements of the two functions called at
``geometry.py:322-323`` (``cotan_laplacian`` and ``vertex_areas``;
potpourri3d==0.0.3 per ``environment.yml``).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import scipy.sparse as sp

REFERENCE_SRC = "/root/reference/src"
# staged copy made by oracle/stage_ref.py (git-ignored, travels to the GPU box with the snapshot): lets
# bench.py time the UNMODIFIED reference modules there, where /root/reference does not exist
STAGED_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_src():
    """Directory holding the reference ``diffusion_net`` package: the mounted reference, else the staged copy."""
    for d in (REFERENCE_SRC, STAGED_SRC):
        if os.path.isfile(os.path.join(d, "diffusion_net", "layers.py")):
            return d
    return None


def reference_available() -> bool:
    return reference_src() is not None


def _cotan_laplacian(V, F, denom_eps=0.0):
    """Positive semi-definite cotan Laplacian, L_ij = -1/2 (cot a_ij + cot b_ij),
    L_ii = -sum_j L_ij; cot = (u.v) / (|u x v| + denom_eps) per triangle corner."""
    V = np.asarray(V, dtype=np.float64)
    F = np.asarray(F)
    n = V.shape[0]
    rows, cols, vals = [], [], []
    for c in range(3):
        i, j, k = F[:, c], F[:, (c + 1) % 3], F[:, (c + 2) % 3]   # corner at i, opposite edge (j,k)
        u = V[j] - V[i]
        v = V[k] - V[i]
        cot = np.einsum("ij,ij->i", u, v) / (np.linalg.norm(np.cross(u, v), axis=1) + denom_eps)
        w = 0.5 * cot
        rows += [j, k, j, k]
        cols += [k, j, j, k]
        vals += [-w, -w, w, w]
    L = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return L.tocsc()


def _vertex_areas(V, F):
    """Lumped (barycentric) vertex areas: one third of each incident face area."""
    V = np.asarray(V, dtype=np.float64)
    F = np.asarray(F)
    fa = 0.5 * np.linalg.norm(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), axis=1)
    out = np.zeros(V.shape[0])
    for c in range(3):
        np.add.at(out, F[:, c], fa / 3.0)
    return out


def import_reference():
    """Returns the reference ``diffusion_net`` package (layers, geometry, utils)."""
    src = reference_src()
    if src is None:
        raise RuntimeError("reference not present at {} nor staged under {}".format(REFERENCE_SRC, STAGED_SRC))
    if "potpourri3d" not in sys.modules:
        pp3d = types.ModuleType("potpourri3d")
        pp3d.cotan_laplacian = _cotan_laplacian
        pp3d.vertex_areas = _vertex_areas
        sys.modules["potpourri3d"] = pp3d
    if "robust_laplacian" not in sys.modules:
        sys.modules["robust_laplacian"] = types.ModuleType("robust_laplacian")
    if src not in sys.path:
        sys.path.insert(0, src)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import diffusion_net  # noqa: F401  (the reference, unmodified)
    return sys.modules["diffusion_net"]
