"""CPU restatement (numpy/scipy) of the reference DiffusionNetBlock hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the *checker* (or the CPU arm being timed) -- the product
path in ``diffusion-net_b200/`` never imports anything under ``oracle/``.

Parity status: PINNED.  The reference (nmwsharp/diffusion-net @ b1019b0) ships
no tests or golden vectors for this path, so the pin is
against outputs of the *unmodified reference modules run live* in the build
container: ``oracle/make_golden.py`` imports ``/root/reference/src/diffusion_net``
(with the two absent native deps stubbed; they are not on this path), runs it on
seeded inputs and commits the inputs + fp32/fp64 outputs under ``tests/golden/``.
``tests/test_oracle.py`` checks every function below against those fixtures.
The data-side neighbours of the path (HKS features, the operator-cache reader) are pinned
the same way by ``oracle/make_golden_geom.py`` -> ``tests/golden/geom_small.npz`` + ``tests/golden/op_cache/``.

Every function works in the dtype of its inputs (float32 reproduces the
reference arithmetic order with numpy kernels; float64 is the gold standard the
CUDA path is compared with).  All reference citations are to
``/root/reference/src/diffusion_net/``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "to_basis", "from_basis", "learned_time_diffusion", "grad_spmm",
    "spatial_gradient_features", "mini_mlp", "diffusion_net_block",
    "diffusion_net", "coo_to_csr", "rel_err", "compute_hks", "hks_autoscale_scales", "cache_key",
    "read_operator_cache",
]


def rel_err(a, b):
    """max|a-b| / max|b| -- the error metric of SURVEY.md section 8(c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (denom if denom > 0 else 1.0))


def coo_to_csr(rows, cols, vals, shape):
    """scipy CSR from the coalesced COO triplets the reference hands over
    (utils.py:50-55 ``sparse_np_to_torch``: int64 indices, fp32 values)."""
    return sp.csr_matrix((np.asarray(vals), (np.asarray(rows), np.asarray(cols))), shape=shape)


def to_basis(values, basis, massvec):
    """geometry.py:572-583: ``basis^T @ (values * massvec[..., None])``.

    values (V,C), basis (V,K), massvec (V,) -> (K,C)."""
    return basis.T @ (values * massvec[:, None])


def from_basis(values, basis):
    """geometry.py:586-598 (real branch, :598): ``basis @ values``.

    values (K,C), basis (V,K) -> (V,C)."""
    return basis @ values


def learned_time_diffusion(x, mass, evals, evecs, diffusion_time):
    """layers.py:44-67 + :90, ``method='spectral'``.

    Returns ``(x_diffuse, clamped_time)``: the reference overwrites the
    Parameter with ``clamp(t, min=1e-8)`` at every forward (layers.py:48-49)."""
    t = np.maximum(diffusion_time, np.asarray(1e-8, dtype=diffusion_time.dtype))
    if x.shape[-1] != t.shape[0]:  # layers.py:51-54
        raise ValueError(
            "Tensor has wrong shape = {}. Last dim shape should have number of channels = {}".format(
                x.shape, t.shape[0]))
    x_spec = to_basis(x, evecs, mass)                       # layers.py:59
    coefs = np.exp(-evals[:, None] * t[None, :])            # layers.py:62-63
    x_diffuse_spec = coefs * x_spec                         # layers.py:64
    return from_basis(x_diffuse_spec, evecs), t             # layers.py:67


def grad_spmm(gradX, gradY, x_diffuse):
    """layers.py:216-223: ``mm(gradX, x)``, ``mm(gradY, x)`` stacked on a last
    axis of size 2.  gradX/gradY are scipy CSR (V,V); returns (V,C,2)."""
    gx = gradX @ x_diffuse
    gy = gradY @ x_diffuse
    return np.stack((gx, gy), axis=-1)


def spatial_gradient_features(vectors, A_re=None, A_im=None, A=None):
    """layers.py:117-130.  ``vectors`` (V,C,2); weights are ``nn.Linear.weight``
    matrices (C,C) applied as ``g @ W.T`` (bias-free, layers.py:110-113)."""
    g0 = vectors[..., 0]
    g1 = vectors[..., 1]
    if A is None:                                            # with_gradient_rotations
        b_re = g0 @ A_re.T - g1 @ A_im.T                     # layers.py:122
        b_im = g1 @ A_re.T + g0 @ A_im.T                     # layers.py:123
    else:
        b_re = g0 @ A.T                                      # layers.py:125
        b_im = g1 @ A.T                                      # layers.py:126
    dots = g0 * b_re + g1 * b_im                             # layers.py:128
    return np.tanh(dots)                                     # layers.py:130


def mini_mlp(x, weights, biases):
    """layers.py:133-164 in eval mode (dropout = identity): Linear+ReLU for all
    but the last layer, which has no activation (layers.py:158-164)."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = x @ w.T + b
        if i + 1 < n:
            x = np.maximum(x, 0)
    return x


def diffusion_net_block(x_in, mass, evals, evecs, gradX, gradY, params,
                        with_gradient_features=True, return_intermediates=False):
    """layers.py:200-241 for one mesh (no batch dim).

    ``params`` keys follow the reference state_dict names relative to the block:
    ``diffusion.diffusion_time``; ``gradient_features.A_re.weight`` /
    ``gradient_features.A_im.weight`` (or ``gradient_features.A.weight``);
    ``mlp.miniMLP_mlp_layer_00{i}.weight`` / ``.bias``."""
    C = params["diffusion.diffusion_time"].shape[0]
    if x_in.shape[-1] != C:  # layers.py:204-207
        raise ValueError(
            "Tensor has wrong shape = {}. Last dim shape should have number of channels = {}".format(
                x_in.shape, C))
    x_diffuse, _ = learned_time_diffusion(x_in, mass, evals, evecs,
                                          params["diffusion.diffusion_time"])   # :210
    inter = {"x_diffuse": x_diffuse}
    if with_gradient_features:
        x_grad = grad_spmm(gradX, gradY, x_diffuse)                            # :216-223
        if "gradient_features.A.weight" in params:
            feats = spatial_gradient_features(x_grad, A=params["gradient_features.A.weight"])
        else:
            feats = spatial_gradient_features(
                x_grad, A_re=params["gradient_features.A_re.weight"],
                A_im=params["gradient_features.A_im.weight"])                  # :226
        inter["x_grad"] = x_grad
        inter["x_grad_features"] = feats
        combined = np.concatenate((x_in, x_diffuse, feats), axis=-1)           # :229
    else:
        combined = np.concatenate((x_in, x_diffuse), axis=-1)                  # :232
    ws, bs = [], []
    i = 0
    while "mlp.miniMLP_mlp_layer_{:03d}.weight".format(i) in params:
        ws.append(params["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)])
        bs.append(params["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)])
        i += 1
    out = mini_mlp(combined, ws, bs) + x_in                                     # :236-239
    if return_intermediates:
        return out, inter
    return out


def diffusion_net(x_in, mass, evals, evecs, gradX, gradY, params, n_block,
                  outputs_at="vertices", faces=None, edges=None,
                  with_gradient_features=True):
    """layers.py:314-407 for one mesh, ``last_activation=None``.

    ``params`` uses the full reference state_dict names (``first_lin.weight``,
    ``block_{i}.…``, ``last_lin.weight``)."""
    x = x_in @ params["first_lin.weight"].T + params["first_lin.bias"]          # :366
    for b in range(n_block):                                                    # :369-370
        pre = "block_{}.".format(b)
        bp = {k[len(pre):]: v for k, v in params.items() if k.startswith(pre)}
        x = diffusion_net_block(x, mass, evals, evecs, gradX, gradY, bp,
                                with_gradient_features=with_gradient_features)
    x = x @ params["last_lin.weight"].T + params["last_lin.bias"]               # :373
    if outputs_at == "vertices":
        return x
    if outputs_at == "edges":                                                   # :379-384
        return x[edges].mean(axis=1)
    if outputs_at == "faces":                                                   # :386-391
        return x[faces].mean(axis=1)
    if outputs_at == "global_mean":                                             # :393-397
        return (x * mass[:, None]).sum(axis=0) / mass.sum()
    raise ValueError("invalid setting for outputs_at")


# ------------------------------------------------------------------------------------------------
# data-side neighbours of the block
# ------------------------------------------------------------------------------------------------
def compute_hks(evals, evecs, scales):
    """geometry.py:600-628: (K),(V,K),(S) -> (V,S), ``sum_k exp(-evals[k] scales[s]) evecs[v,k]^2``."""
    power_coefs = np.exp(-evals[None, :] * scales[:, None])                    # :619  (S,K)
    return (evecs * evecs) @ power_coefs.T                                     # :620-622 (the "could be a matmul")


def hks_autoscale_scales(count, dtype=np.float32):
    """geometry.py:632: ``torch.logspace(-2, 0, steps=count)``."""
    return np.logspace(-2.0, 0.0, num=count).astype(dtype)


def cache_key(verts, faces):
    """utils.py:71-76 via geometry.py:450: sha1 over the raw bytes of verts then faces."""
    import hashlib
    h = hashlib.sha1()
    for a in (verts, faces):
        h.update(np.ascontiguousarray(a).view(np.uint8))
    return h.hexdigest()


def read_operator_cache(npz, k_eig):
    """geometry.py:494-519 (the cache-hit branch): CSC triples -> scipy matrices, spectrum truncated to k_eig.
    Returns (frames, mass, L, evals, evecs, gradX, gradY) with the sparse ones as scipy CSR."""
    def mat(prefix):                                                           # :494-500
        shape = tuple(int(v) for v in npz[prefix + "_shape"])
        return sp.csc_matrix((npz[prefix + "_data"], npz[prefix + "_indices"], npz[prefix + "_indptr"]),
                             shape=shape).tocsr()
    return (npz["frames"], npz["mass"], mat("L"), npz["evals"][:k_eig], npz["evecs"][:, :k_eig],
            mat("gradX"), mat("gradY"))


def edge_tangent_vectors(verts, frames, edges):
    """geometry.py:198-207: tangent-plane coordinates of every edge vector in the frame of its tail vertex.
    verts (V,3), frames (V,3,3) with rows (basisX, basisY, normal), edges (2,E) -> (E,2), in verts' dtype."""
    edge_vecs = verts[edges[1]] - verts[edges[0]]
    bx, by = frames[edges[0], 0, :], frames[edges[0], 1, :]
    return np.stack(((edge_vecs * bx).sum(-1), (edge_vecs * by).sum(-1)), axis=-1)


def build_grad(n_verts, edges, edge_tangent):
    """geometry.py:209-273: per vertex, the least-squares gradient of a scalar from its outgoing edges,
    ``(lhs^T lhs + 1e-5 I)^-1 lhs^T`` applied to value differences; row v of the result holds the complex coefficient of
    v itself (minus the sum of the others) and of every neighbour.  fp64 like the reference (np.zeros / np.linalg.inv).
    Returns scipy CSR complex128 (V,V); duplicate (row, col) pairs are summed as coo_matrix -> tocsc() does."""
    edges = np.asarray(edges)
    et = np.asarray(edge_tangent, dtype=np.float64)
    keep = edges[0] != edges[1]                                                # :228 (self loops are skipped)
    tail, tip, et = edges[0][keep], edges[1][keep], et[keep]
    order = np.argsort(tail, kind="stable")                                    # outgoing lists in edge order (:224-229)
    tail, tip, et = tail[order], tip[order], et[order]
    start = np.searchsorted(tail, np.arange(n_verts + 1))
    rows, cols, vals = [], [], []
    for v in range(n_verts):
        lhs = et[start[v]:start[v + 1]]                                        # (n,2)  :245-252
        inv = np.linalg.inv(lhs.T @ lhs + 1e-5 * np.identity(2)) @ lhs.T       # :255-256  (2,n)
        coef = inv[0] + 1j * inv[1]                                            # :258-259: column i+1 of sol_mat
        rows += [v] * (len(lhs) + 1)
        cols += [v] + list(tip[start[v]:start[v + 1]])
        vals += [-coef.sum()] + list(coef)                                     # column 0: rhs_mat[:,0] = -1
    return sp.coo_matrix((np.array(vals), (np.array(rows), np.array(cols))), shape=(n_verts, n_verts)).tocsr()
