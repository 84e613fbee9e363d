"""Pins oracle/dn_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden, golden_params, ROOT, GOLDEN

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dn_oracle as O  # noqa: E402


def ops_of(fx, dtype, prefix=""):
    V = fx[prefix + "mass"].shape[0]
    rows, cols = fx[prefix + "g_rows"].astype(np.int64), fx[prefix + "g_cols"].astype(np.int64)
    gX = O.coo_to_csr(rows, cols, fx[prefix + "gx_vals"].astype(dtype), (V, V))
    gY = O.coo_to_csr(rows, cols, fx[prefix + "gy_vals"].astype(dtype), (V, V))
    return (fx[prefix + "mass"].astype(dtype), fx[prefix + "evals"].astype(dtype),
            fx[prefix + "evecs"].astype(dtype), gX, gY)


@pytest.mark.parametrize("name,kw", [("block_small", {}), ("block_norot", {}),
                                     ("block_nograd", {"with_gradient_features": False})])
def test_block_matches_reference(name, kw):
    base = load_golden("block_small")
    fx = load_golden(name)
    for tag, dt, tol in (("f64", np.float64, 1e-12), ("f32", np.float32, 2e-6)):
        mass, evals, evecs, gX, gY = ops_of(base, dt)
        p = golden_params(fx, dt)
        out, inter = O.diffusion_net_block(fx["x_in"].astype(dt), mass, evals, evecs, gX, gY, p,
                                           return_intermediates=True, **kw)
        assert O.rel_err(inter["x_diffuse"], fx["x_diffuse_" + tag]) < tol
        if "x_grad_features_" + tag in fx:
            assert O.rel_err(inter["x_grad_features"], fx["x_grad_features_" + tag]) < tol * 20
        assert O.rel_err(out, fx["out_" + tag]) < tol
        assert out.dtype == dt


def test_clamp_matches_reference():
    fx = load_golden("block_small")
    t = golden_params(fx)["diffusion.diffusion_time"]
    assert t[3] < 0
    mass, evals, evecs, _, _ = ops_of(fx, np.float32)
    _, tc = O.learned_time_diffusion(fx["x_in"], mass, evals, evecs, t)
    np.testing.assert_array_equal(tc, fx["time_after_f32"])
    assert tc[3] == np.float32(1e-8)


def test_k128_block():
    fx = load_golden("block_k128")
    mass, evals, evecs, gX, gY = ops_of(fx, np.float64)
    out, inter = O.diffusion_net_block(fx["x_in"].astype(np.float64), mass, evals, evecs, gX, gY,
                                       golden_params(fx, np.float64), return_intermediates=True)
    assert O.rel_err(inter["x_diffuse"], fx["x_diffuse_f64_as32"]) < 2e-7
    assert O.rel_err(out, fx["out_f64_as32"]) < 2e-7


@pytest.mark.parametrize("mode", ["vertices", "edges", "faces", "global_mean"])
def test_net_matches_reference(mode):
    fx = load_golden("net_small")
    mass, evals, evecs, gX, gY = ops_of(fx, np.float64, "m0_")
    out = O.diffusion_net(fx["verts0"].astype(np.float64), mass, evals, evecs, gX, gY,
                          golden_params(fx, np.float64), n_block=2, outputs_at=mode,
                          faces=fx["faces"].astype(np.int64), edges=fx["edges"].astype(np.int64))
    assert O.rel_err(out, fx["out_{}_f64".format(mode)]) < 1e-12


def test_batch_equals_per_mesh():
    fx = load_golden("net_small")
    p = golden_params(fx, np.float64)
    for b, pre in enumerate(("m0_", "m1_")):
        mass, evals, evecs, gX, gY = ops_of(fx, np.float64, pre)
        out = O.diffusion_net(fx["verts{}".format(b)].astype(np.float64), mass, evals, evecs, gX, gY, p, n_block=2)
        assert O.rel_err(out, fx["out_batch2_f64"][b]) < 1e-12


def test_wrong_channels_raises():
    fx = load_golden("block_small")
    mass, evals, evecs, gX, gY = ops_of(fx, np.float32)
    with pytest.raises(ValueError):
        O.diffusion_net_block(fx["x_in"][:, :5], mass, evals, evecs, gX, gY, golden_params(fx))


def test_torch_port_matches_reference():
    """The torch-CPU port timed by bench.py reproduces the live-reference outputs."""
    import torch
    import dn_oracle_torch as T
    for name, kw in (("block_small", {}), ("block_norot", {}), ("block_nograd", {"with_gradient_features": False})):
        base = load_golden("block_small")
        fx = load_golden(name)
        V = base["mass"].shape[0]
        idx = torch.from_numpy(np.stack((base["g_rows"], base["g_cols"])).astype(np.int64))
        gX = torch.sparse_coo_tensor(idx, torch.from_numpy(base["gx_vals"]), (V, V)).coalesce()
        gY = torch.sparse_coo_tensor(idx, torch.from_numpy(base["gy_vals"]), (V, V)).coalesce()
        p = {k: torch.from_numpy(v) for k, v in golden_params(fx).items()}
        b = lambda a: torch.from_numpy(a).unsqueeze(0)
        out = T.block_forward(b(fx["x_in"]), b(base["mass"]), b(base["evals"]), b(base["evecs"]), [gX], [gY], p, **kw)
        assert O.rel_err(out[0].numpy(), fx["out_f32"]) < 1e-6
        assert O.rel_err(out[0].numpy(), fx["out_f64"]) < 2e-6


# ---- data-side neighbours: pinned by oracle/make_golden_geom.py -----------------------
def _cache_file():
    d = os.path.join(GOLDEN, "op_cache")
    files = sorted(os.listdir(d))
    assert len(files) == 1
    return os.path.join(d, files[0])


def test_hks_matches_reference():
    fx = load_golden("geom_small")
    sc = O.hks_autoscale_scales(16)
    got32 = O.compute_hks(fx["evals"], fx["evecs"], sc)
    assert O.rel_err(got32, fx["hks_f32"]) < 2e-6
    got64 = O.compute_hks(fx["evals"].astype(np.float64), fx["evecs"].astype(np.float64),
                          np.logspace(-2.0, 0.0, num=16))
    assert O.rel_err(got64, fx["hks_f64"]) < 1e-12
    got3 = O.compute_hks(fx["evals"].astype(np.float64), fx["evecs"].astype(np.float64),
                         fx["hks3_scales"].astype(np.float64))
    assert O.rel_err(got3, fx["hks3_f64"]) < 1e-12


def test_operator_cache_reader_matches_reference_hit_branch():
    fx = load_golden("geom_small")
    path = _cache_file()
    assert os.path.basename(path) == str(fx["cache_file"])
    assert os.path.basename(path) == O.cache_key(fx["verts"], fx["faces"]) + "_0.npz"
    with np.load(path, allow_pickle=True) as npz:
        frames, mass, L, evals, evecs, gX, gY = O.read_operator_cache(npz, 16)
        e12, v12 = O.read_operator_cache(npz, 12)[3:5]
    for got, key in ((frames, "frames"), (mass, "mass"), (evals, "evals"), (evecs, "evecs"),
                     (e12, "evals12"), (v12, "evecs12")):
        assert np.array_equal(got, fx[key]), key
    V = mass.shape[0]
    for got, pre in ((L, "L"), (gX, "gradX"), (gY, "gradY")):
        want = O.coo_to_csr(fx[pre + "_rows"].astype(np.int64), fx[pre + "_cols"].astype(np.int64),
                            fx[pre + "_vals"], (V, V))
        got = got.copy()
        got.sort_indices()
        want.sort_indices()
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices), pre
        assert np.array_equal(got.data, want.data), pre


def test_build_grad_matches_reference_fixture():
    """oracle.build_grad / edge_tangent_vectors (geometry.py:198-273) against gradX / gradY the live reference produced
    for tests/golden/geom_small.npz (edges = the Laplacian's pattern, geometry.py:333-335, 374-376)."""
    import scipy.sparse as sp
    fx = load_golden("geom_small")
    V = fx["verts"].shape[0]
    edges = np.stack((fx["L_rows"], fx["L_cols"])).astype(np.int64)
    et = O.edge_tangent_vectors(fx["verts"], fx["frames"], edges)
    G = O.build_grad(V, edges, et).tocsr()
    G.sum_duplicates(); G.sort_indices()
    for name, part in (("gradX", np.real), ("gradY", np.imag)):
        ref = sp.coo_matrix((fx[name + "_vals"], (fx[name + "_rows"], fx[name + "_cols"])), shape=(V, V)).tocsr()
        ref.sort_indices()
        assert np.array_equal(G.indptr, ref.indptr) and np.array_equal(G.indices, ref.indices)
        got = part(G.data).astype(np.float32)
        assert np.abs(got - ref.data).max() <= 2e-6 * np.abs(ref.data).max()
