"""``to_basis`` / ``from_basis`` with the reference signatures (geometry.py:572-598), running the
hand-written kernels.  Batched (B,V,*) or single-mesh (V,*) inputs, as the reference accepts."""
from __future__ import annotations

import torch

from . import ops


def to_basis(values, basis, massvec):
    """(B,V,D),(B,V,K),(B,V) -> (B,K,D): ``basis^T @ (values * massvec[...,None])`` (geometry.py:572-583)."""
    if values.dim() == 2:
        return ops.to_basis_raw(values, basis, massvec)
    return torch.stack([ops.to_basis_raw(values[b], basis[b], massvec[b]) for b in range(values.shape[0])], 0)


def from_basis(values, basis):
    """(B,K,D),(B,V,K) -> (B,V,D): ``basis @ values`` (geometry.py:586-598, real branch)."""
    if values.is_complex() or basis.is_complex():
        raise NotImplementedError("complex from_basis is dead code in the reference (utils.cmatmul does not exist)")
    if values.dim() == 2:
        return ops.from_basis_raw(values, basis)
    return torch.stack([ops.from_basis_raw(values[b], basis[b]) for b in range(values.shape[0])], 0)
