"""Rigid-body geometry in R^3 for structure features (reference ppfleetx/models/protein_folding/r3.py:1-518, the AlphaFold ``r3``
API: vectors, rotations and rigid transforms with broadcasting batch axes).

Layout differs from the reference on purpose: the reference keeps nine separate ``xx … zz`` arrays per rotation and three per
vector (a TPU-era struct-of-arrays choice).  Here a vector is one ``[..., 3]`` tensor, a rotation one ``[..., 3, 3]`` tensor and
a rigid a ``Rigids(rot, trans)`` pair, so compositions are batched mat-muls on contiguous memory and every function is a handful
of fused kernels instead of ~30 element-wise launches.  Function names follow the reference so call sites read the same.
"""
from __future__ import annotations

from typing import NamedTuple

import torch

Vecs = torch.Tensor      # [..., 3]
Rots = torch.Tensor      # [..., 3, 3]  (columns are the images of e_x, e_y, e_z)


class Rigids(NamedTuple):
    rot: torch.Tensor    # [..., 3, 3]
    trans: torch.Tensor  # [..., 3]

    @property
    def shape(self):
        return self.trans.shape[:-1]

    def apply(self, fn) -> "Rigids":
        """Apply ``fn`` (an index / unsqueeze / expand on the *leading* batch axes) to both members: ``r.apply(lambda t: t[:, None])``."""
        return Rigids(fn(self.rot), fn(self.trans))


# ------------------------------------------------------------------------------------------------ vectors
def vecs_from_tensor(x: torch.Tensor) -> Vecs:
    assert x.shape[-1] == 3
    return x


def vecs_to_tensor(v: Vecs) -> torch.Tensor:
    return v


def vecs_add(a: Vecs, b: Vecs) -> Vecs:
    return a + b


def vecs_sub(a: Vecs, b: Vecs) -> Vecs:
    return a - b


def vecs_scale(v: Vecs, s) -> Vecs:
    return v * (s.unsqueeze(-1) if torch.is_tensor(s) and s.dim() == v.dim() - 1 else s)


def vecs_dot_vecs(a: Vecs, b: Vecs) -> torch.Tensor:
    return (a * b).sum(-1)


def vecs_cross_vecs(a: Vecs, b: Vecs) -> Vecs:
    return torch.linalg.cross(a, b, dim=-1)


def vecs_robust_norm(v: Vecs, epsilon: float = 1e-8) -> torch.Tensor:
    return torch.sqrt((v * v).sum(-1) + epsilon)


def vecs_robust_normalize(v: Vecs, epsilon: float = 1e-8) -> Vecs:
    return v / vecs_robust_norm(v, epsilon).unsqueeze(-1)


def vecs_squared_distance(a: Vecs, b: Vecs) -> torch.Tensor:
    d = a - b
    return (d * d).sum(-1)


# ------------------------------------------------------------------------------------------------ rotations
def rots_from_tensor3x3(m: torch.Tensor) -> Rots:
    assert m.shape[-2:] == (3, 3)
    return m


def rots_to_tensor3x3(r: Rots) -> torch.Tensor:
    return r


def rots_from_two_vecs(e0_unnormalized: Vecs, e1_unnormalized: Vecs) -> Rots:
    """Gram-Schmidt frame: e0 along the first vector, e1 in the plane of both, e2 = e0 x e1 (columns of the result)."""
    e0 = vecs_robust_normalize(e0_unnormalized)
    c = vecs_dot_vecs(e1_unnormalized, e0).unsqueeze(-1)
    e1 = vecs_robust_normalize(e1_unnormalized - c * e0)
    e2 = vecs_cross_vecs(e0, e1)
    return torch.stack([e0, e1, e2], dim=-1)


def invert_rots(r: Rots) -> Rots:
    return r.transpose(-1, -2)


def rots_mul_rots(a: Rots, b: Rots) -> Rots:
    return a @ b


def rots_mul_vecs(r: Rots, v: Vecs) -> Vecs:
    return (r @ v.unsqueeze(-1)).squeeze(-1)


# ------------------------------------------------------------------------------------------------ rigids
def rigids_from_3_points(point_on_neg_x_axis: Vecs, origin: Vecs, point_on_xy_plane: Vecs) -> Rigids:
    """Frame with ``origin`` at 0, the first point on the negative x axis and the third in the xy-plane with y > 0 (Jumper et al.
    Suppl. Alg. 21)."""
    rot = rots_from_two_vecs(origin - point_on_neg_x_axis, point_on_xy_plane - origin)
    return Rigids(rot, origin.expand(rot.shape[:-1]) if origin.shape != rot.shape[:-1] else origin)


rigids_from_3_points_vecs = rigids_from_3_points      # the reference has a Vecs form and a tensor form; one layout serves both here


def squared_difference(x, y):
    return torch.square(x - y)


def broadcast_shape(x_shape, y_shape):
    """Shape two batch shapes broadcast to (NumPy rules; the reference's hand-rolled version, r3.py:409-424, only handles equal ranks)."""
    return list(torch.broadcast_shapes(tuple(x_shape), tuple(y_shape)))


def broadcast_to(x: torch.Tensor, broadcast_shape) -> torch.Tensor:
    return x if list(x.shape) == list(broadcast_shape) else x.expand(*broadcast_shape)


def invert_rigids(r: Rigids) -> Rigids:
    inv = invert_rots(r.rot)
    return Rigids(inv, -rots_mul_vecs(inv, r.trans))


def rigids_mul_vecs(r: Rigids, v: Vecs) -> Vecs:
    return rots_mul_vecs(r.rot, v) + r.trans


def rigids_mul_rots(r: Rigids, m: Rots) -> Rigids:
    return Rigids(rots_mul_rots(r.rot, m), r.trans)


def rigids_mul_rigids(a: Rigids, b: Rigids) -> Rigids:
    return Rigids(rots_mul_rots(a.rot, b.rot), rots_mul_vecs(a.rot, b.trans) + a.trans)


def rigids_from_tensor4x4(m: torch.Tensor) -> Rigids:
    assert m.shape[-2:] == (4, 4)
    return Rigids(m[..., :3, :3], m[..., :3, 3])


def rigids_to_tensor4x4(r: Rigids) -> torch.Tensor:
    top = torch.cat([r.rot, r.trans.unsqueeze(-1)], dim=-1)
    bottom = torch.zeros_like(top[..., :1, :])
    bottom[..., 0, 3] = 1
    return torch.cat([top, bottom], dim=-2)


def rigids_from_tensor_flat9(m: torch.Tensor) -> Rigids:
    """[..., 9] = two un-normalised frame vectors + translation."""
    assert m.shape[-1] == 9
    return Rigids(rots_from_two_vecs(m[..., 0:3], m[..., 3:6]), m[..., 6:9])


def rigids_to_tensor_flat9(r: Rigids) -> torch.Tensor:
    """[..., 9] = first two COLUMNS of the rotation (the images of e_x and e_y) + translation — what ``rigids_from_tensor_flat9`` rebuilds the
    frame from by Gram-Schmidt (reference r3.py:352-357)."""
    return torch.cat([r.rot[..., :, 0], r.rot[..., :, 1], r.trans], dim=-1)


def rigids_from_tensor_flat12(m: torch.Tensor) -> Rigids:
    """[..., 12] = row-major rotation (9) + translation (3)."""
    assert m.shape[-1] == 12
    return Rigids(m[..., :9].reshape(m.shape[:-1] + (3, 3)), m[..., 9:])


def rigids_to_tensor_flat12(r: Rigids) -> torch.Tensor:
    return torch.cat([r.rot.reshape(r.rot.shape[:-2] + (9,)), r.trans], dim=-1)


def rigids_from_list(flat) -> Rigids:
    """12 component arrays (xx, xy, …, zz, x, y, z) -> Rigids (reference r3.py component order)."""
    assert len(flat) == 12
    rot = torch.stack(list(flat[:9]), dim=-1)
    return Rigids(rot.reshape(rot.shape[:-1] + (3, 3)), torch.stack(list(flat[9:]), dim=-1))


def rigids_to_list(r: Rigids):
    return [r.rot[..., i, j] for i in range(3) for j in range(3)] + [r.trans[..., i] for i in range(3)]


def rigids_from_quataffine(a) -> Rigids:
    return Rigids(a.rotation, a.translation)


def rigids_to_quataffine(r: Rigids):
    from .quat_affine import QuatAffine

    return QuatAffine(quaternion=None, translation=r.trans, rotation=r.rot)


def identity_rigids(shape, dtype=torch.float32, device=None) -> Rigids:
    rot = torch.eye(3, dtype=dtype, device=device).expand(tuple(shape) + (3, 3)).clone()
    return Rigids(rot, torch.zeros(tuple(shape) + (3,), dtype=dtype, device=device))
