"""Quaternion + translation affine transforms (reference ppfleetx/models/protein_folding/quat_affine.py:1-613, the AlphaFold
``quat_affine`` API used for backbone frames and template unit-vector features).

Same tensor layout decision as ``r3.py``: points are ``[..., 3]`` tensors and rotations ``[..., 3, 3]`` tensors (the reference
passes lists of three / nine component arrays), so applying a frame to a point cloud is one batched mat-mul.  Quaternions are
``[..., 4]`` as ``(w, x, y, z)``.
"""
from __future__ import annotations

from typing import Optional

import torch


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """Unit (or un-normalised — the formula is homogeneous of degree 2) quaternion -> rotation matrix ``[..., 3, 3]``."""
    w, x, y, z = q.unbind(-1)
    rows = [
        torch.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], -1),
    ]
    return torch.stack(rows, -2)


def rot_to_quat(rot: torch.Tensor, unstack_inputs: bool = False) -> torch.Tensor:
    """Rotation matrix -> unit quaternion with ``w >= 0``: the eigenvector of the largest eigenvalue of the symmetric 4x4 matrix K(R)
    (robust for every rotation angle, no branch on the trace)."""
    del unstack_inputs          # accepted for call-site compatibility: rotations are always one [..., 3, 3] tensor here
    xx, xy, xz = rot[..., 0, 0], rot[..., 0, 1], rot[..., 0, 2]
    yx, yy, yz = rot[..., 1, 0], rot[..., 1, 1], rot[..., 1, 2]
    zx, zy, zz = rot[..., 2, 0], rot[..., 2, 1], rot[..., 2, 2]
    k = torch.stack([
        torch.stack([xx + yy + zz, zy - yz, xz - zx, yx - xy], -1),
        torch.stack([zy - yz, xx - yy - zz, xy + yx, xz + zx], -1),
        torch.stack([xz - zx, xy + yx, yy - xx - zz, yz + zy], -1),
        torch.stack([yx - xy, xz + zx, yz + zy, zz - xx - yy], -1),
    ], -2) / 3.0
    _, vecs = torch.linalg.eigh(k.float())
    q = vecs[..., -1].to(rot.dtype)
    return q * torch.where(q[..., :1] < 0, -1.0, 1.0).to(q.dtype)


def quat_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw], -1)


def quat_multiply_by_vec(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """q * (0, v) — the first-order quaternion update used by the structure module."""
    return quat_multiply(q, torch.cat([torch.zeros_like(v[..., :1]), v], -1))


def apply_rot_to_vec(rot: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    return (rot @ vec.unsqueeze(-1)).squeeze(-1)


def apply_inverse_rot_to_vec(rot: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    return (rot.transpose(-1, -2) @ vec.unsqueeze(-1)).squeeze(-1)


class QuatAffine:
    """x -> R(q) x + t over broadcast batch axes."""

    def __init__(self, quaternion: Optional[torch.Tensor], translation: torch.Tensor, rotation: Optional[torch.Tensor] = None,
                 normalize: bool = True, unstack_inputs: bool = False):
        del unstack_inputs          # accepted for call-site compatibility: members are always stacked tensors here
        if quaternion is None:
            assert rotation is not None, "give a quaternion or a rotation"
            quaternion = rot_to_quat(rotation)
        elif normalize:
            quaternion = quaternion / torch.linalg.norm(quaternion, dim=-1, keepdim=True)
        if rotation is None:
            rotation = quat_to_rot(quaternion)
        assert quaternion.shape[-1] == 4 and rotation.shape[-2:] == (3, 3) and translation.shape[-1] == 3
        self.quaternion, self.rotation, self.translation = quaternion, rotation, translation

    # -- (de)serialisation as [..., 7] = quaternion + translation
    def to_tensor(self) -> torch.Tensor:
        return torch.cat([self.quaternion, self.translation], -1)

    @classmethod
    def from_tensor(cls, tensor: torch.Tensor, normalize: bool = False) -> "QuatAffine":
        assert tensor.shape[-1] == 7
        return cls(tensor[..., :4], tensor[..., 4:], normalize=normalize)

    def apply_tensor_fn(self, fn) -> "QuatAffine":
        """``fn`` acts on the leading batch axes (index, unsqueeze, stop-gradient …) of every member."""
        return QuatAffine(fn(self.quaternion), fn(self.translation), rotation=fn(self.rotation), normalize=False)

    def apply_rotation_tensor_fn(self, fn) -> "QuatAffine":
        return QuatAffine(fn(self.quaternion), self.translation, rotation=fn(self.rotation), normalize=False)

    def scale_translation(self, position_scale: float) -> "QuatAffine":
        return QuatAffine(self.quaternion, self.translation * position_scale, rotation=self.rotation, normalize=False)

    def stop_rot_gradient(self) -> "QuatAffine":
        return self.apply_rotation_tensor_fn(lambda t: t.detach())

    # -- composition
    def pre_compose(self, update: torch.Tensor) -> "QuatAffine":
        """New affine = self ∘ small update, the update given as ``[..., 6]`` = imaginary quaternion part (b, c, d) + translation expressed
        in the *local* frame (Jumper et al. Suppl. Alg. 23 "BackboneUpdate")."""
        assert update.shape[-1] == 6
        vec, trans = update[..., :3], update[..., 3:]
        new_q = self.quaternion + quat_multiply_by_vec(self.quaternion, vec)
        new_t = self.translation + apply_rot_to_vec(self.rotation, trans)
        return QuatAffine(new_q, new_t)

    def _expand(self, t: torch.Tensor, extra_dims: int, tail: int) -> torch.Tensor:
        for _ in range(extra_dims):
            t = t.unsqueeze(-1 - tail)
        return t

    def apply_to_point(self, point: torch.Tensor, extra_dims: int = 0) -> torch.Tensor:
        """``point`` has ``extra_dims`` more batch axes (inserted just before the coordinate axis) than the affine."""
        rot, trans = self._expand(self.rotation, extra_dims, 2), self._expand(self.translation, extra_dims, 1)
        return apply_rot_to_vec(rot, point) + trans

    def invert_point(self, transformed_point: torch.Tensor, extra_dims: int = 0) -> torch.Tensor:
        rot, trans = self._expand(self.rotation, extra_dims, 2), self._expand(self.translation, extra_dims, 1)
        return apply_inverse_rot_to_vec(rot, transformed_point - trans)

    def __repr__(self) -> str:
        return f"QuatAffine(batch={tuple(self.translation.shape[:-1])})"


def make_canonical_transform(n_xyz: torch.Tensor, ca_xyz: torch.Tensor, c_xyz: torch.Tensor):
    """(translation, rotation) that moves CA to the origin, C onto the +x axis and N into the xy-plane.

    Two Givens rotations align C with +x (about z, then about y), a third about x brings N into the plane."""
    assert n_xyz.shape[-1] == ca_xyz.shape[-1] == c_xyz.shape[-1] == 3
    translation = -ca_xyz
    n, c = n_xyz + translation, c_xyz + translation
    cx, cy, cz = c.unbind(-1)
    norm_xy = torch.sqrt(1e-20 + cx * cx + cy * cy)
    s1, c1 = -cy / norm_xy, cx / norm_xy
    zeros, ones = torch.zeros_like(s1), torch.ones_like(s1)
    r1 = torch.stack([torch.stack([c1, -s1, zeros], -1), torch.stack([s1, c1, zeros], -1), torch.stack([zeros, zeros, ones], -1)], -2)
    norm = torch.sqrt(1e-20 + cx * cx + cy * cy + cz * cz)
    s2, c2 = cz / norm, norm_xy / norm
    r2 = torch.stack([torch.stack([c2, zeros, s2], -1), torch.stack([zeros, ones, zeros], -1), torch.stack([-s2, zeros, c2], -1)], -2)
    rc = r2 @ r1
    n = (rc @ n.unsqueeze(-1)).squeeze(-1)
    _, ny, nz = n.unbind(-1)
    norm_yz = torch.sqrt(1e-20 + ny * ny + nz * nz)
    s3, c3 = -nz / norm_yz, ny / norm_yz
    r3_ = torch.stack([torch.stack([ones, zeros, zeros], -1), torch.stack([zeros, c3, -s3], -1), torch.stack([zeros, s3, c3], -1)], -2)
    return translation, r3_ @ rc


def make_transform_from_reference(n_xyz: torch.Tensor, ca_xyz: torch.Tensor, c_xyz: torch.Tensor):
    """(rotation, translation) that maps the canonical backbone (CA at 0, C on +x, N in the xy-plane) onto the given one — the inverse of
    ``make_canonical_transform``; this is each residue's backbone frame."""
    translation, rot = make_canonical_transform(n_xyz, ca_xyz, c_xyz)
    return rot.transpose(-1, -2), -translation


# ---- NumPy forms used by feature pipelines that run on the host (reference quat_affine.py:162-174, 513-613)
def apply_rot_to_vec_np(rot, vec, unstack: bool = False):
    """``rot`` indexable as ``rot[i][j]`` (arrays or scalars), ``vec`` a list ``[x, y, z]`` (or one ``[..., 3]`` array with ``unstack``);
    returns the rotated coordinates as a list of three arrays."""
    x, y, z = (vec[..., 0], vec[..., 1], vec[..., 2]) if unstack else vec
    return [rot[i][0] * x + rot[i][1] * y + rot[i][2] * z for i in range(3)]


def make_canonical_transform_np(n_xyz, ca_xyz, c_xyz):
    """NumPy ``make_canonical_transform``: ``(translation [b, 3], rotation [b, 3, 3])`` moving CA to the origin, C onto +x, N into the xy-plane."""
    import numpy as np

    assert n_xyz.ndim == 2 and n_xyz.shape[-1] == 3, n_xyz.shape
    assert n_xyz.shape == ca_xyz.shape == c_xyz.shape, (n_xyz.shape, ca_xyz.shape, c_xyz.shape)
    t, r = make_canonical_transform(*(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)) for a in (n_xyz, ca_xyz, c_xyz)))
    return t.numpy().astype(n_xyz.dtype, copy=False), r.numpy().astype(n_xyz.dtype, copy=False)


def make_transform_from_reference_np(n_xyz, ca_xyz, c_xyz):
    """NumPy ``make_transform_from_reference``: ``(rotation, translation)`` of each residue's backbone frame (rotation first, then translation)."""
    import numpy as np

    translation, rotation = make_canonical_transform_np(n_xyz, ca_xyz, c_xyz)
    return np.transpose(rotation, (0, 2, 1)), -translation
