"""Protein-folding trunk: geometry libraries against independent formulas, feature builders, attention variants, and DAP / BP layouts
against the unsharded model (2 gloo ranks)."""
import math

import pytest
import torch

from dist_utils import run_distributed


def _dihedral(p0, p1, p2, p3):
    b0, b1, b2 = p0 - p1, p2 - p1, p3 - p2
    b1n = b1 / b1.norm(dim=-1, keepdim=True)
    v = b0 - (b0 * b1n).sum(-1, keepdim=True) * b1n
    w = b2 - (b2 * b1n).sum(-1, keepdim=True) * b1n
    return torch.atan2((torch.linalg.cross(b1n, v) * w).sum(-1), (v * w).sum(-1))


def test_quaternion_rotation_roundtrip_and_composition():
    from paddlefleetx_b200.models.protein_folding import quat_affine as qa

    torch.manual_seed(0)
    q = torch.randn(64, 4, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.where(q[:, :1] < 0, -1.0, 1.0)
    rot = qa.quat_to_rot(q)
    assert torch.allclose(rot @ rot.transpose(-1, -2), torch.eye(3, dtype=torch.float64).expand(64, 3, 3), atol=1e-12) and torch.allclose(torch.det(rot), torch.ones(64, dtype=torch.float64))
    assert torch.allclose(qa.rot_to_quat(rot), q, atol=1e-6)
    half_turn = qa.rot_to_quat(torch.diag(torch.tensor([1.0, -1.0, -1.0])))            # trace = -1: the branchy formulas divide by ~0 here
    assert torch.allclose(half_turn.abs(), torch.tensor([0.0, 1.0, 0.0, 0.0]), atol=1e-6)
    a, b = q[:32], q[32:]
    assert torch.allclose(qa.quat_to_rot(qa.quat_multiply(a, b)), qa.quat_to_rot(a) @ qa.quat_to_rot(b), atol=1e-12)
    # affine: apply / invert are inverse maps, with extra point axes; pre_compose with a zero update is the identity
    aff = qa.QuatAffine(q[:5], torch.randn(5, 3, dtype=torch.float64))
    pts = torch.randn(5, 7, 3, dtype=torch.float64)
    assert torch.allclose(aff.invert_point(aff.apply_to_point(pts, extra_dims=1), extra_dims=1), pts, atol=1e-12)
    same = aff.pre_compose(torch.zeros(5, 6, dtype=torch.float64))
    assert torch.allclose(same.to_tensor(), aff.to_tensor(), atol=1e-12)
    upd = torch.cat([torch.zeros(5, 3, dtype=torch.float64), torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64).expand(5, 3)], -1)
    moved = aff.pre_compose(upd)                                                          # translate by the frame's own x axis
    assert torch.allclose(moved.translation - aff.translation, aff.rotation[..., :, 0], atol=1e-12)
    assert torch.allclose(qa.QuatAffine.from_tensor(aff.to_tensor()).rotation, aff.rotation, atol=1e-12)


def test_backbone_frames_and_r3_rigids():
    from paddlefleetx_b200.models.protein_folding import quat_affine as qa
    from paddlefleetx_b200.models.protein_folding import r3

    torch.manual_seed(1)
    n, ca, c = torch.randn(3, 10, 3, dtype=torch.float64).unbind(0)
    t, rot = qa.make_canonical_transform(n, ca, c)
    c_can, n_can = (rot @ (c + t).unsqueeze(-1)).squeeze(-1), (rot @ (n + t).unsqueeze(-1)).squeeze(-1)
    assert torch.allclose(c_can[:, 1:], torch.zeros(10, 2, dtype=torch.float64), atol=1e-9) and (c_can[:, 0] > 0).all()
    assert torch.allclose(n_can[:, 2], torch.zeros(10, dtype=torch.float64), atol=1e-9) and (n_can[:, 1] > 0).all()
    frame_rot, frame_t = qa.make_transform_from_reference(n, ca, c)
    assert torch.allclose((frame_rot @ c_can.unsqueeze(-1)).squeeze(-1) + frame_t, c, atol=1e-9)
    # r3: same frame from three points (x axis towards C means: N-side point on the negative axis is the mirrored C)
    rig = r3.rigids_from_3_points(point_on_neg_x_axis=2 * ca - c, origin=ca, point_on_xy_plane=n)
    assert torch.allclose(rig.rot, frame_rot, atol=1e-7) and torch.allclose(rig.trans, ca)
    inv = r3.invert_rigids(rig)
    ident = r3.rigids_mul_rigids(rig, inv)
    assert torch.allclose(ident.rot, torch.eye(3, dtype=torch.float64).expand(10, 3, 3), atol=1e-6) and torch.allclose(ident.trans, torch.zeros(10, 3, dtype=torch.float64), atol=1e-6)
    v = torch.randn(10, 3, dtype=torch.float64)
    assert torch.allclose(r3.rigids_mul_vecs(inv, r3.rigids_mul_vecs(rig, v)), v, atol=1e-6)
    flat = r3.rigids_to_tensor_flat12(rig)
    back = r3.rigids_from_tensor_flat12(flat)
    assert torch.allclose(back.rot, rig.rot) and torch.allclose(r3.rigids_from_tensor4x4(r3.rigids_to_tensor4x4(rig)).trans, rig.trans)
    assert torch.allclose(r3.rigids_from_list(r3.rigids_to_list(rig)).rot, rig.rot)
    assert torch.allclose(r3.vecs_cross_vecs(rig.rot[..., :, 0], rig.rot[..., :, 1]), rig.rot[..., :, 2], atol=1e-6)
    assert torch.allclose(r3.rigids_to_quataffine(rig).rotation, rig.rot) and r3.rigids_from_quataffine(r3.rigids_to_quataffine(rig)).trans.shape == (10, 3)


def test_residue_constants_are_consistent():
    from paddlefleetx_b200.models.protein_folding import residue_constants as rc

    assert len(rc.restypes) == 20 and len(set(rc.restypes)) == 20 and len(rc.atom_types) == 37 and rc.atom_order["OXT"] == 36
    heavy = {"ALA": 5, "ARG": 11, "ASN": 8, "ASP": 8, "CYS": 6, "GLN": 9, "GLU": 9, "GLY": 4, "HIS": 10, "ILE": 8, "LEU": 8, "LYS": 9, "MET": 8,
             "PHE": 11, "PRO": 7, "SER": 6, "THR": 7, "TRP": 14, "TYR": 12, "VAL": 7}
    for name, n in heavy.items():
        atoms = [a for a in rc.restype_name_to_atom14_names[name] if a]
        assert len(atoms) == n == len(rc.residue_atoms[name]) and atoms[:4] == ["N", "CA", "C", "O"] and all(a in rc.atom_order for a in atoms)
        for chi in rc.chi_angles_atoms[name]:
            assert all(a in atoms for a in chi)
        for a, b in rc.residue_atom_renaming_swaps.get(name, {}).items():
            assert a in atoms and b in atoms
    assert sum(map(sum, rc.chi_angles_mask)) == sum(len(v) for v in rc.chi_angles_atoms.values()) and sum(map(sum, rc.chi_pi_periodic)) == 4
    i = rc.restype_order["W"]
    assert rc.restype_atom14_mask[i].sum() == 14 and rc.restype_atom37_mask[i].sum() == 14 and rc.restype_atom14_mask[20].sum() == 0
    for slot in range(14):
        assert rc.restype_atom37_to_atom14[i, rc.restype_atom14_to_atom37[i, slot]] == slot
    onehot = rc.sequence_to_onehot("ACDZ")
    assert onehot.shape == (4, 21) and onehot[3, 20] == 1 and rc.aatype_to_str_sequence(onehot.argmax(-1)) == "ACDX"
    with pytest.raises(ValueError):
        rc.sequence_to_onehot("AZ", map_unknown_to_x=False)


def test_torsion_angles_match_dihedral_formula():
    from paddlefleetx_b200.models.protein_folding import all_atom
    from paddlefleetx_b200.models.protein_folding import residue_constants as rc

    torch.manual_seed(2)
    B, T, N = 1, 2, 6
    aatype = torch.randint(0, 20, (B, T, N))
    aatype[0, 0, 1], aatype[0, 0, 2] = rc.restype_order["D"], rc.restype_order["G"]
    pos, mask = torch.randn(B, T, N, 37, 3, dtype=torch.float64), torch.ones(B, T, N, 37)
    out = all_atom.atom37_to_torsion_angles(aatype, pos, mask)
    sc = out["torsion_angles_sin_cos"]
    defined = out["torsion_angles_mask"] > 0
    assert torch.allclose((sc ** 2).sum(-1)[defined], torch.ones_like(sc[..., 0])[defined], atol=1e-6)      # unit circle wherever the angle exists
    ang = torch.atan2(sc[..., 0], sc[..., 1])
    wrap = lambda d: (d + math.pi) % (2 * math.pi) - math.pi  # noqa: E731
    phi = _dihedral(pos[:, :, :-1, 2], pos[:, :, 1:, 0], pos[:, :, 1:, 1], pos[:, :, 1:, 2])
    omega = _dihedral(pos[:, :, :-1, 1], pos[:, :, :-1, 2], pos[:, :, 1:, 0], pos[:, :, 1:, 1])
    psi_o = _dihedral(pos[..., 0, :], pos[..., 1, :], pos[..., 2, :], pos[..., 4, :])
    assert wrap(ang[:, :, 1:, 1] - phi).abs().max() < 1e-5 and wrap(ang[:, :, 1:, 0] - omega).abs().max() < 1e-5
    assert wrap(ang[..., 2] - (psi_o + math.pi)).abs().max() < 1e-5                     # O sits opposite the next residue's N
    chi_idx = all_atom.get_chi_atom_indices()
    for t in range(T):
        for r in range(N):
            for k, atoms in enumerate(rc.chi_angles_atoms[rc.restype_1to3[rc.restypes[int(aatype[0, t, r])]]]):
                ref = _dihedral(*(pos[0, t, r, i] for i in chi_idx[int(aatype[0, t, r])][k]))
                assert abs(float(wrap(ang[0, t, r, 3 + k] - ref))) < 1e-5
    m = out["torsion_angles_mask"]
    assert m[0, 0, 0, :2].sum() == 0 and m[0, 0, 2, 3:].sum() == 0 and m[0, 0, 1, 3:].tolist() == [1, 1, 0, 0]      # first residue, glycine, aspartate
    alt = out["alt_torsion_angles_sin_cos"]
    assert torch.allclose(alt[0, 0, 1, 4], -sc[0, 0, 1, 4]) and torch.allclose(alt[0, 0, 1, 3], sc[0, 0, 1, 3])       # ASP chi2 is pi-periodic
    mask[0, 0, 3, rc.atom_order["CB"]] = 0
    ph = all_atom.atom37_to_torsion_angles(aatype, pos, mask, placeholder_for_undefined=True)
    if rc.chi_angles_atoms[rc.restype_1to3[rc.restypes[int(aatype[0, 0, 3])]]]:
        assert ph["torsion_angles_mask"][0, 0, 3, 3] == 0 and ph["torsion_angles_sin_cos"][0, 0, 3, 3].tolist() == [1.0, 0.0]


def test_common_helpers():
    from paddlefleetx_b200.models.protein_folding import common as cm

    torch.manual_seed(3)
    x = torch.randn(2, 5, 3)
    d = cm.dgram_from_positions(x, num_bins=15, min_bin=3.25, max_bin=20.75)
    assert d.shape == (2, 5, 5, 15) and d.sum(-1).max() <= 1 and d[:, range(5), range(5)].sum() == 0        # zero distance falls below the first bin
    far = cm.dgram_from_positions(torch.tensor([[0.0, 0, 0], [100.0, 0, 0]]), 15, 3.25, 20.75)
    assert far[0, 1, -1] == 1
    params, idx = torch.randn(2, 6, 4), torch.tensor([[0, 5, 2], [1, 1, 3]])
    got = cm.batched_gather(params, idx, axis=1, batch_dims=1)
    assert got.shape == (2, 3, 4) and torch.equal(got[1, 2], params[1, 3]) and torch.equal(cm.batched_gather(params, torch.tensor([4, 0]), axis=1)[:, 0], params[:, 4])
    val, mask = torch.randn(2, 4, 3), (torch.rand(2, 4, 1) > 0.3).float()
    ref = (val * mask).sum(1) / (mask.sum(1) + 1e-10)
    assert torch.allclose(cm.mask_mean(mask, val, axis=1), ref, atol=1e-6)
    f = lambda a, b: a @ b.transpose(-1, -2)  # noqa: E731
    a, b = torch.randn(7, 4), torch.randn(5, 4)
    assert torch.allclose(cm.subbatch(f, [0], [0], 3, 0)(a, b), f(a, b))
    drop = cm.Dropout(0.5, axis=1).train()
    y = drop(torch.ones(4, 6, 8))
    assert (y[:, :1] == y).all() and set(y.unique().tolist()) <= {0.0, 2.0}
    assert torch.equal(drop.eval()(torch.ones(2, 2)), torch.ones(2, 2))


def test_global_attention_matches_explicit_formula():
    from paddlefleetx_b200.models.protein_folding import GlobalAttention

    torch.manual_seed(4)
    att = GlobalAttention(8, 8, num_head=2, out_dim=8).double()
    torch.nn.init.normal_(att.o.weight, std=0.3)
    q = torch.randn(1, 3, 5, 8, dtype=torch.float64)
    mask = torch.tensor([1, 1, 0, 1, 1.0], dtype=torch.float64).view(1, 1, 5, 1).expand(1, 3, 5, 1)
    out = att(q, q, mask)
    h, d = 2, 4
    for g in range(3):
        qa_ = (q[0, g] * mask[0, g]).sum(0) / mask[0, g].sum()
        qh = (att.q.weight @ qa_).view(h, d) * d ** -0.5
        k, v = q[0, g] @ att.k.weight.T, q[0, g] @ att.v.weight.T
        w = torch.softmax((qh @ k.T).masked_fill(mask[0, g, :, 0] == 0, -1e9), -1)
        avg = w @ v
        gate = torch.sigmoid(q[0, g] @ att.g.weight.T + att.g.bias).view(5, h, d)
        ref = (gate * avg[None]).reshape(5, h * d) @ att.o.weight.T + att.o.bias
        assert torch.allclose(out[0, g], ref, atol=1e-9)
    # masked-out sequences do not influence the others
    q2 = q.clone()
    q2[:, :, 2] += 10
    assert torch.allclose(att(q2, q2, mask)[:, :, [0, 1, 3, 4]], out[:, :, [0, 1, 3, 4]], atol=1e-9)


def test_template_features_and_embedding():
    from paddlefleetx_b200.models.protein_folding import TemplateEmbedding
    from paddlefleetx_b200.models.protein_folding.template import SingleTemplateEmbedding, template_pair_features

    torch.manual_seed(5)
    B, R = 2, 6
    batch = dict(template_aatype=torch.randint(0, 20, (B, R)), template_pseudo_beta=torch.randn(B, R, 3) * 6, template_pseudo_beta_mask=torch.ones(B, R),
                 template_all_atom_positions=torch.randn(B, R, 37, 3) * 3, template_all_atom_masks=torch.ones(B, R, 37))
    feats = template_pair_features(batch, torch.ones(B, R, R), torch.float32, use_template_unit_vector=True)
    assert feats.shape == (B, R, R, SingleTemplateEmbedding.FEATURE_DIM == 88 and 88)
    unit = feats[..., 84:87]
    off = ~torch.eye(R, dtype=torch.bool)
    assert torch.allclose(unit.norm(dim=-1)[:, off], torch.ones(B, R * R - R), atol=1e-3) and unit[:, ~off].abs().max() < 1e-2
    # rotating + translating the whole template leaves every feature unchanged (frames are local)
    from paddlefleetx_b200.models.protein_folding import quat_affine as qa

    rot = qa.quat_to_rot(torch.nn.functional.normalize(torch.randn(4), dim=0))
    moved = dict(batch, template_pseudo_beta=batch["template_pseudo_beta"] @ rot.T + 3.0, template_all_atom_positions=batch["template_all_atom_positions"] @ rot.T + 3.0)
    assert torch.allclose(template_pair_features(moved, torch.ones(B, R, R), torch.float32, use_template_unit_vector=True), feats, atol=2e-4)
    assert template_pair_features(batch, torch.ones(B, R, R), torch.float32)[..., 84:87].abs().sum() == 0           # unit vectors off by default
    emb = TemplateEmbedding(c_z=8, c_t=8, num_block=1, num_head=2, attn_key_dim=8, use_template_unit_vector=True)
    torch.nn.init.normal_(emb.attention.o.weight, std=0.2)
    tb = {k: v[None] for k, v in batch.items()}
    query = torch.randn(1, R, R, 8)
    both = emb(query, dict(tb, template_mask=torch.ones(1, B)), torch.ones(1, R, R))
    assert both.shape == (1, R, R, 8) and both.abs().sum() > 0
    none = emb(query, dict(tb, template_mask=torch.zeros(1, B)), torch.ones(1, R, R))
    assert none.abs().sum() == 0                                                                                      # no valid template -> no update
    first = emb(query, dict(tb, template_mask=torch.tensor([[1.0, 0.0]])), torch.ones(1, R, R))
    only = emb(query, dict({k: v[:, :1] for k, v in tb.items()}, template_mask=torch.ones(1, 1)), torch.ones(1, R, R))
    assert torch.allclose(first, only, atol=1e-5)                                                                     # a masked template is ignored


@pytest.mark.parametrize("mode", ["dap", "bp"])
def test_evoformer_with_templates_and_extra_msa_parallel_matches_single(mode):
    run_distributed("dist_fns:evoformer_parallel_matches_single", 2, mode)


def test_dap_split_phase_pairs_match_one_shot_ops():
    run_distributed("dist_fns:dap_split_phase_pairs", 2)


def test_geometry_and_chemistry_helpers_of_the_reference(tmp_path):
    import numpy as np

    from paddlefleetx_b200.models.protein_folding import quat_affine as QA
    from paddlefleetx_b200.models.protein_folding import r3
    from paddlefleetx_b200.models.protein_folding import residue_constants as rc

    torch.manual_seed(0)
    rig = r3.rigids_from_3_points_vecs(torch.randn(5, 3), torch.randn(5, 3), torch.randn(5, 3))
    back = r3.rigids_from_tensor_flat9(r3.rigids_to_tensor_flat9(rig))
    torch.testing.assert_close(back.rot, rig.rot, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(back.trans, rig.trans)
    assert r3.broadcast_shape([4, 1, 3], [5, 1]) == [4, 5, 3] and r3.broadcast_to(torch.zeros(1, 3), [4, 3]).shape == (4, 3)
    assert float(r3.squared_difference(torch.tensor(3.0), torch.tensor(1.0))) == 4.0
    # NumPy forms agree with the tensor forms and put the canonical backbone back where it came from
    n, ca, c = (np.random.RandomState(i).randn(6, 3) for i in range(3))
    rot, trans = QA.make_transform_from_reference_np(n, ca, c)
    rot_t, trans_t = QA.make_transform_from_reference(*(torch.from_numpy(a) for a in (n, ca, c)))
    np.testing.assert_allclose(rot, rot_t.numpy(), atol=1e-12)
    np.testing.assert_allclose(trans, trans_t.numpy(), atol=1e-12)
    t_can, r_can = QA.make_canonical_transform_np(n, ca, c)
    c_can = np.einsum("bij,bj->bi", r_can, c + t_can)
    n_can = np.einsum("bij,bj->bi", r_can, n + t_can)
    assert np.abs(c_can[:, 1:]).max() < 1e-9 and (c_can[:, 0] > 0).all() and np.abs(n_can[:, 2]).max() < 1e-9
    x, y, z = QA.apply_rot_to_vec_np(np.eye(3) * 2.0, np.array([[1.0, 2.0, 3.0]]), unstack=True)
    assert (x, y, z) == (2.0, 4.0, 6.0) or np.allclose([x[0], y[0], z[0]], [2, 4, 6])
    # chi-angle one-hots: ARG chi1 is N-CA-CB-CG, so atom 1 of chi 0 is CA, atom 2 is CB; GLY has none
    arg, gly = rc.restype_order["R"], rc.restype_order["G"]
    assert rc.chi_atom_1_one_hot.shape == (21, 37, 4)
    assert rc.chi_atom_1_one_hot[arg, rc.atom_order["CA"], 0] == 1 and rc.chi_atom_2_one_hot[arg, rc.atom_order["CB"], 0] == 1
    assert rc.chi_atom_1_one_hot[gly].sum() == 0 and rc.chi_angle_atom(3)[arg, rc.atom_order["CZ"], 3] == 1
    # stereo-chemistry table in the upstream text format
    table = tmp_path / "stereo_chemical_props.txt"
    table.write_text("Bond Residue Mean StdDev\nN-CA ALA 1.459 0.020\nCA-C ALA 1.525 0.026\nCA-CB ALA 1.520 0.021\n-\n\n"
                     "Angle Residue Mean StdDev\nN-CA-C ALA 111.0 2.7\nN-CA-CB ALA 110.1 1.4\n-\n")
    bonds, virtual, angles = rc.load_stereo_chemical_props(str(table))
    assert [b.atom2_name for b in bonds["ALA"]] == ["CA", "C", "CB"] and bonds["UNK"] == [] and len(angles["ALA"]) == 2
    nc = virtual["ALA"][0]
    want = np.sqrt(1.459 ** 2 + 1.525 ** 2 - 2 * 1.459 * 1.525 * np.cos(np.deg2rad(111.0)))
    assert (nc.atom1_name, nc.atom2_name) == ("N", "C") and abs(nc.length - want) < 1e-9 and 0 < nc.stddev < 0.1
    bounds = rc.make_atom14_dists_bounds(props=(bonds, virtual, angles))
    ala = rc.restype_order["A"]
    assert bounds["lower_bound"].shape == (21, 14, 14)
    assert abs(bounds["lower_bound"][ala, 0, 1] - (1.459 - 15 * 0.020)) < 1e-6 and abs(bounds["upper_bound"][ala, 1, 0] - (1.459 + 15 * 0.020)) < 1e-6
    assert abs(bounds["lower_bound"][ala, 0, 3] - (1.55 + 1.52 - 1.5)) < 1e-6 and bounds["upper_bound"][ala, 0, 3] == np.float32(1e10)    # N..O: clash bound only
    assert bounds["lower_bound"][ala, 5, 6] == 0                                                      # empty slots
    with pytest.raises(FileNotFoundError):
        rc.load_stereo_chemical_props(str(tmp_path / "missing.txt"))
