"""CPU: oracle/pem_oracle.py against the golden vectors produced by the reference's own modules
(tools/make_golden.py, run where /root/reference exists)."""
import os

import pytest
import torch

from oracle import pem_oracle as po


def _run(gold, inputs):
    sd = po.make_state_dict(seed=gold["meta"]["seed"])
    rand = gold["rand"]
    if rand is None:
        torch.manual_seed(1)
        rand = torch.rand(gold["meta"]["B"], po.N_PROPOSAL1 * 3)
    return po.pem_forward(sd, inputs["pts"], inputs["dense_fm"], inputs["dense_po"], inputs["dense_fo"],
                          inputs["model"], rand=rand, coarse_npoint=gold["meta"]["coarse_npoint"],
                          return_stages=True, completion="both")


def _compare(gold, out):
    assert torch.equal(out["fps_idx_m"], gold["fps_idx_m"])
    assert torch.equal(out["fps_idx_o"], gold["fps_idx_o"])
    pick = gold["geo_pick"]
    torch.testing.assert_close(out["geo_m"][:, pick[:, 0], pick[:, 1], :], gold["geo_m_pick"], atol=2e-5, rtol=1e-5)
    torch.testing.assert_close(out["geo_o"][:, pick[:, 0], pick[:, 1], :], gold["geo_o_pick"], atol=2e-5, rtol=1e-5)
    for k in ("init_R", "init_t", "pred_R", "pred_t"):
        torch.testing.assert_close(out[k], gold[k], atol=1e-4, rtol=0)
    torch.testing.assert_close(out["pred_pose_score"], gold["pred_pose_score"], atol=2e-3, rtol=0)
    # the deterministic-completion variant (the comparator of the GPU tests on rank-deficient winners) is reproducible too
    for k in ("det_init_R", "det_init_t", "det_pred_R", "det_pred_t"):
        torch.testing.assert_close(out[k], gold[k], atol=1e-4, rtol=0)
    # ... and it IS the reference wherever the reference's winner has three distinct correspondences on both sides
    same = (gold["det_init_R"] - gold["init_R"]).abs().amax(dim=(1, 2)) == 0
    torch.testing.assert_close(out["det_pred_R"][same], gold["pred_R"][same], atol=1e-4, rtol=0)


def test_oracle_matches_reference_small(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "pem_small.pt"), weights_only=False)
    out = _run(gold, gold["inputs"])
    _compare(gold, out)
    torch.testing.assert_close(out["atten_coarse"], gold["atten_coarse"], atol=1e-4, rtol=1e-5)


def test_oracle_matches_reference_full(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "pem_full.pt"), weights_only=False)
    m = gold["meta"]
    inputs = po.make_inputs(B=m["B"], n=m["n"], seed=m["seed"])
    out = _run(gold, inputs)
    _compare(gold, out)
    for R in (out["init_R"], out["pred_R"]):
        torch.testing.assert_close(R @ R.transpose(1, 2), torch.eye(3).expand_as(R), atol=1e-5, rtol=0)
        torch.testing.assert_close(torch.det(R), torch.ones(R.shape[0]), atol=1e-5, rtol=0)


def test_state_dict_layout():
    sd = po.make_state_dict()
    assert sd["coarse_point_matching.transformers.0.layers.0.attention.attention.proj_p.weight"].shape == (256, 256)
    assert "coarse_point_matching.transformers.0.layers.1.attention.attention.proj_p.weight" not in sd
    assert sd["fine_point_matching.PE.mlp1.layer0.conv.weight"].shape == (32, 6, 1, 1)
    assert sd["fine_point_matching.transformers.2.dense_layer.attention.attention.scale"].shape == (1, 1, 256)
    n_param = sum(v.numel() for k, v in sd.items() if "running" not in k and "tracked" not in k and "div_term" not in k)
    assert n_param == 131584 + 3491840 + 5161472        # SURVEY.md 3.2 parameter counts


def test_rank1_rotation_rule():
    """the documented deviation: least rotation taking the source direction onto the reference direction; invariant to the
    sign of the singular pair, a proper rotation, identity for H = 0, half turn for opposite directions"""
    g = torch.Generator().manual_seed(0)
    u = torch.nn.functional.normalize(torch.randn(64, 3, generator=g, dtype=torch.float64), dim=1)
    v = torch.nn.functional.normalize(torch.randn(64, 3, generator=g, dtype=torch.float64), dim=1)
    H = 0.37 * u[:, :, None] * v[:, None, :]
    R = po.rank1_rotation(H)
    torch.testing.assert_close((R @ u[:, :, None]).squeeze(2), v, atol=1e-12, rtol=0)
    torch.testing.assert_close(R @ R.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand_as(R), atol=1e-12, rtol=0)
    torch.testing.assert_close(torch.det(R), torch.ones(64, dtype=torch.float64), atol=1e-12, rtol=0)
    # least rotation: the angle equals the angle between u and v
    ang = torch.acos(((R.diagonal(dim1=1, dim2=2).sum(1) - 1) / 2).clamp(-1, 1))
    torch.testing.assert_close(ang, torch.acos((u * v).sum(1).clamp(-1, 1)), atol=1e-7, rtol=0)
    torch.testing.assert_close(po.rank1_rotation(torch.zeros(1, 3, 3, dtype=torch.float64))[0], torch.eye(3, dtype=torch.float64))
    Rf = po.rank1_rotation(-(u[:1, :, None] * u[:1, None, :]))                       # v = -u
    torch.testing.assert_close((Rf @ u[:1, :, None]).squeeze(2), -u[:1], atol=1e-12, rtol=0)
    torch.testing.assert_close(torch.det(Rf), torch.ones(1, dtype=torch.float64), atol=1e-12, rtol=0)


def test_deterministic_completion_only_touches_rank_deficient_hypotheses():
    g = torch.Generator().manual_seed(3)
    B, n, n1 = 1, 50, 400
    pts2 = torch.randn(B, n, 3, generator=g)
    pts1 = pts2 @ po.random_rotation(B, g).transpose(1, 2) + 0.3
    i1 = torch.randint(0, n, (B, n1 * 3), generator=g)
    i2 = torch.randint(0, n, (B, n1 * 3), generator=g)
    r1, r0 = po._triplet_ranks(i1, i2, B, n1)
    p1 = torch.gather(pts1, 1, i1.unsqueeze(2).repeat(1, 1, 3)).reshape(B * n1, 3, 3)
    p2 = torch.gather(pts2, 1, i2.unsqueeze(2).repeat(1, 1, 3)).reshape(B * n1, 3, 3)
    Ra, ta = po.weighted_procrustes(p2, p1, None, weight_thresh=0.5)
    Rb, tb = po.weighted_procrustes(p2, p1, None, weight_thresh=0.5, rank1=r1, rank0=r0)
    keep = ~(r1 | r0)
    assert keep.any() and (~keep).any()
    assert torch.equal(Ra[keep], Rb[keep]) and torch.equal(ta[keep], tb[keep])
    torch.testing.assert_close(Rb @ Rb.transpose(1, 2), torch.eye(3).expand_as(Rb), atol=1e-5, rtol=0)
