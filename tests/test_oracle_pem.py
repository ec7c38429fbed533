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
                          return_stages=True)


def _compare(gold, out):
    assert torch.equal(out["fps_idx_m"], gold["fps_idx_m"])
    assert torch.equal(out["fps_idx_o"], gold["fps_idx_o"])
    pick = gold["geo_pick"]
    torch.testing.assert_close(out["geo_m"][:, pick[:, 0], pick[:, 1], :], gold["geo_m_pick"], atol=2e-5, rtol=1e-5)
    torch.testing.assert_close(out["geo_o"][:, pick[:, 0], pick[:, 1], :], gold["geo_o_pick"], atol=2e-5, rtol=1e-5)
    for k in ("init_R", "init_t", "pred_R", "pred_t"):
        torch.testing.assert_close(out[k], gold[k], atol=1e-4, rtol=0)
    torch.testing.assert_close(out["pred_pose_score"], gold["pred_pose_score"], atol=2e-3, rtol=0)


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
