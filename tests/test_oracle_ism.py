"""CPU: oracle/ism_oracle.py against tests/golden/ism_scoring.pt -- outputs of the reference's own PairwiseSimilarity /
compute_semantic_score / best_template_pose (tools/make_golden_ism.py)."""
import os

import pytest
import torch

from oracle import ism_oracle as io


@pytest.mark.parametrize("case", ["config5_ycbv", "config3_ism"])
def test_oracle_matches_reference_scoring(golden_dir, case):
    gold = torch.load(os.path.join(golden_dir, "ism_scoring.pt"), weights_only=False)
    c = gold["cases"][case]
    q, r = io.make_descriptors(P=c["P"], O=c["O"], T=c["T"], C=c["C"], seed=c["seed"])
    assert q.double().sum().item() == c["input_checksum"]["q"] and r.double().sum().item() == c["input_checksum"]["ref"]
    idx_sel, pred_obj, sem, best_t, scores, _ = io.compute_semantic_score(q, r, 0.2)
    assert torch.equal(idx_sel, c["idx_selected"])
    assert torch.equal(pred_obj, c["pred_idx_objects"])
    assert torch.equal(best_t, c["best_template"])                 # bit-exact argmax template indices
    torch.testing.assert_close(sem, c["semantic_score"], atol=1e-6, rtol=0)
    torch.testing.assert_close(scores, c["sim"], atol=1e-6, rtol=0)
    assert 0 < len(idx_sel) < c["P"]                               # the threshold splits the synthetic proposals
