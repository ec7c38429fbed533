"""tools/make_golden_ism.py -- DEV CONTAINER ONLY (needs /root/reference).

Pins oracle/ism_oracle.py (and, through the GPU tests, csrc/ism.cu) against the reference's OWN template-scoring code:
    PairwiseSimilarity.forward                                    ISM/model/loss.py:21-44
    Instance_Segmentation_Model.compute_semantic_score            ISM/model/detector.py:260-296
    Instance_Segmentation_Model.best_template_pose                ISM/model/detector.py:198-207
imported unmodified from /root/reference (tools/ref_ism_import.py stubs the absent third-party imports) and called on a
bare object carrying `matching_config` and `ref_data` -- the only attributes those methods read.
Writes tests/golden/ism_scoring.pt: BASELINE config #5 (P=200, O=21, T=42) and config #3 (P=64, O=8, T=42), C=1024.

Usage: python tools/make_golden_ism.py"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import ism_oracle as io  # noqa: E402
from ref_ism_import import import_reference_ism, STUBBED  # noqa: E402


def main():
    loss, detector = import_reference_ism()
    print("reference ISM modules imported; stubbed third-party imports:", STUBBED)
    ISMModel = detector.Instance_Segmentation_Model
    cases = {}
    for tag, (P, O, T, seed) in dict(config5_ycbv=(200, 21, 42, 5), config3_ism=(64, 8, 42, 1)).items():
        q, ref = io.make_descriptors(P=P, O=O, T=T, C=1024, seed=seed)
        host = types.SimpleNamespace(
            matching_config=types.SimpleNamespace(metric=loss.PairwiseSimilarity(), aggregation_function="avg_5",
                                                  confidence_thresh=0.2),
            ref_data={"descriptors": ref})
        host.best_template_pose = types.MethodType(ISMModel.best_template_pose, host)
        with torch.no_grad():
            sim = host.matching_config.metric(q, ref)                                     # (P,O,T)
            idx_sel, pred_obj, sem, best_t = ISMModel.compute_semantic_score(host, q)
        # the restatement must reproduce the reference bit for bit (same torch primitives in the same order)
        o_idx, o_obj, o_sem, o_bt, o_sim, o_per = io.compute_semantic_score(q, ref, 0.2)
        assert torch.equal(o_sim, sim), "PairwiseSimilarity restatement differs"
        assert torch.equal(o_idx, idx_sel) and torch.equal(o_obj, pred_obj) and torch.equal(o_bt, best_t)
        assert torch.equal(o_sem, sem)
        print(f"  {tag}: P={P} O={O} T={T}: {len(idx_sel)} proposals above 0.2; oracle == reference bit for bit "
              f"(sim, idx_selected, pred_idx_objects, semantic_score, best_template)")
        cases[tag] = dict(P=P, O=O, T=T, C=1024, seed=seed, sim=sim.clone(), idx_selected=idx_sel, pred_idx_objects=pred_obj,
                          semantic_score=sem, best_template=best_t,
                          input_checksum=dict(q=q.double().sum().item(), ref=ref.double().sum().item()))
    out = os.path.join(ROOT, "tests", "golden", "ism_scoring.pt")
    torch.save(dict(meta=dict(source="ISM/model/loss.py PairwiseSimilarity + ISM/model/detector.py compute_semantic_score / "
                              "best_template_pose imported from /root/reference (CPU, fp32)", torch=torch.__version__,
                              stubbed_imports=list(STUBBED)), cases=cases), out)
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
