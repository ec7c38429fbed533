"""Instance Segmentation Model pieces on the hot path: per-proposal template scoring.

Mirrors  PairwiseSimilarity                         ISM/model/loss.py:21-44
         Instance_Segmentation_Model.compute_semantic_score / best_template_pose
                                                    ISM/model/detector.py:198-207, 260-296
with the same call signatures and return values.  One fused sm_100a kernel (csrc/ism.cu) computes the clamped cosine
matrix, the avg-5 aggregation, the object argmax and the best-template argmax; the reference's P-fold replication of the
reference descriptors is never formed.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops


class PairwiseSimilarity(nn.Module):
    """forward(query (P,C), reference (O,T,C)) -> (P,O,T) clamped cosine similarity."""

    def __init__(self, metric="cosine", chunk_size=64):
        super().__init__()
        self.metric = metric
        self.chunk_size = chunk_size

    @torch.no_grad()
    def forward(self, query, reference):
        qn = ops.l2norm_rows(query.float().contiguous())
        rn = ops.l2norm_rows(reference.float().contiguous())
        sim, _, _, _, _ = ops.template_score(qn, rn, want_sim=True)
        return sim


def compute_semantic_score(proposal_descriptors, ref_descriptors, aggregation_function="avg_5", confidence_thresh=0.2):
    """detector.py:260-296 -> (idx_selected_proposals, pred_idx_objects, semantic_score, best_template), all int64/f32
    like the reference.  Only 'avg_5' (ISM/configs/model/ISM_sam.yaml) runs fused."""
    if aggregation_function != "avg_5":
        raise NotImplementedError("SAM-6D's ISM configuration uses aggregation_function='avg_5'")
    qn = ops.l2norm_rows(proposal_descriptors.float().contiguous())
    rn = ops.l2norm_rows(ref_descriptors.float().contiguous())
    _, _, best_obj, best_score, best_tmpl = ops.template_score(qn, rn, want_sim=False)
    keep = best_score > confidence_thresh
    idx_selected = torch.arange(best_score.shape[0], device=best_score.device)[keep]
    return idx_selected, best_obj[keep].long(), best_score[keep], best_tmpl[keep].long()


class SemanticScorer(nn.Module):
    """Holds `ref_data["descriptors"]` and `matching_config` like Instance_Segmentation_Model does, exposing
    compute_semantic_score(proposal_descriptors) with the reference signature."""

    def __init__(self, ref_descriptors, aggregation_function="avg_5", confidence_thresh=0.2):
        super().__init__()
        self.ref_data = {"descriptors": ref_descriptors}
        self.matching_config = SimpleNamespace(metric=PairwiseSimilarity(), aggregation_function=aggregation_function,
                                               confidence_thresh=confidence_thresh)

    def compute_semantic_score(self, proposal_decriptors):
        return compute_semantic_score(proposal_decriptors, self.ref_data["descriptors"],
                                      self.matching_config.aggregation_function, self.matching_config.confidence_thresh)
