"""oracle/ism_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the ISM template-scoring path:
    pairwise_similarity      ISM/model/loss.py:27-44   (PairwiseSimilarity.forward, incl. the repeat/normalize/cosine order)
    compute_semantic_score   ISM/model/detector.py:260-296 with aggregation 'avg_5', and best_template_pose :198-207
    appearance_score         ISM/model/loss.py:52-63   (MaskedPatch_MatrixSimilarity.compute_straight), detector.py:298-309
    visible_ratio            ISM/model/loss.py:65-77   (compute_visible_ratio), detector.py:311-323
Parity status: PINNED.  tools/make_golden_ism.py imports the reference's own loss.py / detector.py from /root/reference
(absent third-party imports stubbed, no reference line changed), runs them on the seeded descriptors of BASELINE configs
#3 and #5 and finds this restatement bit-identical (similarity tensor, selected proposals, object indices, scores, template
indices); the outputs are committed as tests/golden/ism_scoring.pt and checked by tests/test_oracle_ism.py (CPU) and
tests/test_gpu_kernels.py (CUDA kernel, bit-exact indices).
"""
import torch
import torch.nn.functional as F


def pairwise_similarity(query: torch.Tensor, reference: torch.Tensor) -> torch.Tensor:
    n_query = query.shape[0]
    n_obj, n_tmpl = reference.shape[0], reference.shape[1]
    references = reference.clone().unsqueeze(0).repeat(n_query, 1, 1, 1)
    queries = query.clone().unsqueeze(1).repeat(1, n_tmpl, 1)
    queries = F.normalize(queries, dim=-1)
    references = F.normalize(references, dim=-1)
    sims = [F.cosine_similarity(queries, references[:, o], dim=-1) for o in range(n_obj)]
    sim = torch.stack(sims).permute(1, 0, 2)            # (P,O,T)
    return sim.clamp(min=0.0, max=1.0)


def compute_semantic_score(desc: torch.Tensor, ref_desc: torch.Tensor, confidence_thresh: float = 0.2):
    scores = pairwise_similarity(desc, ref_desc)
    k = min(5, scores.shape[-1])
    per_obj = torch.mean(torch.topk(scores, k=k, dim=-1)[0], dim=-1)
    score_per_proposal, assigned = torch.max(per_obj, dim=-1)
    idx_sel = torch.arange(len(score_per_proposal))[score_per_proposal > confidence_thresh]
    pred_obj = assigned[idx_sel]
    sem = score_per_proposal[idx_sel]
    filt = scores[idx_sel, ...]
    _, best_t = torch.max(filt, dim=-1)                  # (P', O)
    best_template = torch.gather(best_t, 1, pred_obj[:, None].repeat(1, best_t.shape[1]))[:, 0]
    return idx_sel, pred_obj, sem, best_template, scores, per_obj


def appearance_score(query_patches: torch.Tensor, ref_patches: torch.Tensor) -> torch.Tensor:
    """compute_straight: query (P,Np,C) masked+normalised patch tokens, ref (P,Np,C) those of the best template -> (P,)"""
    sim = torch.matmul(query_patches, ref_patches.permute(0, 2, 1))
    max_ref = torch.max(sim, dim=-1).values
    factor = torch.count_nonzero(query_patches.sum(dim=-1), dim=-1) + 1e-6
    return (torch.sum(max_ref, dim=-1) / factor).clamp(min=0.0, max=1.0)


def visible_ratio(query_patches: torch.Tensor, ref_patches: torch.Tensor, thred: float = 0.5) -> torch.Tensor:
    """compute_visible_ratio: share of the template's valid patches that some query patch matches above `thred`"""
    sim = torch.matmul(query_patches, ref_patches.permute(0, 2, 1)).max(1)[0]
    valid = torch.count_nonzero(sim, dim=(1,)) + 1e-6
    hit = torch.count_nonzero(sim * (sim > thred), dim=(1,))
    return hit / valid


from sam6d_b200.synth import make_descriptors  # noqa: E402,F401  (synthetic descriptors: shared with bench.py)
