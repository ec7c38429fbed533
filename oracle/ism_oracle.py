"""oracle/ism_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the ISM template-scoring path.  The reference modules cannot be imported anywhere in this
environment (ISM/model/loss.py pulls ruamel.yaml through utils/inout.py; ISM/model/detector.py needs pytorch_lightning,
hydra), so this file restates them op for op:
    pairwise_similarity      ISM/model/loss.py:27-44   (PairwiseSimilarity.forward, incl. the repeat/normalize/cosine order)
    compute_semantic_score   ISM/model/detector.py:260-296 with aggregation 'avg_5', and best_template_pose :198-207
Parity status: UNPINNED by reference outputs (no importable reference, no golden vectors in the repository); pinned only
by construction against torch's own F.normalize / F.cosine_similarity / topk / max primitives the reference calls.
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


from sam6d_b200.synth import make_descriptors  # noqa: E402,F401  (synthetic descriptors: shared with bench.py)
