"""oracle/pem_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement (torch fp32 on the host, no custom kernels) of SAM-6D's Pose
Estimation Model matching path -- everything `Net.forward` does after the ViT
feature extractor -- written as plain functions over a flat weight dictionary
whose keys are the reference's own `state_dict` names.  The formulation follows
the reference op by op (it materialises the same intermediates, e.g. proj_p on
the full (B,S,S,C) embedding), so it doubles as the "port" CPU baseline.

Paths below are relative to SAM-6D/Pose_Estimation_Model/ in the reference.

Parity status: PINNED against the reference's own Python modules (imported from
/root/reference in the dev container with identical seeded weights and inputs) by
tools/make_golden.py, which also writes tests/golden/*.pt;  tests/test_oracle_pem.py
re-checks the oracle against those vectors wherever it runs.  The reference ships no
golden vectors of its own for this path.
"""
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import pn2

SD = Dict[str, torch.Tensor]

# model/pointnet2 hyper-parameters fixed by the reference code / config/base.yaml:17-54
COARSE_NPOINT = 196
FINE_NPOINT = 2048
NUM_HEADS = 4            # coarse_point_matching.py:31, fine_point_matching.py:29
SIGMA_D = 0.2
SIGMA_A = 15.0
ANGLE_K = 3
TEMP = 0.1
N_PROPOSAL1 = 6000
N_PROPOSAL2 = 300
PE_R1, PE_NS1 = 0.1, 32  # fine_point_matching.py:91
PE_R2, PE_NS2 = 0.2, 64
FOCUS = 3


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _heads(x: torch.Tensor, h: int) -> torch.Tensor:
    b, n, c = x.shape
    return x.view(b, n, h, c // h).permute(0, 2, 1, 3)  # 'b n (h c) -> b h n c'


def _unheads(x: torch.Tensor) -> torch.Tensor:
    b, h, n, c = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, n, h * c)


# --------------------------------------------------------------------------------------
# sampling (utils/model_utils.py:53-66)
# --------------------------------------------------------------------------------------
def sample_pts_feats(pts: torch.Tensor, feats: torch.Tensor, npoint: int):
    idx = pn2.furthest_point_sampling(pts, npoint)
    p = pn2.gather_points(pts.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    f = pn2.gather_points(feats.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    return p, f, idx


# --------------------------------------------------------------------------------------
# pairwise squared distance, expanded form (utils/model_utils.py:84-111)
# --------------------------------------------------------------------------------------
def pairwise_sqdist(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    xy = torch.matmul(x, y.transpose(-1, -2))
    x2 = torch.sum(x ** 2, dim=-1).unsqueeze(-1)
    y2 = torch.sum(y ** 2, dim=-1).unsqueeze(-2)
    return (x2 - 2 * xy + y2).clamp(min=0.0)


# --------------------------------------------------------------------------------------
# geometric structure embedding (model/transformer.py:257-349)
# --------------------------------------------------------------------------------------
def sinusoidal_embedding(idx: torch.Tensor, d_model: int) -> torch.Tensor:
    """Interleaved [sin(w0 x), cos(w0 x), sin(w1 x), ...]; transformer.py:262-283."""
    div = torch.exp(torch.arange(0, d_model, 2, device=idx.device).float() * (-math.log(10000.0) / d_model))
    om = idx.reshape(-1, 1, 1) * div.view(1, -1, 1)
    emb = torch.cat([torch.sin(om), torch.cos(om)], dim=2)
    return emb.view(*idx.shape, d_model)


def geo_embedding_indices(points: torch.Tensor):
    """transformer.py:302-332.  points (B,S,3) -> d_idx (B,S,S), a_idx (B,S,S,k)."""
    b, s, _ = points.shape
    dist = torch.sqrt(pairwise_sqdist(points, points))
    d_idx = dist / SIGMA_D
    knn = dist.topk(k=ANGLE_K + 1, dim=2, largest=False)[1][:, :, 1:]           # (B,S,k)
    knn_pts = torch.gather(points.unsqueeze(1).expand(b, s, s, 3), 2, knn.unsqueeze(3).expand(b, s, ANGLE_K, 3))
    ref = (knn_pts - points.unsqueeze(2)).unsqueeze(2).expand(b, s, s, ANGLE_K, 3)   # knn(i,k) - p_i
    anc = (points.unsqueeze(1) - points.unsqueeze(2)).unsqueeze(3).expand(b, s, s, ANGLE_K, 3)  # p_j - p_i
    sin_v = torch.linalg.norm(torch.cross(ref, anc, dim=-1), dim=-1)
    cos_v = torch.sum(ref * anc, dim=-1)
    a_idx = torch.atan2(sin_v, cos_v) * (180.0 / (SIGMA_A * math.pi))
    return d_idx, a_idx


def geo_embedding(sd: SD, points: torch.Tensor, prefix: str = "geo_embedding") -> torch.Tensor:
    """transformer.py:334-349 with reduction_a = 'max'.  -> (B,S,S,C)."""
    c = sd[prefix + ".proj_d.weight"].shape[0]
    d_idx, a_idx = geo_embedding_indices(points)
    d_emb = _lin(sd, prefix + ".proj_d", sinusoidal_embedding(d_idx, c))
    a_emb = _lin(sd, prefix + ".proj_a", sinusoidal_embedding(a_idx, c)).max(dim=3)[0]
    return d_emb + a_emb


# --------------------------------------------------------------------------------------
# transformer layers (model/transformer.py:93-513)
# --------------------------------------------------------------------------------------
def attention_output(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AttentionOutput, transformer.py:182-197: LN(x + squeeze(relu(expand(x))))."""
    h = _lin(sd, p + ".squeeze", torch.relu(_lin(sd, p + ".expand", x)))
    return _ln(sd, p + ".norm", x + h)


def rpe_self_layer(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """RPETransformerLayer, transformer.py:352-465.  p = '<...>.layers.0'."""
    a = p + ".attention.attention"
    q = _heads(_lin(sd, a + ".proj_q", x), NUM_HEADS)
    k = _heads(_lin(sd, a + ".proj_k", x), NUM_HEADS)
    v = _heads(_lin(sd, a + ".proj_v", x), NUM_HEADS)
    b, s, _, c = emb.shape
    pe = _lin(sd, a + ".proj_p", emb).view(b, s, s, NUM_HEADS, c // NUM_HEADS).permute(0, 3, 1, 2, 4)
    sc_p = torch.einsum("bhnc,bhnmc->bhnm", q, pe)
    sc_e = torch.einsum("bhnc,bhmc->bhnm", q, k)
    att = F.softmax((sc_e + sc_p) / (c // NUM_HEADS) ** 0.5, dim=-1)
    hid = _unheads(torch.matmul(att, v))
    y = _ln(sd, p + ".attention.norm", _lin(sd, p + ".attention.linear", hid) + x)
    return attention_output(sd, p + ".output", y)


def cross_layer(sd: SD, p: str, x: torch.Tensor, mem: torch.Tensor) -> torch.Tensor:
    """TransformerLayer, transformer.py:93-224.  p = '<...>.layers.1'."""
    a = p + ".attention.attention"
    q = _heads(_lin(sd, a + ".proj_q", x), NUM_HEADS)
    k = _heads(_lin(sd, a + ".proj_k", mem), NUM_HEADS)
    v = _heads(_lin(sd, a + ".proj_v", mem), NUM_HEADS)
    d = q.shape[-1]
    att = F.softmax(torch.einsum("bhnc,bhmc->bhnm", q, k) / d ** 0.5, dim=-1)
    hid = _unheads(torch.matmul(att, v))
    y = _ln(sd, p + ".attention.norm", _lin(sd, p + ".attention.linear", hid) + x)
    return attention_output(sd, p + ".output", y)


def geometric_transformer(sd: SD, p: str, f0, e0, f1, e1):
    """GeometricTransformer(['self','cross'], parallel=False), transformer.py:469-513."""
    f0 = rpe_self_layer(sd, p + ".layers.0", f0, e0)
    f1 = rpe_self_layer(sd, p + ".layers.0", f1, e1)
    f0 = cross_layer(sd, p + ".layers.1", f0, f1)
    f1 = cross_layer(sd, p + ".layers.1", f1, f0)   # sees the already-updated f0
    return f0, f1


# --------------------------------------------------------------------------------------
# score matrix + pose solvers (utils/model_utils.py:114-136, 187-383)
# --------------------------------------------------------------------------------------
def feature_similarity(f1: torch.Tensor, f2: torch.Tensor) -> torch.Tensor:
    f1 = F.normalize(f1, p=2, dim=2)
    f2 = F.normalize(f2, p=2, dim=2)
    return (f1 @ f2.transpose(1, 2)) / TEMP


def _any_orth(a: torch.Tensor) -> torch.Tensor:
    """unit vector orthogonal to the unit vectors a (n,3): a x e_k, k = the smallest |component| (first on ties)"""
    ab = a.abs()
    k = torch.where((ab[:, 0] <= ab[:, 1]) & (ab[:, 0] <= ab[:, 2]), 0, torch.where(ab[:, 1] <= ab[:, 2], 1, 2))
    e = torch.nn.functional.one_hot(k, 3).to(a.dtype)
    o = torch.linalg.cross(a, e)
    return o / o.norm(dim=1, keepdim=True)


def rank1_rotation(H: torch.Tensor) -> torch.Tensor:
    """THE documented deviation from model_utils.py:352-358 (sam6d_b200/csrc/svd3.cuh: rank1_rotation).

    H (n,3,3) float64 with one singular value above rounding noise: sigma u1 v1^T (u1 on the source side).  The reference's
    R = V diag(1,1,det) U^T then rotates about the v1 axis by whatever angle LAPACK's noise-level second singular pair
    implies.  The deterministic completion is the least rotation taking u1 to v1:
        R = c I + [w]x + w w^T / (1 + c),  w = u1 x v1,  c = u1 . v1      (half turn about any_orth(u1) when c -> -1)."""
    U, S, Vh = torch.linalg.svd(H)
    u1, v1 = U[:, :, 0], Vh[:, 0, :]
    c = (u1 * v1).sum(1)
    w = torch.linalg.cross(u1, v1)
    eye = torch.eye(3, dtype=H.dtype, device=H.device).expand_as(H)
    K = torch.zeros_like(H)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    R = c[:, None, None] * eye + K + w[:, :, None] * w[:, None, :] / (1.0 + c).clamp_min(1e-300)[:, None, None]
    flip = (1.0 + c) < 1e-9
    if flip.any():
        a = _any_orth(u1[flip])
        R[flip] = 2.0 * a[:, :, None] * a[:, None, :] - torch.eye(3, dtype=H.dtype, device=H.device)
    zero = ~(S[:, 0] > 0)
    R[zero] = torch.eye(3, dtype=H.dtype, device=H.device)
    return R


def weighted_procrustes(src, ref, weights=None, weight_thresh=0.0, eps=1e-5, rank1=None, rank0=None):
    """model_utils.py:287-363: R, t with ref ~= R src + t.
    rank1 / rank0 (bool masks over the batch, default None = the reference's behaviour for every element): elements
    whose rotation the caller wants from the deterministic completion (rank1_rotation) / set to the identity."""
    bsz = src.shape[0]
    if weights is None:
        weights = torch.ones_like(src[:, :, 0])
    weights = torch.where(weights < weight_thresh, torch.zeros_like(weights), weights)
    weights = (weights / (weights.sum(dim=1, keepdim=True) + eps)).unsqueeze(2)
    c_s = (src * weights).sum(dim=1, keepdim=True)
    c_r = (ref * weights).sum(dim=1, keepdim=True)
    H = (src - c_s).permute(0, 2, 1) @ (weights * (ref - c_r))
    U, _, V = torch.svd(H)
    Ut = U.transpose(1, 2)
    eye = torch.eye(3, device=src.device).unsqueeze(0).repeat(bsz, 1, 1)
    eye[:, -1, -1] = torch.sign(torch.det(V @ Ut))
    R = V @ eye @ Ut
    if rank1 is not None and rank1.any():
        H64 = (src[rank1] - c_s[rank1]).double().permute(0, 2, 1) @ (weights[rank1] * (ref[rank1] - c_r[rank1])).double()
        R[rank1] = rank1_rotation(H64).float()
    if rank0 is not None and rank0.any():
        R[rank0] = torch.eye(3, device=src.device)
    t = (c_r.permute(0, 2, 1) - R @ c_s.permute(0, 2, 1)).squeeze(2)
    return R, t


def soft_assignment(atten: torch.Tensor):
    """model_utils.py:206-214 / 262-266: dual softmax, bg-aware labels, masked inner block."""
    score = torch.softmax(atten, dim=2) * torch.softmax(atten, dim=1)
    lab1 = torch.max(score[:, 1:, :], dim=2)[1]
    lab2 = torch.max(score[:, :, 1:], dim=1)[1]
    w1 = (lab1 > 0).float()
    w2 = (lab2 > 0).float()
    inner = score[:, 1:, 1:] * w1.unsqueeze(2) * w2.unsqueeze(1)
    return inner, w1, w2, lab1, lab2


def _triplet_ranks(i1, i2, B, n1):
    """(rank1, rank0) masks (B*n1,) of the hypotheses whose triplet repeats a point of either cloud: their centred points
    are collinear (rank-1 cross-covariance), or a single point on one side (rank 0)."""
    t1, t2 = i1.reshape(B * n1, 3), i2.reshape(B * n1, 3)
    eq = lambda a: (a[:, 0] == a[:, 1]).int() + (a[:, 0] == a[:, 2]).int() + (a[:, 1] == a[:, 2]).int()  # noqa: E731
    e1, e2 = eq(t1), eq(t2)
    rank0 = (e1 == 3) | (e2 == 3)
    return ~rank0 & ((e1 + e2) > 0), rank0


def coarse_Rt(atten, pts1, pts2, model_pts, rand: Optional[torch.Tensor] = None,
              n1: int = N_PROPOSAL1, n2: int = N_PROPOSAL2, return_debug: bool = False, completion: str = "reference"):
    """compute_coarse_Rt, model_utils.py:187-246.  `rand` replaces torch.rand(B, 3*n1).
    completion="reference": every hypothesis through torch.svd like the reference.  completion="deterministic": the
    rank-deficient hypotheses (triplets that repeat a point) get the rotation the CUDA path gives them (rank1_rotation /
    identity) -- the one place where the reference's output is not a function of its inputs."""
    assert completion in ("reference", "deterministic")
    B, N1, _ = pts1.shape
    N2 = pts2.shape[1]
    inner, w1, _, _, _ = soft_assignment(atten)
    score = inner.contiguous().reshape(B, N1 * N2) ** 1.5
    cdf = torch.cumsum(score, dim=1)
    cdf = cdf / (cdf[:, -1].unsqueeze(1).contiguous() + 1e-8)
    if rand is None:
        rand = torch.rand(B, n1 * 3)
    idx = torch.searchsorted(cdf, rand)
    i1 = idx.div(N2, rounding_mode="floor").clamp(max=N1 - 1)
    i2 = (idx % N2).clamp(max=N2 - 1)
    p1 = torch.gather(pts1, 1, i1.unsqueeze(2).repeat(1, 1, 3)).reshape(B * n1, 3, 3)
    p2 = torch.gather(pts2, 1, i2.unsqueeze(2).repeat(1, 1, 3)).reshape(B * n1, 3, 3)
    if completion == "deterministic":
        r1, r0 = _triplet_ranks(i1, i2, B, n1)
        Rs, ts = weighted_procrustes(p2, p1, None, weight_thresh=0.5, rank1=r1, rank0=r0)
    else:
        Rs, ts = weighted_procrustes(p2, p1, None, weight_thresh=0.5)
    Rs = Rs.reshape(B, n1, 3, 3)
    ts = ts.reshape(B, n1, 1, 3)
    p1 = p1.reshape(B, n1, 3, 3)
    p2 = p2.reshape(B, n1, 3, 3)
    resid = torch.norm((p1 - ts) @ Rs - p2, dim=3).mean(2)
    top = torch.topk(resid, n2, dim=1, largest=False)[1]
    Rsel = torch.gather(Rs, 1, top.reshape(B, n2, 1, 1).repeat(1, 1, 3, 3))
    tsel = torch.gather(ts, 1, top.reshape(B, n2, 1, 1).repeat(1, 1, 1, 3))
    tp = ((pts1.unsqueeze(1) - tsel) @ Rsel).reshape(B * n2, -1, 3)
    mp = model_pts.unsqueeze(1).repeat(1, n2, 1, 1).reshape(B * n2, -1, 3)
    dis = torch.sqrt(pairwise_sqdist(tp, mp)).min(2)[0].reshape(B, n2, -1)
    scores = w1.unsqueeze(1).sum(2) / ((dis * w1.unsqueeze(1)).sum(2) + 1e-8)
    best = scores.max(1)[1]
    R = torch.gather(Rsel, 1, best.reshape(B, 1, 1, 1).repeat(1, 1, 3, 3)).squeeze(1)
    t = torch.gather(tsel, 1, best.reshape(B, 1, 1, 1).repeat(1, 1, 1, 3)).squeeze(2).squeeze(1)
    if return_debug:
        # a hypothesis is rank-deficient when a correspondence repeats inside its triplet: H then has one singular value
        # above fp32 noise and the reference's rotation is decided by LAPACK on rounding noise ("Parity")
        tri1, tri2 = i1.reshape(B, n1, 3), i2.reshape(B, n1, 3)
        rep = lambda a: (a[..., 0] == a[..., 1]) | (a[..., 0] == a[..., 2]) | (a[..., 1] == a[..., 2])  # noqa: E731
        degenerate = rep(tri1) | rep(tri2)
        win = torch.gather(top, 1, best.unsqueeze(1)).squeeze(1)
        return R, t, dict(score=score, cdf=cdf, idx=idx, w1=w1, Rs=Rs, ts=ts.squeeze(2), resid=resid, top=top,
                          sel_scores=scores, best=best, degenerate=degenerate,
                          winner_degenerate=torch.gather(degenerate, 1, win.unsqueeze(1)).squeeze(1),
                          best_score=scores.max(1)[0])
    return R, t


def fine_Rt(atten, pts1, pts2, model_pts, dis_thres: float = 0.15, return_debug: bool = False):
    """compute_fine_Rt, model_utils.py:250-283."""
    inner, w1, _, lab1, lab2 = soft_assignment(atten)
    row = inner.sum(2, keepdim=True)
    pred = (inner / (row + 1e-6)) @ pts2
    wts = inner.sum(2)
    R, t = weighted_procrustes(pred, pts1, wts, weight_thresh=0.0)
    tp = (pts1 - t.unsqueeze(1)) @ R
    dis = torch.sqrt(pairwise_sqdist(tp, model_pts)).min(2)[0]
    hit = (dis < dis_thres).float()
    score = (hit * w1).sum(1) / (w1.sum(1) + 1e-8)
    score = score * w1.mean(1)
    if return_debug:
        return R, t, score, dict(lab1=lab1, lab2=lab2, wts=wts, pred=pred)
    return R, t, score


# --------------------------------------------------------------------------------------
# coarse point matching (model/coarse_point_matching.py:38-81, inference branch)
# --------------------------------------------------------------------------------------
def coarse_features(sd: SD, f1, geo1, f2, geo2, prefix: str = "coarse_point_matching"):
    B = f1.shape[0]
    bg = sd[prefix + ".bg_token"].repeat(B, 1, 1)
    f1 = torch.cat([bg, _lin(sd, prefix + ".in_proj", f1)], dim=1)
    f2 = torch.cat([bg, _lin(sd, prefix + ".in_proj", f2)], dim=1)
    nblock = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(prefix + ".transformers."))
    for i in range(nblock):
        f1, f2 = geometric_transformer(sd, f"{prefix}.transformers.{i}", f1, geo1, f2, geo2)
    return f1, f2


def coarse_point_matching(sd: SD, p1, f1, geo1, p2, f2, geo2, radius, model,
                          rand: Optional[torch.Tensor] = None, prefix: str = "coarse_point_matching",
                          return_debug: bool = False, completion: str = "reference"):
    f1, f2 = coarse_features(sd, f1, geo1, f2, geo2, prefix)
    atten = feature_similarity(_lin(sd, prefix + ".out_proj", f1), _lin(sd, prefix + ".out_proj", f2))
    init_R, init_t, dbg = coarse_Rt(atten, p1, p2, model / (radius.reshape(-1, 1, 1) + 1e-6), rand, return_debug=True,
                                    completion=completion)
    if return_debug:
        return init_R, init_t, atten, dbg
    return init_R, init_t, atten


# --------------------------------------------------------------------------------------
# positional encoding (model/fine_point_matching.py:90-125, pointnet2_utils.py:317-376,
# pytorch_utils.py:25-206 -- Conv2d(1x1, no bias) -> BatchNorm2d(eval) -> ReLU)
# --------------------------------------------------------------------------------------
def _shared_mlp(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    for j in range(3):
        lp = f"{p}.layer{j}"
        x = F.conv2d(x, sd[lp + ".conv.weight"])
        x = F.batch_norm(x, sd[lp + ".normlayer.bn.running_mean"], sd[lp + ".normlayer.bn.running_var"],
                         sd[lp + ".normlayer.bn.weight"], sd[lp + ".normlayer.bn.bias"], False, 0.0, 1e-5)
        x = torch.relu(x)
    return x


def _query_and_group(pts: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    """QueryAndGroup(use_xyz=True)(xyz=pts, new_xyz=pts, features=pts^T) -> (B,6,N,ns)."""
    idx = pn2.ball_query(pts, pts, radius, nsample)
    xyz_t = pts.transpose(1, 2).contiguous()
    grouped = pn2.group_points(xyz_t, idx)
    rel = grouped - pts.transpose(1, 2).unsqueeze(-1)
    return torch.cat([rel, grouped], dim=1)


def positional_encoding(sd: SD, pts: torch.Tensor, prefix: str = "fine_point_matching.PE") -> torch.Tensor:
    pts = pts.contiguous()
    f1 = _shared_mlp(sd, prefix + ".mlp1", _query_and_group(pts, PE_R1, PE_NS1)).max(dim=3)[0]
    f2 = _shared_mlp(sd, prefix + ".mlp2", _query_and_group(pts, PE_R2, PE_NS2)).max(dim=3)[0]
    feat = torch.cat([f1, f2], dim=1)                                   # (B,256,N)
    feat = F.conv1d(feat, sd[prefix + ".mlp3.conv.weight"], sd[prefix + ".mlp3.conv.bias"])
    return feat.transpose(1, 2)


# --------------------------------------------------------------------------------------
# sparse-to-dense transformer (model/transformer.py:518-673)
# --------------------------------------------------------------------------------------
def linear_attention(sd: SD, p: str, xq: torch.Tensor, xkv: torch.Tensor) -> torch.Tensor:
    """LinearAttention, transformer.py:534-564 (focused linear attention)."""
    q = torch.relu(_lin(sd, p + ".proj_q", xq)) + 1e-6
    k = torch.relu(_lin(sd, p + ".proj_k", xkv)) + 1e-6
    v = _lin(sd, p + ".proj_v", xkv)
    scale = F.softplus(sd[p + ".scale"])
    q = q / scale
    k = k / scale
    qn = q.norm(dim=-1, keepdim=True)
    kn = k.norm(dim=-1, keepdim=True)
    q = q ** FOCUS
    k = k ** FOCUS
    q = (q / q.norm(dim=-1, keepdim=True)) * qn
    k = (k / k.norm(dim=-1, keepdim=True)) * kn
    b, n, c = q.shape
    h = NUM_HEADS
    q, k, v = (t.view(b, -1, h, c // h).permute(0, 2, 1, 3).reshape(b * h, -1, c // h) for t in (q, k, v))
    i, j, cc, d = q.shape[-2], k.shape[-2], k.shape[-1], v.shape[-1]
    z = 1 / (torch.einsum("bic,bc->bi", q, k.sum(dim=1)) + 1e-6)
    if i * j * (cc + d) > cc * d * (i + j):
        kv = torch.einsum("bjc,bjd->bcd", k, v)
        x = torch.einsum("bic,bcd,bi->bid", q, kv, z)
    else:
        qk = torch.einsum("bic,bjc->bij", q, k)
        x = torch.einsum("bij,bjd,bi->bid", qk, v, z)
    return x.view(b, h, n, c // h).permute(0, 2, 1, 3).reshape(b, n, c)


def linear_transformer_layer(sd: SD, p: str, x: torch.Tensor, mem: torch.Tensor) -> torch.Tensor:
    """LinearTransformerLayer, transformer.py:567-608."""
    hid = linear_attention(sd, p + ".attention.attention", x, mem)
    y = _ln(sd, p + ".attention.norm", _lin(sd, p + ".attention.linear", hid) + x)
    return attention_output(sd, p + ".output", y)


def _sample_feats(dense: torch.Tensor, fps_idx: torch.Tensor) -> torch.Tensor:
    """SparseToDenseTransformer._sample_feats, transformer.py:651-658.  NOTE (quirk Q1): the
    gather runs on the token sequence that already includes the bg token at row 0, with an
    index that addresses the N dense points, so sparse token j is dense row fps_idx[j] of the
    (N+1)-long sequence."""
    bg = dense[:, 0:1, :]
    g = pn2.gather_points(dense.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
    return torch.cat([bg, g], dim=1)


def sparse_to_dense(sd: SD, p: str, d0, e0, idx0, d1, e1, idx1):
    """SparseToDenseTransformer.forward, transformer.py:642-673 (with/replace bg token)."""
    s0 = _sample_feats(d0, idx0)
    s1 = _sample_feats(d1, idx1)
    s0, s1 = geometric_transformer(sd, p + ".sparse_layer", s0, e0, s1, e1)
    out = []
    for dense, sparse in ((d0, s0), (d1, s1)):
        y = linear_transformer_layer(sd, p + ".dense_layer", dense[:, 1:, :].contiguous(), sparse[:, 1:, :].contiguous())
        out.append(torch.cat([sparse[:, 0:1, :], y], dim=1))
    return out[0], out[1]


# --------------------------------------------------------------------------------------
# fine point matching (model/fine_point_matching.py:39-86, inference branch)
# --------------------------------------------------------------------------------------
def fine_features(sd: SD, p1, f1, geo1, idx1, p2, f2, geo2, idx2, init_R, init_t,
                  prefix: str = "fine_point_matching"):
    B = p1.shape[0]
    p1_ = (p1 - init_t.unsqueeze(1)) @ init_R
    bg = sd[prefix + ".bg_token"].repeat(B, 1, 1)
    f1 = torch.cat([bg, _lin(sd, prefix + ".in_proj", f1) + positional_encoding(sd, p1_, prefix + ".PE")], dim=1)
    f2 = torch.cat([bg, _lin(sd, prefix + ".in_proj", f2) + positional_encoding(sd, p2, prefix + ".PE")], dim=1)
    nblock = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(prefix + ".transformers."))
    for i in range(nblock):
        f1, f2 = sparse_to_dense(sd, f"{prefix}.transformers.{i}", f1, geo1, idx1, f2, geo2, idx2)
    return f1, f2


def fine_point_matching(sd: SD, p1, f1, geo1, idx1, p2, f2, geo2, idx2, radius, model, init_R, init_t,
                        prefix: str = "fine_point_matching", return_atten: bool = False):
    f1, f2 = fine_features(sd, p1, f1, geo1, idx1, p2, f2, geo2, idx2, init_R, init_t, prefix)
    atten = feature_similarity(_lin(sd, prefix + ".out_proj", f1), _lin(sd, prefix + ".out_proj", f2))
    R, t, score = fine_Rt(atten, p1, p2, model / (radius.reshape(-1, 1, 1) + 1e-6))
    t = t * (radius.reshape(-1, 1) + 1e-6)
    if return_atten:
        return R, t, score, atten
    return R, t, score


# --------------------------------------------------------------------------------------
# Net.forward after the feature extractor (model/pose_estimation_model.py:23-53 and the
# inference branch of ViTEncoder.forward, model/feature_extraction.py:135-142)
# --------------------------------------------------------------------------------------
def pem_forward(sd: SD, pts, dense_fm, dense_po, dense_fo, model,
                rand: Optional[torch.Tensor] = None, coarse_npoint: int = COARSE_NPOINT,
                return_stages: bool = False, completion: str = "reference"):
    """pts (B,N,3) observed cloud, dense_fm (B,N,C) its features, dense_po/dense_fo the template
    bank, model (B,Nm,3) CAD samples -> dict(init_R, init_t, pred_R, pred_t, pred_pose_score).
    completion: see coarse_Rt ("reference" = the reference's behaviour; "deterministic" = the CUDA path's rule for
    rank-deficient pose hypotheses, the single documented deviation)."""
    B = pts.shape[0]
    radius = torch.norm(dense_po, dim=2).max(1)[0]
    dense_pm = pts / (radius.reshape(-1, 1, 1) + 1e-6)
    dense_po = dense_po / (radius.reshape(-1, 1, 1) + 1e-6)
    bg_point = torch.ones(B, 1, 3, device=pts.device) * 100
    sp_m, sf_m, idx_m = sample_pts_feats(dense_pm, dense_fm, coarse_npoint)
    geo_m = geo_embedding(sd, torch.cat([bg_point, sp_m], dim=1))
    sp_o, sf_o, idx_o = sample_pts_feats(dense_po, dense_fo, coarse_npoint)
    geo_o = geo_embedding(sd, torch.cat([bg_point, sp_o], dim=1))
    first = "reference" if completion == "both" else completion
    init_R, init_t, atten_c, cdbg = coarse_point_matching(sd, sp_m, sf_m, geo_m, sp_o, sf_o, geo_o, radius, model, rand,
                                                          return_debug=True, completion=first)
    pred_R, pred_t, score, atten_f = fine_point_matching(
        sd, dense_pm, dense_fm, geo_m, idx_m, dense_po, dense_fo, geo_o, idx_o, radius, model,
        init_R, init_t, return_atten=True)
    out = dict(init_R=init_R, init_t=init_t, pred_R=pred_R, pred_t=pred_t, pred_pose_score=score)
    if completion == "both":
        # the same coarse score matrix through the deterministic completion; the fine stage is re-run only for the
        # proposals whose initial pose changed (proposals are independent)
        dR, dt, ddbg = coarse_Rt(atten_c, sp_m, sp_o, model / (radius.reshape(-1, 1, 1) + 1e-6), rand, return_debug=True,
                                 completion="deterministic")
        ch = ((dR - init_R).abs().amax(dim=(1, 2)) > 0) | ((dt - init_t).abs().amax(dim=1) > 0)
        pR, pt, ps = pred_R.clone(), pred_t.clone(), score.clone()
        if ch.any():
            pR[ch], pt[ch], ps[ch] = fine_point_matching(
                sd, dense_pm[ch], dense_fm[ch], geo_m[ch], idx_m[ch], dense_po[ch], dense_fo[ch], geo_o[ch], idx_o[ch],
                radius[ch], model[ch], dR[ch], dt[ch])
        out.update(det_init_R=dR, det_init_t=dt, det_pred_R=pR, det_pred_t=pt, det_pred_pose_score=ps,
                   det_init_score=ddbg["best_score"], det_init_degenerate=ddbg["winner_degenerate"])
    if return_stages:
        out.update(fps_idx_m=idx_m, fps_idx_o=idx_o, sparse_pm=sp_m, sparse_po=sp_o, geo_m=geo_m, geo_o=geo_o,
                   atten_coarse=atten_c, atten_fine=atten_f, radius=radius,
                   init_degenerate=cdbg["winner_degenerate"], init_score=cdbg["best_score"])
    return out


# --------------------------------------------------------------------------------------
# seeded weights and synthetic proposals: shared with bench.py, so they live in the product package (data, not algorithm)
# --------------------------------------------------------------------------------------
from sam6d_b200.synth import make_pem_state_dict as make_state_dict, make_pem_inputs as make_inputs, random_rotation  # noqa: E402,F401
