"""helpers shared by the GPU parity tests"""
import math

import torch

from oracle import pem_oracle as po


def exact_indices(pts):
    """the embedding indices of transformer.py:302-332 evaluated in float64 (what the fp32 reference approximates:
    its expanded-form distance x2 - 2xy + y2 carries up to ~2e-3 of cancellation noise in d_idx, more on hosts whose
    fp32 matmul is not IEEE)"""
    p = pts.double()
    b, s, _ = p.shape
    dist = torch.cdist(p, p)
    knn = dist.topk(k=4, dim=2, largest=False)[1][:, :, 1:]
    knn_pts = torch.gather(p.unsqueeze(1).expand(b, s, s, 3), 2, knn.unsqueeze(3).expand(b, s, 3, 3))
    ref = (knn_pts - p.unsqueeze(2)).unsqueeze(2).expand(b, s, s, 3, 3)
    anc = (p.unsqueeze(1) - p.unsqueeze(2)).unsqueeze(3).expand(b, s, s, 3, 3)
    ang = torch.atan2(torch.linalg.norm(torch.cross(ref, anc, dim=-1), dim=-1), torch.sum(ref * anc, dim=-1))
    return (dist / po.SIGMA_D).float(), (ang * (180.0 / (po.SIGMA_A * math.pi))).float()



def exact_geo_embedding(sd, pts):
    """GeometricStructureEmbedding.forward (transformer.py:334-349) on float64-exact indices, projections in fp32"""
    d_idx, a_idx = exact_indices(pts)
    return (po._lin(sd, "geo_embedding.proj_d", po.sinusoidal_embedding(d_idx, 256)) +
            po._lin(sd, "geo_embedding.proj_a", po.sinusoidal_embedding(a_idx, 256)).max(dim=3)[0])
