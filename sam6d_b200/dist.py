"""Multi-GPU plumbing: proposals are independent, so the path shards by contiguous proposal chunks with no
data-path collective; the only exchange is one all-gather of the final poses (16 fp32 per pose: R 9, t 3, score 1, pad 3)
over NCCL / NVLink.  One process per GPU, torch.distributed for the rendezvous."""
from typing import Dict, Tuple

import torch
import torch.distributed as dist

POSE_FLOATS = 16


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous chunk [lo, hi) of `total` proposals owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_poses(end_points: Dict[str, torch.Tensor]) -> torch.Tensor:
    R, t, s = end_points["pred_R"], end_points["pred_t"], end_points["pred_pose_score"]
    B = R.shape[0]
    out = torch.zeros(B, POSE_FLOATS, dtype=torch.float32, device=R.device)
    out[:, :9] = R.reshape(B, 9)
    out[:, 9:12] = t
    out[:, 12] = s
    return out


def unpack_poses(p: torch.Tensor) -> Dict[str, torch.Tensor]:
    return dict(pred_R=p[:, :9].reshape(-1, 3, 3), pred_t=p[:, 9:12], pred_pose_score=p[:, 12])


def all_gather_poses(local: torch.Tensor, counts=None) -> torch.Tensor:
    """local (B_local,16) -> (sum B,16) on every rank.  Equal chunk sizes use one all_gather_into_tensor; ragged chunks
    (counts = per-rank sizes) pad to the maximum and trim."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None or len(set(counts)) == 1:
        out = torch.empty(world * local.shape[0], POSE_FLOATS, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    mx = max(counts)
    padded = torch.zeros(mx, POSE_FLOATS, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * mx, POSE_FLOATS, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)
