"""CPU: the C restatement of the reference's CUDA-only PointNet++ ops (oracle/pn2_oracle.c)
against brute-force definitions and the edge cases the kernels define
(PEM/model/pointnet2/_ext_src/src/{sampling_gpu,ball_query_gpu,group_points_gpu}.cu)."""
import numpy as np
import torch

from oracle import pn2


def test_fps_basic_and_ties():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 300, 3, generator=g)
    idx = pn2.furthest_point_sampling(x, 40)
    assert idx.dtype == torch.int32 and idx.shape == (3, 40)
    assert (idx[:, 0] == 0).all()                      # sampling_gpu.cu:90-91
    for b in range(3):
        assert len(set(idx[b].tolist())) == 40
    # greedy definition: every pick maximises the running min distance
    for b in range(3):
        p = x[b].numpy().astype(np.float32)
        mind = np.full(300, 1e10, np.float32)
        for j in range(1, 40):
            d = ((p - p[idx[b, j - 1]]) ** 2).sum(1)
            mind = np.minimum(mind, d)
            assert abs(mind[idx[b, j]] - mind.max()) <= 1e-5 * max(1.0, mind.max())


def test_fps_duplicates_tie_rule():
    # all-identical points: every distance ties at 0 -> slot 0 survives the tree -> 0
    x = torch.ones(1, 100, 3)
    idx = pn2.furthest_point_sampling(x, 5)
    assert idx.tolist() == [[0, 0, 0, 0, 0]]
    # exact ties: 1 and 65 share thread 1 (bs = 64) -> smallest k; thread 2 (bit-reversed 16) beats thread 1
    # (bit-reversed 32) because the tree folds slot t+s into slot t and a tie keeps slot t
    x = torch.zeros(1, 100, 3)
    x[0, 1] = x[0, 65] = torch.tensor([1.0, 0, 0])
    assert pn2.lib().pn2_oracle_block_size(100) == 64
    assert pn2.furthest_point_sampling(x, 2).tolist() == [[0, 1]]
    x[0, 2] = torch.tensor([1.0, 0, 0])
    assert pn2.furthest_point_sampling(x, 2).tolist() == [[0, 2]]


def test_ball_query_semantics():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 200, 3, generator=g)
    q = x[:, :50].contiguous()
    r, ns = 0.25, 16
    idx = pn2.ball_query(q, x, r, ns)
    for b in range(2):
        for j in range(50):
            d2 = ((x[b] - q[b, j]) ** 2).sum(1)
            hits = torch.nonzero(d2 < np.float32(r) * np.float32(r) - 1e-6).flatten().tolist()
            got = idx[b, j].tolist()
            n = min(len(hits), ns)
            assert got[:n] == hits[:n] or abs(len(hits) - len(set(got))) <= 1
            if n < ns and n > 0:
                assert all(v == got[0] for v in got[n:])  # padded with the first hit
    # no hit at all -> zeros (ball_query.cpp:24-26)
    far = torch.full((1, 3, 3), 50.0)
    assert pn2.ball_query(far, x[:1], 0.1, 4).abs().sum() == 0


def test_gather_group():
    g = torch.Generator().manual_seed(2)
    pts = torch.randn(2, 5, 30, generator=g)
    idx = torch.randint(0, 30, (2, 7), generator=g, dtype=torch.int32)
    out = pn2.gather_points(pts, idx)
    ref = torch.gather(pts, 2, idx.long().unsqueeze(1).expand(2, 5, 7))
    assert torch.equal(out, ref)
    gi = torch.randint(0, 30, (2, 7, 4), generator=g, dtype=torch.int32)
    out = pn2.group_points(pts, gi)
    ref = torch.gather(pts.unsqueeze(2).expand(2, 5, 7, 30), 3, gi.long().unsqueeze(1).expand(2, 5, 7, 4))
    assert torch.equal(out, ref)
