"""CPU, world_size 2 over gloo: proposal sharding and the single all-gather of final poses (sam6d_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sam6d_b200 import dist as sdist


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 32, 200):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    ep = dict(pred_R=torch.randn(5, 3, 3, generator=g), pred_t=torch.randn(5, 3, generator=g), pred_pose_score=torch.rand(5, generator=g))
    p = sdist.pack_poses(ep)
    assert p.shape == (5, 16)
    back = sdist.unpack_poses(p)
    for k in ep:
        assert torch.equal(back[k], ep[k])


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sdist.shard_range(total, rank, world)
    g = torch.Generator().manual_seed(123)
    allp = torch.randn(total, 16, generator=g)             # every rank can rebuild the full answer
    counts = [b - a for a, b in (sdist.shard_range(total, r, world) for r in range(world))]
    out = sdist.all_gather_poses(allp[lo:hi].clone(), counts=counts)
    q.put((rank, bool(torch.equal(out, allp))))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_all_gather_poses_world2_even_and_ragged():
    ctx = mp.get_context("spawn")
    for total in (32, 7):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in range(2))
        for p in procs:
            p.join(timeout=60)
        assert res == [(0, True), (1, True)]


def _oracle_local_best(desc, refs):
    """the CPU oracle as the per-shard scorer (the product path uses the CUDA kernel)"""
    from oracle import ism_oracle as io
    _, _, _, _, scores, per_obj = io.compute_semantic_score(desc, refs, confidence_thresh=-1.0)
    score, obj = per_obj.max(dim=1)
    tmpl = scores.argmax(dim=-1).gather(1, obj[:, None])[:, 0]
    return obj, score, tmpl


def _score_worker(rank, world, port, O, q):
    from oracle import ism_oracle as io
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc, refs = io.make_descriptors(P=24, O=O, T=42, C=64, seed=5)[:2]
    lo, hi = sdist.shard_range(O, rank, world)
    got = sdist.sharded_semantic_score(desc, refs[lo:hi], lo, confidence_thresh=0.2, local_best=_oracle_local_best)
    ref = io.compute_semantic_score(desc, refs, confidence_thresh=0.2)[:4]
    q.put((rank, all(bool(torch.equal(a, b)) for a, b in zip(got, ref))))
    dist.destroy_process_group()


def test_object_sharded_template_scoring_world2():
    """reference descriptors sharded by object, 12 bytes per proposal exchanged, identical
    (idx_selected, object, score, best_template) on every rank -- including a rank that owns no object."""
    ctx = mp.get_context("spawn")
    for O in (7, 1):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_score_worker, args=(r, 2, port, O, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in range(2))
        for p in procs:
            p.join(timeout=60)
        assert res == [(0, True), (1, True)]
