"""GPU, NCCL, world_size > 1 (skipped in a single-process run): object-sharded template scoring with the CUDA scorer and
proposal-sharded pose estimation reproduce the unsharded kernels bit for bit (BASELINE configs #4 / #5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        -m pytest tests/test_gpu_dist_nccl.py -q -x
Every rank runs the same assertions on its own GPU."""
import os

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu

WORLD = int(os.environ.get("WORLD_SIZE", "1"))


@pytest.fixture(scope="module")
def pg():
    if WORLD < 2:
        pytest.skip("needs torchrun with at least 2 ranks")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    yield dist.get_rank(), dist.get_world_size()


def test_object_sharded_scoring_equals_unsharded_kernel(pg):
    """config #5 shape: 21 objects x 42 templates x 1024-d, 200 proposals; each rank scores all proposals against its object shard
    with the fused CUDA kernel, one all-gather of 12 bytes per proposal and rank, identical result on every rank"""
    rank, world = pg
    from sam6d_b200 import dist as sdist, ism, synth
    q, r = synth.make_descriptors(P=200, O=21, T=42, C=1024, seed=5)
    q, r = q.cuda(), r.cuda()
    lo, hi = sdist.shard_range(21, rank, world)
    got = sdist.sharded_semantic_score(q, r[lo:hi].contiguous(), lo, confidence_thresh=0.2)
    want = ism.compute_semantic_score(q, r)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[3], want[3])
    assert torch.equal(got[2], want[2])                       # the per-object avg-5 score does not depend on the sharding: bitwise
    # more ranks than objects: some ranks own nothing
    lo1, hi1 = sdist.shard_range(1, rank, world)
    got1 = sdist.sharded_semantic_score(q, r[:1][lo1:hi1].contiguous(), lo1, confidence_thresh=0.2)
    want1 = ism.compute_semantic_score(q, r[:1].contiguous())
    assert all(torch.equal(a, b) for a, b in zip(got1, want1))


def test_proposal_sharded_pem_equals_unsharded(pg):
    """13 proposals (ragged over the ranks) through Net.forward on their shards + the pose all-gather == all 13 on one GPU"""
    rank, world = pg
    from sam6d_b200 import dist as sdist, synth
    from sam6d_b200.pem import Net
    B = 13
    net = Net(precision="bf16").cuda().eval()
    net.load_state_dict(synth.make_pem_state_dict(seed=1), strict=True)
    inp = {k: v.cuda() for k, v in synth.make_pem_inputs(B=B, n=2048, n_model=1024, seed=9).items()}
    torch.manual_seed(1)
    rand = torch.rand(B, synth.N_PROPOSAL1 * 3).cuda()
    keys = ("pts", "dense_fm", "dense_po", "dense_fo", "model")
    full = sdist.pack_poses(net({k: inp[k] for k in keys}, rand=rand))
    lo, hi = sdist.shard_range(B, rank, world)
    counts = [b - a for a, b in (sdist.shard_range(B, r, world) for r in range(world))]
    if hi > lo:
        local = sdist.pack_poses(net({k: inp[k][lo:hi].contiguous() for k in keys}, rand=rand[lo:hi].contiguous()))
    else:
        local = torch.zeros(0, sdist.POSE_FLOATS, device="cuda")
    got = sdist.all_gather_poses(local, counts=counts)
    assert got.shape == full.shape
    torch.testing.assert_close(got, full, atol=1e-6, rtol=0)   # proposals are independent: a shard computes the same numbers
