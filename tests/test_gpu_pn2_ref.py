"""GPU: pins the point-cloud ops against the REFERENCE's own CUDA kernels.

oracle/_ref/pointnet2_ref_ext.so is the reference's pointnet2._ext compiled from the sources under /root/reference by
oracle/build_ref_ext.py (dev container) and shipped to the GPU box as a binary.  Here the reference kernels run on the B200
and both the C restatement (oracle/pn2_oracle.c) and the sam6d_b200 kernels must reproduce their index outputs bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import build_ref_ext, pn2     # noqa: E402


@pytest.fixture(scope="module")
def ref():
    mod = build_ref_ext.load_module()
    if mod is None:
        pytest.skip("oracle/_ref/ not present (reference extension is built only where /root/reference exists)")
    return mod


def _clouds(b, n, seed, dup=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n, 3, generator=g)
    x = x / x.norm(dim=2, keepdim=True) * (0.5 + 0.5 * torch.rand(b, n, 1, generator=g))
    if dup:
        pick = torch.randint(0, n // 6, (b, n), generator=g)
        x = torch.gather(x, 1, pick.unsqueeze(2).expand(b, n, 3))
    return x.contiguous()


@pytest.mark.parametrize("b,n,m,dup", [(4, 2048, 196, False), (2, 2048, 196, True), (2, 1000, 100, False), (1, 5000, 64, False),
                                        (2, 300, 40, True)])
def test_fps_matches_reference_kernel(ref, b, n, m, dup):
    from sam6d_b200 import ops
    x = _clouds(b, n, n + m, dup)
    want = ref.furthest_point_sampling(x.cuda(), m).cpu()
    assert torch.equal(pn2.furthest_point_sampling(x, m), want), "C restatement != reference CUDA kernel"
    assert torch.equal(ops.furthest_point_sampling(x.cuda(), m).cpu(), want), "sam6d_b200 kernel != reference CUDA kernel"


@pytest.mark.parametrize("b,n,m", [(1, 210000, 2048), (2, 50000, 512), (3, 4097, 100), (1, 106496, 300), (1, 106497, 300)])
def test_fps_cluster_kernel_matches_reference_kernel(ref, b, n, m):
    """large clouds: the thread-block-cluster FPS (8 / 16 CTAs per cloud, points in distributed shared memory) against the
    reference's own kernel and the one-CTA kernel; 210 000 -> 2048 is the template bank of get_obj_feats
    (PEM/model/feature_extraction.py:170-181).  Timings are written next to each other into gpurun_out/fps_big.json."""
    import json
    import os
    from sam6d_b200 import ops
    x = _clouds(b, n, n + m, dup=(n == 50000)).cuda()
    want = ref.furthest_point_sampling(x, m)
    got = ops.furthest_point_sampling(x, m)
    assert torch.equal(got, want), "cluster FPS != reference CUDA kernel"
    assert torch.equal(ops.furthest_point_sampling_single_cta(x, m), want)

    def t_ms(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    rec = dict(b=b, n=n, m=m, reference_ext_ms=t_ms(lambda: ref.furthest_point_sampling(x, m)),
               cluster_ms=t_ms(lambda: ops.furthest_point_sampling(x, m)), single_cta_ms=t_ms(lambda: ops.furthest_point_sampling_single_cta(x, m)))
    print("FPS", rec)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fps_big.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    old = json.load(open(path)) if os.path.exists(path) else []
    json.dump(old + [rec], open(path, "w"), indent=1)


@pytest.mark.parametrize("n,r,ns", [(2048, 0.1, 32), (2048, 0.2, 64), (700, 0.3, 16)])
def test_ball_query_matches_reference_kernel(ref, n, r, ns):
    from sam6d_b200 import ops
    x = _clouds(3, n, n + ns)
    want = ref.ball_query(x.cuda(), x.cuda(), r, ns).cpu()
    assert torch.equal(pn2.ball_query(x, x, r, ns), want)
    assert torch.equal(ops.ball_query(x.cuda(), x.cuda(), r, ns).cpu(), want)


def test_gather_group_match_reference_kernel(ref):
    from sam6d_b200 import ops
    g = torch.Generator().manual_seed(3)
    pts = torch.randn(2, 9, 500, generator=g)
    idx = torch.randint(0, 500, (2, 77), generator=g, dtype=torch.int32)
    gi = torch.randint(0, 500, (2, 77, 8), generator=g, dtype=torch.int32)
    assert torch.equal(ops.gather_points(pts.cuda(), idx.cuda()).cpu(), ref.gather_points(pts.cuda(), idx.cuda()).cpu())
    assert torch.equal(ops.group_points(pts.cuda(), gi.cuda()).cpu(), ref.group_points(pts.cuda(), gi.cuda()).cpu())
