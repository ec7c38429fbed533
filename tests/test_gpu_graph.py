"""The captured forward (sam6d_b200/graph.py, Net.enable_graphs) against the launch-by-launch forward: same kernels on the same
inputs, so every output must be BIT-identical (no kernel on the path uses floating-point atomics), in both arithmetic modes,
for caller-provided uniforms; and the bookkeeping (one capture per input signature, launches counted per replay, another
weight version never replays a stale graph)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pem_oracle as po      # noqa: E402

KEYS = ("init_R", "init_t", "pred_R", "pred_t", "pred_pose_score")


def _net(cfg=None):
    from sam6d_b200.pem import Net
    net = (Net(cfg) if cfg is not None else Net()).cuda().eval()
    net.load_state_dict(po.make_state_dict(seed=1), strict=True)
    return net


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_is_bit_identical(precision):
    from sam6d_b200 import _lib
    net = _net().set_precision(precision)
    B = 2
    inp = po.make_inputs(B=B, n=2048, seed=3)
    dev = {k: inp[k].cuda() for k in ("pts", "dense_fm", "dense_po", "dense_fo", "model")}
    torch.manual_seed(5)
    rands = [torch.rand(B, po.N_PROPOSAL1 * 3, device="cuda") for _ in range(2)]
    want = [{k: v.clone() for k, v in net(dict(dev), rand=r).items() if k in KEYS} for r in rands]
    l0 = _lib.launch_count()
    net(dict(dev), rand=rands[0])
    per_step = _lib.launch_count() - l0
    net.enable_graphs()
    outs = []
    for i in range(4):                       # sighting (launch by launch), capture + replay, replay, replay
        l0 = _lib.launch_count()
        outs.append(net(dict(dev), rand=rands[i % 2]))
        assert _lib.launch_count() - l0 == per_step, "a replay counts the kernels its graph launches"
    torch.cuda.synchronize()
    sg = net._graphs
    assert not sg.disabled and sg.captures == 1 and sg.replays == 3
    for i, out in enumerate(outs):
        for k in KEYS:
            assert torch.equal(out[k], want[i % 2][k]), (precision, i, k)
    # results are copies: a later replay does not change what an earlier call returned
    assert torch.equal(outs[1]["pred_R"], want[1]["pred_R"]) and torch.equal(outs[2]["pred_R"], want[0]["pred_R"])
    # inputs are read in place: new values at the same addresses are this call's inputs
    dev["pts"].add_(0.01)
    fresh = net(dict(dev), rand=rands[0])
    net.disable_graphs()
    eager = net(dict(dev), rand=rands[0])
    for k in KEYS:
        assert torch.equal(fresh[k], eager[k]), k
    assert sg.captures == 1


def test_graph_not_replayed_after_weight_update():
    net = _net().set_precision("bf16").enable_graphs()
    B = 2
    inp = po.make_inputs(B=B, n=2048, seed=4)
    dev = {k: inp[k].cuda() for k in ("pts", "dense_fm", "dense_po", "dense_fo", "model")}
    rand = torch.rand(B, po.N_PROPOSAL1 * 3, device="cuda")
    for _ in range(3):
        before = net(dict(dev), rand=rand)
    assert net._graphs.replays == 2
    with torch.no_grad():
        net.fine_point_matching.out_proj.weight.mul_(-1.0)          # bumps the parameter version
    after = net(dict(dev), rand=rand)                               # new signature: launch by launch with the new weights
    assert net._graphs.replays == 2
    net.disable_graphs()
    eager = net(dict(dev), rand=rand)
    assert torch.equal(after["pred_R"], eager["pred_R"]) and torch.equal(after["init_R"], before["init_R"])


def test_graph_default_rand_draw():
    """rand=None: the uniforms are drawn per call (torch.rand, like the reference) outside the graph"""
    net = _net().set_precision("bf16").enable_graphs()
    inp = po.make_inputs(B=2, n=2048, seed=6)
    dev = {k: inp[k].cuda() for k in ("pts", "dense_fm", "dense_po", "dense_fo", "model")}
    outs = []
    for _ in range(4):
        out = net(dict(dev))
        outs.append(out["pred_R"].clone())
        R = out["pred_R"]
        assert torch.isfinite(R).all() and torch.allclose(R @ R.transpose(1, 2), torch.eye(3, device="cuda").expand_as(R), atol=1e-5)
    assert net._graphs.replays == 3
    buf = next(iter(net._graphs.graphs.values())).rand
    assert 0.45 < float(buf.mean()) < 0.55
