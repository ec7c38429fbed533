"""Host logic of the CUDA-graph step cache (sam6d_b200/graph.py) with the capture itself replaced by a stand-in: the policy
(first sighting launch by launch, second sighting captured, later ones replayed, least recently used graph dropped), the call
signature and the unpacking of the result block.  The real capture is covered on the GPU (tests/test_gpu_graph.py)."""
import torch

from sam6d_b200 import _lib, graph


class _FakeGraph:
    def __init__(self, log):
        self.log = log

    def replay(self):
        self.log.append("replay")


def _patched(monkeypatch, max_graphs=2):
    sg = graph.StepGraphs(max_graphs=max_graphs)
    log = []

    def fake_capture(fn, ep, n_rand):
        B = ep["pts"].shape[0]
        log.append("capture")
        return graph._Captured(_FakeGraph(log), torch.arange(B * graph.OUT_FLOATS, dtype=torch.float32), torch.zeros(B, n_rand), 7, B)

    def cpu_signature(ep, extra=()):
        return tuple((k, v.data_ptr(), tuple(v.shape)) for k, v in sorted(ep.items())
                     if isinstance(v, torch.Tensor) and k not in graph._OUT_NAMES) + tuple(extra)

    monkeypatch.setattr(sg, "_capture", fake_capture)
    monkeypatch.setattr(graph, "signature", cpu_signature)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    return sg, log


def test_signature_needs_device_tensors_and_ignores_results():
    ep = dict(pts=torch.zeros(2, 8, 3), note="x")
    assert graph.signature(ep) is None                     # host tensors: the call runs launch by launch (and fails there)
    assert graph.OUT_FLOATS == 25
    assert set(graph._OUT_NAMES) == {"init_R", "init_t", "pred_R", "pred_t", "pred_pose_score"}


def test_policy_sighting_capture_replay(monkeypatch):
    sg, log = _patched(monkeypatch)
    B = 3
    ep = dict(pts=torch.zeros(B, 8, 3), model=torch.zeros(B, 4, 3))
    eager = lambda e, r: e
    assert sg.run(eager, dict(ep), None, 6) is None and log == []            # first sighting
    l0 = _lib.launch_count()
    rand = torch.full((B, 6), 0.25)
    out = sg.run(eager, dict(ep), rand, 6)                                     # second: capture + replay
    assert log == ["capture", "replay"] and _lib.launch_count() - l0 == 7
    assert out["init_R"].shape == (B, 3, 3) and out["pred_t"].shape == (B, 3) and out["pred_pose_score"].shape == (B,)
    flat = torch.arange(B * 25, dtype=torch.float32)
    assert torch.equal(out["init_R"].reshape(-1), flat[:B * 9]) and torch.equal(out["init_t"].reshape(-1), flat[B * 9:B * 12])
    assert torch.equal(out["pred_pose_score"], flat[B * 24:])
    assert all(v.is_contiguous() for v in (out["init_R"], out["pred_R"], out["pred_t"]))
    cap = next(iter(sg.graphs.values()))
    assert torch.equal(cap.rand, rand)                                         # the caller's uniforms reach the graph's buffer
    out["init_R"].zero_()                                                      # results are copies: the graph's block is untouched
    assert cap.flat[1] == 1
    ep2 = dict(out)                                                            # a dict that carries results keeps its signature
    assert sg.run(eager, ep2, None, 6) is not None and log[-1] == "replay" and sg.captures == 0 and sg.replays == 2
    assert ((cap.rand >= 0) & (cap.rand < 1)).all() and not torch.equal(cap.rand, rand)   # rand=None: a fresh torch.rand draw


def test_policy_lru_and_weight_change(monkeypatch):
    sg, log = _patched(monkeypatch, max_graphs=2)
    eager = lambda e, r: e
    sets = [dict(pts=torch.zeros(2, 8, 3)) for _ in range(3)]
    for ep in sets:
        for _ in range(2):
            sg.run(eager, dict(ep), None, 6)
    assert log.count("capture") == 3 and len(sg.graphs) == 2                   # the oldest graph was dropped
    assert sg.run(eager, dict(sets[2]), None, 6) is not None and log.count("capture") == 3
    # another weight version / precision is another signature: sighting again, no stale replay
    assert sg.run(eager, dict(sets[2]), None, 6, extra=("bf16", 123)) is None


def test_capture_failure_turns_the_cache_off(monkeypatch):
    sg, log = _patched(monkeypatch)

    def boom(fn, ep, n_rand):
        raise RuntimeError("capture invalidated")

    monkeypatch.setattr(sg, "_capture", boom)
    ep = dict(pts=torch.zeros(2, 8, 3))
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert sg.run(lambda e, r: e, dict(ep), None, 6) is None
        assert sg.run(lambda e, r: e, dict(ep), None, 6) is None
    assert sg.disabled and any("launch by launch" in str(x.message) for x in w)
    assert sg.run(lambda e, r: e, dict(ep), None, 6) is None
