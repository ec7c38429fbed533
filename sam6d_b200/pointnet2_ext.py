"""Drop-in for the reference's pybind module `pointnet2._ext` (PEM/model/pointnet2/_ext_src/src/bindings.cpp:11-24),
restricted to the four forward ops PEM inference calls.  Same names, argument order, dtypes and error behaviour
(RuntimeError for CPU / non-contiguous / wrong-dtype tensors); outputs are allocated by the callee on the current device and
the kernels run asynchronously on the current CUDA stream, like the reference.

    import sam6d_b200.pointnet2_ext as _ext     # instead of: import pointnet2._ext as _ext
"""
from . import ops


def furthest_point_sampling(points, nsamples):
    return ops.furthest_point_sampling(points, int(nsamples))


def gather_points(points, idx):
    return ops.gather_points(points, idx)


def ball_query(new_xyz, xyz, radius, nsample):
    return ops.ball_query(new_xyz, xyz, float(radius), int(nsample))


def group_points(points, idx):
    return ops.group_points(points, idx)


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError(f"pointnet2._ext.{name} is not on the SAM-6D inference path (training / unused op)")
    return fn


gather_points_grad = _unsupported("gather_points_grad")
group_points_grad = _unsupported("group_points_grad")
three_nn = _unsupported("three_nn")
three_interpolate = _unsupported("three_interpolate")
three_interpolate_grad = _unsupported("three_interpolate_grad")
