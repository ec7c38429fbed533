"""tools/ref_ism_import.py -- DEV CONTAINER ONLY (needs /root/reference).

Imports the reference's ISM Python modules (model/loss.py, model/detector.py, model/dinov2.py ...) as they are.  Their
missing third-party dependencies (ruamel.yaml, pytorch_lightning, hydra, omegaconf, trimesh, ...) never execute on the
scoring path -- they are only imported at module top level -- so each absent one is replaced by an empty stand-in module
in sys.modules; pl.LightningModule becomes torch.nn.Module.  No reference source is copied or modified."""
import importlib
import sys
import types

import torch

ISM = "/root/reference/SAM-6D/Instance_Segmentation_Model"


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return None


def _stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)
            sys.modules[n].__path__ = []


STUBBED = []


def import_reference_ism():
    """-> (loss module, detector module) of the reference"""
    if ISM not in sys.path:
        sys.path.insert(0, ISM)
    for n in ["ruamel", "ruamel.yaml", "pytorch_lightning", "hydra", "hydra.utils", "omegaconf", "trimesh", "pycocotools",
              "pycocotools.mask", "distinctipy", "skimage", "skimage.feature", "skimage.morphology", "imageio"]:   # (xformers is NOT stubbed: the reference falls back to eager attention on ImportError)
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
            STUBBED.append(n)
    import pytorch_lightning as pl
    if isinstance(pl, _Stub):
        pl.LightningModule = torch.nn.Module
    from model import loss, detector
    return loss, detector
