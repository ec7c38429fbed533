"""CPU: the cheap parts of oracle/dinov2_oracle.py against tests/golden/dinov2.pt (outputs of the reference's own
CropResizePad / CustomDINOv2 / vit_large, tools/make_golden_dinov2.py).  The 24-block ViT-L/14 itself is exercised on the GPU
side (minutes on the CPU); here: the proposal preprocessing and the positional-embedding interpolation."""
import os

import pytest
import torch

from oracle import dinov2_oracle as do


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "dinov2.pt"), weights_only=False)


def test_proposal_preprocessing_matches_reference(gold):
    image, masks, boxes = do.make_proposals(P=gold["meta"]["P"], seed=gold["meta"]["seed"])
    assert image.double().sum().item() == gold["input_checksum"]["image"] and masks.double().sum().item() == gold["input_checksum"]["masks"]
    assert torch.equal(boxes, gold["boxes"])
    rgbs = do.process_rgb_proposals(image, masks.clone(), boxes)
    pm = do.process_masks_proposals(masks.clone(), boxes)
    assert torch.equal(rgbs[:, :, ::7, ::7], gold["rgbs_sub"])
    torch.testing.assert_close(rgbs.double().sum(dim=(1, 2, 3)), gold["rgbs_sum"], atol=1e-6, rtol=0)
    assert torch.equal((pm > 0.5).to(torch.uint8), gold["pmasks_packed"])
    keep = torch.nn.functional.avg_pool2d(pm.unsqueeze(1), 14, 14).flatten(-2).squeeze(1) > 0.5
    assert torch.equal(keep, gold["keep"])


def test_pos_embed_interpolation_shapes():
    pe = torch.randn(1, 1 + 37 * 37, 64)
    out = do.interpolate_pos_encoding(pe, 256, 224, 224)
    assert out.shape == (1, 257, 64)
    assert torch.equal(out[:, 0], pe[:, 0])
    assert torch.equal(do.interpolate_pos_encoding(pe, 37 * 37, 518, 518), pe)
