"""Pin the oracle (oracle/sam_ref.py) against an independent implementation of the same arithmetic:
HuggingFace transformers' SamModel, fed the SAME seeded weights through a key map.  The reference
(micro-sam) holds no numeric fixture for embeddings/logits, so this is the pin."""
import numpy as np
import pytest
import torch

from oracle import sam_ref
from tests.hf_map import upstream_to_hf


def _hf_model(model_type):
    from transformers import SamConfig, SamMaskDecoderConfig, SamModel, SamPromptEncoderConfig, SamVisionConfig
    a = sam_ref.ARCH[model_type]
    vc = SamVisionConfig(hidden_size=a["embed_dim"], num_hidden_layers=a["depth"], num_attention_heads=a["num_heads"],
                         global_attn_indexes=list(a["global_attn_indexes"]), mlp_dim=4 * a["embed_dim"],
                         output_channels=256, image_size=1024, patch_size=16, window_size=14)
    cfg = SamConfig(vision_config=vc, prompt_encoder_config=SamPromptEncoderConfig(),
                    mask_decoder_config=SamMaskDecoderConfig(layer_norm_eps=1e-5))
    return SamModel(cfg).eval()


# vit_b = a real architecture of micro_sam/models/build_sam.py:40-76: 12 heads (head-major qkv split), 4 global blocks
# (2, 5, 8, 11), depth 12 -- the toy shapes alone would not pin those.
@pytest.mark.parametrize("model_type", ["vit_test", "vit_test80", "vit_b"])
def test_oracle_matches_hf(model_type):
    torch.manual_seed(0)
    sd = sam_ref.seeded_state_dict(model_type, seed=1)
    sam = sam_ref.build_sam(model_type)
    sam.load_state_dict(sd)
    hf = _hf_model(model_type)
    hf_sd = {upstream_to_hf(k): v for k, v in sd.items()}
    missing, unexpected = hf.load_state_dict(hf_sd, strict=False)
    assert not unexpected, unexpected
    assert not [m for m in missing if "shared_image_embedding" not in m and "position" not in m], missing

    x = torch.rand(1, 3, 1024, 1024) * 255
    xin = sam.preprocess(x)
    with torch.no_grad():
        feat = sam.image_encoder(xin)
        feat_hf = hf.get_image_embeddings(xin)
    assert feat.shape == (1, 256, 64, 64)
    np.testing.assert_allclose(feat.numpy(), feat_hf.numpy(), rtol=1e-3, atol=2e-4)

    pred = sam_ref.SamPredictor(sam)
    pred.features, pred.is_image_set = feat, True
    pred.original_size = pred.input_size = (1024, 1024)
    pts = torch.tensor([[[100.0, 200.0]], [[700.5, 333.25]], [[512.0, 512.0]]])
    lbl = torch.ones(3, 1, dtype=torch.int)
    _, iou, low = pred.predict_torch(pts, lbl, multimask_output=True, return_logits=True)
    with torch.no_grad():
        out = hf(image_embeddings=feat_hf, input_points=pts[None], input_labels=lbl[None], multimask_output=True)
    np.testing.assert_allclose(low.numpy(), out.pred_masks[0].numpy(), rtol=1e-3, atol=5e-4)
    np.testing.assert_allclose(iou.numpy(), out.iou_scores[0].numpy(), rtol=1e-3, atol=5e-4)

    # box prompts (no padding point)
    boxes = torch.tensor([[100.0, 120.0, 300.0, 400.0], [10.0, 20.0, 1000.0, 900.0]])
    _, iou_b, low_b = pred.predict_torch(None, None, boxes=boxes, multimask_output=False, return_logits=True)
    with torch.no_grad():
        out = hf(image_embeddings=feat_hf, input_boxes=boxes[None], multimask_output=False)
    np.testing.assert_allclose(low_b.numpy(), out.pred_masks[0].numpy(), rtol=1e-3, atol=5e-4)
    np.testing.assert_allclose(iou_b.numpy(), out.iou_scores[0].numpy(), rtol=1e-3, atol=5e-4)
