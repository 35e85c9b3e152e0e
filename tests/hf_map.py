"""Key map upstream-SAM state dict -> HuggingFace ``SamModel``.  Test helper."""
import re


def upstream_to_hf(key: str) -> str:
    k = key
    k = k.replace("image_encoder.", "vision_encoder.")
    k = k.replace("vision_encoder.blocks.", "vision_encoder.layers.")
    k = re.sub(r"(vision_encoder\.layers\.\d+)\.norm1\.", r"\1.layer_norm1.", k)
    k = re.sub(r"(vision_encoder\.layers\.\d+)\.norm2\.", r"\1.layer_norm2.", k)
    k = k.replace("patch_embed.proj.", "patch_embed.projection.")
    k = k.replace("neck.0.", "neck.conv1.").replace("neck.1.", "neck.layer_norm1.")
    k = k.replace("neck.2.", "neck.conv2.").replace("neck.3.", "neck.layer_norm2.")
    if k == "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix":
        return "shared_image_embedding.positional_embedding"
    k = k.replace("prompt_encoder.point_embeddings.", "prompt_encoder.point_embed.")
    for a, b in (("0", "conv1"), ("1", "layer_norm1"), ("3", "conv2"), ("4", "layer_norm2"), ("6", "conv3")):
        k = k.replace(f"mask_downscaling.{a}.", f"mask_embed.{b}.")
    if k.startswith("mask_decoder.transformer.layers."):
        k = re.sub(r"\.norm(\d)\.", r".layer_norm\1.", k)
    k = k.replace("norm_final_attn.", "layer_norm_final_attn.")
    k = k.replace("output_upscaling.0.", "upscale_conv1.").replace("output_upscaling.1.", "upscale_layer_norm.")
    k = k.replace("output_upscaling.3.", "upscale_conv2.")
    m = re.match(r"(mask_decoder\.(?:output_hypernetworks_mlps\.\d+|iou_prediction_head))\.layers\.(\d)\.(.*)", k)
    if m:
        name = {"0": "proj_in", "1": "layers.0", "2": "proj_out"}[m.group(2)]
        k = f"{m.group(1)}.{name}.{m.group(3)}"
    return k
