"""vision-zephyr_b200: the image -> LLM-embedding path of Vision-Zephyr on B200 (sm_100a).

Public surface (same names as the reference):
    build_vision_tower, build_multimodal_projector (alias build_vision_projector),
    VisZephyrB200MetaModel / VisZephyrB200MetaForCausalLM with encode_images and
    prepare_inputs_labels_for_multimodal,
plus the fused GPU preprocessing (process_any_resolution_images, process_fixed_images).
"""
from .constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX
from .anyres import (calculate_grid_shape, select_best_fit_resolution, unpad_bounds, anyres_views,
                     lanczos_table, slot_descriptor)
from .preprocess import (PatchBatch, VisualPrompt, clip_lut, lut_from_processor,
                         process_any_resolution_images, process_fixed_images, build_plan, run_plan)
from .vision_tower import CLIPVisionTowerB200, build_vision_tower
from .projector import QFormerB200, TextPack, build_multimodal_projector, build_vision_projector
from .arch import (VisZephyrB200MetaModel, VisZephyrB200MetaForCausalLM, merge_rows, splice_plan,
                   splice_scatter, text_gather)
from .text_inputs import tokenizer_image_token, collate_supervised
from .visual_prompts import PromptedImage, image_blending, resolve_visual_prompt
from . import dist as parallel

__all__ = [
    "IGNORE_INDEX", "IMAGE_TOKEN_INDEX", "build_vision_tower", "build_multimodal_projector",
    "build_vision_projector", "CLIPVisionTowerB200", "QFormerB200", "TextPack", "PatchBatch",
    "VisualPrompt", "VisZephyrB200MetaModel", "VisZephyrB200MetaForCausalLM",
    "process_any_resolution_images", "process_fixed_images", "clip_lut", "lut_from_processor",
    "calculate_grid_shape", "select_best_fit_resolution", "unpad_bounds", "merge_rows",
    "tokenizer_image_token", "collate_supervised", "image_blending", "resolve_visual_prompt", "PromptedImage",
]
