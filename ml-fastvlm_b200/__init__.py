"""fastvithd-b200: B200-native FastViTHD vision tower + mm_projector for apple/ml-fastvlm.

Host side (Python/PyTorch plumbing) of `libfastvithd_b200.so` (hand-written sm_100a CUDA behind the
C ABI in include/fastvithd_b200.h).  Mirrors the reference's plugin surface for ONE path,
`LlavaMetaForCausalLM.encode_images` (llava/model/llava_arch.py:141-144):

    build_vision_tower      <- llava/model/multimodal_encoder/builder.py:6-19
    FastViTHDVisionTower    <- llava/model/multimodal_encoder/mobileclip_encoder.py:13-116
    build_vision_projector  <- llava/model/multimodal_projector/builder.py:17-35
    encode_images / EncodeImagesMixin / patch_llava  <- llava/model/llava_arch.py:141-144
    LlmPrefill              <- the first forward of LlavaQwen2ForCausalLM.generate (llava_qwen.py:57-143): time to first token

There is no CPU fallback: importing works anywhere, computing requires a B200 and the built library.
"""
from .lib import library_path, load_library, FvhdError  # noqa: F401
from .arch import reference_param_specs, projector_param_specs  # noqa: F401
from .packer import pack_tower, pack_projector  # noqa: F401
from .engine import Engine  # noqa: F401
from .tower import FastViTHDVisionTower, build_vision_tower  # noqa: F401
from .projector import FastVLMProjector, build_vision_projector  # noqa: F401
from .glue import (encode_images, splice_visual_tokens, splice_layout, prepare_inputs_embeds, EncodeImagesMixin, patch_llava,  # noqa: F401
                   IMAGE_TOKEN_INDEX)
from .checkpoint import read_state_dicts, load_pretrained  # noqa: F401
from .preprocess import (process_images, preprocess_into, resample_coeffs, process_anyres_image, select_best_resolution,  # noqa: F401
                         get_anyres_image_grid_shape, anyres_geometry)
from .llm import LlmPrefill, pack_qwen2  # noqa: F401
from .parallel import shard_bounds, shard_batch, all_gather_tokens, encode_images_sharded, gather_slots, GatheredEncoder  # noqa: F401

__version__ = "0.1.0"
