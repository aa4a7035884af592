from .vector_quantizer import VectorQuantizer, all_reduce, get_world_size  # noqa: F401
from .vq_vae_patch_embedd import VQVAEPatch  # noqa: F401
