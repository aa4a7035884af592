"""B200-native vector-quantisation hot path for VQ-VAE-Patch (drop-in for the reference's
model/vector_quantizer.py and the encode call around it).  See DESIGN.md."""
from . import _lib, ops  # noqa: F401
from .model.vector_quantizer import VectorQuantizer, all_reduce, get_world_size  # noqa: F401
from .model.vq_vae_patch_embedd import VQVAEPatch  # noqa: F401

__all__ = ["VectorQuantizer", "VQVAEPatch", "ops", "all_reduce", "get_world_size"]
