"""Mirror of terrain_diffusion.models (reference: terrain_diffusion/models/)."""
from .edm_unet import EDMUnet2D  # noqa: F401
