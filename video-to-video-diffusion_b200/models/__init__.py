from .vae import VideoVAE, SliceInterpolationVAE
from .unet3d import UNet3D
from .diffusion import GaussianDiffusion
from .model import VideoToVideoDiffusion

__all__ = ["VideoVAE", "SliceInterpolationVAE", "UNet3D", "GaussianDiffusion", "VideoToVideoDiffusion"]
