"""Drop-in names of the reference's `models` package, backed by libb2v.so (see ../_lib.py, include/b2v.h).

    UNet3D                     eps-predictor; forward = b2v_unet_forward
    SliceInterpolationVAE      4x-spatial 3-D autoencoder (alias VideoVAE); encode / decode = b2v_vae_*
    GaussianDiffusion          schedule buffers + ancestral sampling loop (b2v_ddpm_step)
    VideoToVideoDiffusion      facade: config resolution, generate()
"""
from . import diffusion as _diffusion
from . import model as _model
from . import unet3d as _unet3d
from . import vae as _vae

UNet3D = _unet3d.UNet3D
SliceInterpolationVAE = _vae.SliceInterpolationVAE
VideoVAE = _vae.VideoVAE
GaussianDiffusion = _diffusion.GaussianDiffusion
VideoToVideoDiffusion = _model.VideoToVideoDiffusion

__all__ = ("UNet3D", "SliceInterpolationVAE", "VideoVAE", "GaussianDiffusion", "VideoToVideoDiffusion")
