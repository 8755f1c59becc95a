"""B200-native (sm_100a) sampling hot path of Kkuntal990/video-to-video-diffusion.

Drop-in mirror of the reference's Python surface (same class names, signatures, state_dict keys):
    from v2v_b200.models import VideoToVideoDiffusion, UNet3D, VideoVAE, GaussianDiffusion
    from v2v_b200.inference import DDIMSampler, DDPMSampler
All compute goes through libb2v.so (include/b2v.h); see DESIGN.md.
"""
__version__ = "0.1.0"
