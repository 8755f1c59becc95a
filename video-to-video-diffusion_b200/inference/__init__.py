from .sampler import DDIMSampler, DDPMSampler

__all__ = ["DDIMSampler", "DDPMSampler"]
