"""Mirror of the reference `models/vae.py` interface: `SliceInterpolationVAE` / `VideoVAE` with the reference's
constructor, parameter names and creation order; `encode` / `decode` run on libb2v.so (b2v_vae_encode /
b2v_vae_decode): the same tcgen05 conv and fused GroupNorm kernels as the U-Net, 4x spatial, depth preserved.
"""
import torch
import torch.nn as nn

from .. import _lib
from ._native import NativeHandle, ParamsOnly, require_cuda


class Conv3DBlock(ParamsOnly):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride, padding)
        self.norm = nn.GroupNorm(8, out_channels)
        self.act = nn.SiLU()


class ResBlock3D(ParamsOnly):
    def __init__(self, channels):
        super().__init__()
        self.conv1 = Conv3DBlock(channels, channels)
        self.conv2 = nn.Sequential(nn.Conv3d(channels, channels, kernel_size=3, padding=1), nn.GroupNorm(8, channels))
        self.act = nn.SiLU()


class _Resample(ParamsOnly):
    def __init__(self, conv, out_channels):
        super().__init__()
        self.conv = conv
        self.norm = nn.GroupNorm(8, out_channels)
        self.act = nn.SiLU()


class DownsampleBlock(_Resample):
    def __init__(self, in_channels, out_channels):
        super().__init__(nn.Conv3d(in_channels, out_channels, (3, 4, 4), (1, 2, 2), (1, 1, 1)), out_channels)


class UpsampleBlock(_Resample):
    def __init__(self, in_channels, out_channels):
        super().__init__(nn.ConvTranspose3d(in_channels, out_channels, (3, 4, 4), (1, 2, 2), (1, 1, 1)), out_channels)


def _pair(ch):
    return nn.Sequential(ResBlock3D(ch), ResBlock3D(ch))


class VideoEncoder(ParamsOnly):
    def __init__(self, in_channels=3, latent_dim=4, base_channels=64):
        super().__init__()
        b = base_channels
        self.conv_in = Conv3DBlock(in_channels, b)
        self.down1 = nn.Sequential(ResBlock3D(b), ResBlock3D(b), DownsampleBlock(b, 2 * b))
        self.down2 = nn.Sequential(ResBlock3D(2 * b), ResBlock3D(2 * b), DownsampleBlock(2 * b, 4 * b))
        self.mid = _pair(4 * b)
        self.conv_out = nn.Conv3d(4 * b, 8, kernel_size=3, padding=1)
        self.quant_conv = nn.Conv3d(8, latent_dim, kernel_size=1)


class VideoDecoder(ParamsOnly):
    def __init__(self, latent_dim=4, out_channels=3, base_channels=64):
        super().__init__()
        b = base_channels
        self.post_quant_conv = nn.Conv3d(latent_dim, 8, kernel_size=1)
        self.conv_in = Conv3DBlock(8, 4 * b)
        self.mid = _pair(4 * b)
        self.up2_upsample = UpsampleBlock(4 * b, 2 * b)
        self.up2_res = _pair(2 * b)
        self.up3_upsample = UpsampleBlock(2 * b, b)
        self.up3_res = _pair(b)
        self.conv_out = nn.Conv3d(b, out_channels, kernel_size=3, padding=1)


class SliceInterpolationVAE(nn.Module):
    """x (B, C, T, H, W) in [-1, 1]  <->  z (B, latent_dim, T, H/4, W/4) (already scaled by scaling_factor)."""

    def __init__(self, in_channels=3, latent_dim=4, base_channels=64, scaling_factor=0.18215,
                 gradient_checkpointing=False):
        super().__init__()
        # limits of the B200 kernels (DESIGN.md "known limits"), raised here rather than at the first encode / decode
        if base_channels % 64:
            raise ValueError(f"SliceInterpolationVAE: base_channels = {base_channels} must be a multiple of 64")
        if latent_dim > 16 or in_channels > 16:
            raise ValueError("SliceInterpolationVAE: latent_dim and in_channels must be <= 16 (narrow-head kernels)")
        self.latent_dim, self.in_channels = latent_dim, in_channels
        self.base_channels = base_channels
        self.gradient_checkpointing = gradient_checkpointing
        self.encoder = VideoEncoder(in_channels, latent_dim, base_channels)
        self.decoder = VideoDecoder(latent_dim, in_channels, base_channels)
        self.scaling_factor = scaling_factor
        self._native = NativeHandle("vae")
        self._native_scale = None

    def native(self, device):
        if self._native_scale != float(self.scaling_factor):  # scaling_factor is folded into the packed weights
            self._native.close()
            self._native_scale = float(self.scaling_factor)
        d = _lib.VAEDesc(self.in_channels, self.latent_dim, self.base_channels, float(self.scaling_factor))
        return self._native.get(self, d, device)

    def invalidate_native(self):
        self._native.invalidate()

    @torch.no_grad()
    def encode(self, x):
        x = require_cuda(x, "VAE.encode")
        B, C, T, H, W = x.shape
        if C != self.in_channels:
            raise ValueError(f"VAE.encode: expected {self.in_channels} channels, got {C}")
        z = torch.empty((B, self.latent_dim, T, H // 4, W // 4), dtype=torch.float32, device=x.device)
        if z.numel() == 0:
            return z
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().b2v_vae_encode(self.native(x.device), _lib.dptr(x), _lib.dptr(z), B, T, H, W,
                                                 _lib.stream()), "vae_encode")
        return z

    @torch.no_grad()
    def decode(self, z):
        z = require_cuda(z, "VAE.decode")
        B, L, T, h, w = z.shape
        if L != self.latent_dim:
            raise ValueError(f"VAE.decode: expected {self.latent_dim} latent channels, got {L}")
        x = torch.empty((B, self.in_channels, T, 4 * h, 4 * w), dtype=torch.float32, device=z.device)
        if x.numel() == 0:
            return x
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().b2v_vae_decode(self.native(z.device), _lib.dptr(z), _lib.dptr(x), B, T, h, w,
                                                 _lib.stream()), "vae_decode")
        return x

    def encode_with_posterior(self, x):
        z = self.encode(x) / self.scaling_factor  # the reference splits the unscaled encoder output
        return torch.chunk(z, 2, dim=1)

    def forward(self, x):
        z = self.encode(x)
        return self.decode(z), z

    def get_latent_shape(self, volume_shape):
        B, C, T, H, W = volume_shape
        return (B, self.latent_dim, T, H // 4, W // 4)

    @classmethod
    def from_pretrained(cls, model_name_or_path, method="auto", inflate_method="central", strict=True, device="cpu",
                        **kwargs):
        raise NotImplementedError("from_pretrained() is not available; load a trained state_dict with "
                                  "load_state_dict() (same behaviour as the reference, models/vae.py:308-321)")


VideoVAE = SliceInterpolationVAE
