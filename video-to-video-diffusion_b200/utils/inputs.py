"""Input side of the sampling path on the device: CT windowing and the aligned thick-patch
extraction the reference does on the CPU in its datasets (data/slice_interpolation_dataset.py:575-592,
data/patch_slice_interpolation_dataset.py:118-195), fused into one kernel that feeds the VAE encoder directly."""
import torch

from .. import _lib

_INF = 3.0e38


def extract_thick_patch(thick_volume, z_thin_start, z_thin_end, depth_thin, y0, x0, patch_depth_thick=8,
                        patch_size=(192, 192), window=None, to_pm1=False):
    """thick_volume: (1, D_thick, H, W) or (D_thick, H, W) fp32 CUDA.  Maps the thin depth window
    [z_thin_start, z_thin_end) of a D_thin-slice volume onto thick slices exactly like the reference
    (int(z * D_thick / D_thin), at least one slice, clamped), crops (y0, x0, patch_size) and resamples the depth axis
    to `patch_depth_thick` slices.  window=(center, width): raw HU input, windowed to [0, 1] first; to_pm1: then
    mapped to [-1, 1].  Returns (1, patch_depth_thick, ph, pw)."""
    vol = thick_volume.reshape(thick_volume.shape[-3:]).contiguous().float()
    if not vol.is_cuda:
        raise RuntimeError("extract_thick_patch runs on the GPU (no CPU fallback)")
    D, H, W = vol.shape
    ph, pw = patch_size
    z0 = int(z_thin_start * D / depth_thin)
    z1 = int(z_thin_end * D / depth_thin)
    if z1 <= z0:
        z1 = z0 + 1
    z0, z1 = max(0, z0), min(D, z1)
    lo, hi, a, b = -_INF, _INF, 1.0, 0.0
    if window is not None:
        center, width = window
        lo, hi = center - width / 2, center + width / 2
        a, b = 1.0 / (hi - lo), -lo / (hi - lo)
    if to_pm1:
        a, b = 2.0 * a, 2.0 * b - 1.0
    out = torch.empty((1, patch_depth_thick, ph, pw), dtype=torch.float32, device=vol.device)
    _lib.check(_lib.lib().b2v_extract_patch(_lib.dptr(vol), _lib.dptr(out), D, H, W, z0, z1, int(y0), int(x0),
                                            patch_depth_thick, ph, pw, lo, hi, a, b, _lib.stream()), "extract_patch")
    return out


def apply_ct_windowing(volume_hu, window_center, window_width):
    """(D, H, W) Hounsfield units -> [0, 1] (reference _apply_ct_windowing), on the device"""
    D, H, W = volume_hu.shape[-3:]
    return extract_thick_patch(volume_hu, 0, D, D, 0, 0, D, (H, W), window=(window_center, window_width))[0]
