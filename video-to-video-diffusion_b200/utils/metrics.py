"""Mirror of the reference `utils/metrics.py` (PSNR, box-filter SSIM, per-slice video metrics) -- what
`Trainer.validate*` runs on every generated volume (training/trainer.py:365-371).  One fused kernel
(b2v_video_metrics) produces the per-slice squared-error and SSIM sums for the whole volume; a single device->host
copy of 2*T floats replaces the reference's two `.item()` synchronisations per slice.
"""
import math

import torch

from .. import _lib


def _slice_sums(v1, v2, max_val):
    """(B,C,T,H,W) x2 -> per-slice (sum sq err, sum ssim) tensor [T,2] on the host, and the element count per slice"""
    if not v1.is_cuda:
        raise RuntimeError("metrics run on the GPU (no CPU fallback)")
    B, C, T, H, W = v1.shape
    # (B,C,T,H,W) is already (BC, T, H, W) in memory
    a = v1.detach().float().contiguous()
    b = v2.detach().float().contiguous()
    out = torch.empty((T, 2), dtype=torch.float32, device=v1.device)
    _lib.check(_lib.lib().b2v_video_metrics(_lib.dptr(a), _lib.dptr(b), _lib.dptr(out), B * C, T, H, W, float(max_val),
                                            _lib.stream()), "video_metrics")
    return out.cpu(), B * C * H * W


def _psnr_from_mse(mse, max_val):
    mse = max(mse, 1e-8)
    return min(max(20.0 * math.log10(max_val / math.sqrt(mse)), 0.0), 100.0)


def _as5d(x):
    if x.dim() == 4:  # (B,C,H,W): one slice
        return x.unsqueeze(2)
    if x.dim() == 5:
        return x
    raise ValueError(f"expected a 4-D or 5-D tensor, got {tuple(x.shape)}")


def calculate_psnr(img1, img2, max_val=1.0):
    """PSNR in dB over the whole tensor (clamped to [0, 100] like the reference)"""
    s, n = _slice_sums(_as5d(img1.reshape(1, 1, *img1.shape[-3:]) if img1.dim() == 3 else img1),
                       _as5d(img2.reshape(1, 1, *img2.shape[-3:]) if img2.dim() == 3 else img2), max_val)
    return _psnr_from_mse(float(s[:, 0].sum()) / (n * s.shape[0]), max_val)


def calculate_ssim(img1, img2, window_size=11, max_val=1.0):
    """mean box-filter SSIM; 5-D inputs are averaged over depth slices (reference :66-82)"""
    if window_size != 11:
        raise NotImplementedError("the fused kernel implements the reference's default 11x11 window")
    s, n = _slice_sums(_as5d(img1), _as5d(img2), max_val)
    per = s[:, 1] / n
    if torch.isnan(per).any():
        return 0.0
    return float(per.mean())


def calculate_video_metrics(video1, video2, max_val=1.0):
    """{'psnr', 'ssim', 'psnr_per_frame', 'ssim_per_frame'} exactly like the reference (NaN inputs -> zeros)"""
    if video1.dim() == 4:
        video1, video2 = video1.unsqueeze(0), video2.unsqueeze(0)
    s, n = _slice_sums(video1, video2, max_val)
    if torch.isnan(s).any():  # the reference returns zeros when either input holds NaNs
        return {"psnr": 0.0, "ssim": 0.0, "psnr_per_frame": [], "ssim_per_frame": []}
    psnr = [_psnr_from_mse(float(s[t, 0]) / n, max_val) for t in range(s.shape[0])]
    ssim = [float(s[t, 1]) / n for t in range(s.shape[0])]
    return {"psnr": sum(psnr) / len(psnr) if psnr else 0.0, "ssim": sum(ssim) / len(ssim) if ssim else 0.0,
            "psnr_per_frame": psnr, "ssim_per_frame": ssim}
