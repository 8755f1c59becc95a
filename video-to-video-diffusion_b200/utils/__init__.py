from .metrics import calculate_psnr, calculate_ssim, calculate_video_metrics

__all__ = ["calculate_psnr", "calculate_ssim", "calculate_video_metrics"]
