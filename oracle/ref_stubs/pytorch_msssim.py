class SSIM: pass
