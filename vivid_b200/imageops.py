"""Image-space helpers of the generation / evaluation drivers, each one CUDA pass in libvividb200.so (csrc/metrics.cu):
per-image PSNR (calculate_metrics.py:148) and the bilinear resizes of generate_images.py:282-283,322."""
import torch

from . import _lib as L


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t, what):
    if t.device.type != "cuda":
        raise RuntimeError(f"vivid_b200.metrics: {what} must live on a CUDA device; there is no CPU fallback")


def psnr_u8(images, tgt, cum=None):
    """Per-image PSNR (fp64 [N]) of uint8 `images` against `tgt` (uint8 or float in [0,255]); `cum[0] += sum`."""
    _require_cuda(images, "images")
    assert images.dtype == torch.uint8 and images.shape == tgt.shape
    images = images.contiguous()
    tgt = tgt.contiguous() if tgt.dtype == torch.uint8 else tgt.to(torch.float32).contiguous()
    n = images.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=images.device)
    if n == 0:
        return out
    per = images[0].numel()
    L.check(L.lib().vb_psnr_u8(images.data_ptr(), tgt.data_ptr(), L.VB_U8 if tgt.dtype == torch.uint8 else L.VB_F32, n, per,
                               per, out.data_ptr(), L.ptr(cum), _stream(images.device)), "vb_psnr_u8")
    return out


def resize_bilinear(x, size, antialias=False):
    """torch.nn.functional.interpolate(x, size=size, mode='bilinear', antialias=antialias) for fp32 NCHW on CUDA."""
    _require_cuda(x, "x")
    size = (size, size) if isinstance(size, int) else tuple(size)
    x = x.to(torch.float32).contiguous()
    out = torch.empty(x.shape[:-2] + size, dtype=torch.float32, device=x.device)
    L.check(L.lib().vb_resize(x.data_ptr(), out.data_ptr(), x.shape[0] * x.shape[1], x.shape[2], x.shape[3], size[0], size[1],
                              int(bool(antialias)), _stream(x.device)), "vb_resize")
    return out
