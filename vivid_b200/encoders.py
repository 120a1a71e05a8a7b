"""Pixel codec — mirror of training/encoders.py:50-62 (StandardRGBEncoder), on CUDA."""
import torch

from . import _lib as L


class StandardRGBEncoder:
    def __init__(self):
        self.device = None

    def init(self, device):
        self.device = device

    def encode_pixels(self, x):
        return x

    def encode(self, x):
        return self.encode_latents(self.encode_pixels(x))

    def encode_latents(self, x):
        """raw uint8-valued pixels -> [-1, 1] : x/127.5 - 1."""
        if x.dtype == torch.uint8 and x.device.type == "cuda":
            x = x.contiguous()
            out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            L.check(L.lib().vb_encode_u8(x.data_ptr(), out.data_ptr(), x.numel(),
                                         torch.cuda.current_stream(x.device).cuda_stream), "vb_encode_u8")
            return out
        return x.to(torch.float32) / 127.5 - 1

    def decode(self, x):
        """latents -> uint8 pixels : clip(x*127.5 + 128, 0, 255)."""
        if x.device.type != "cuda":
            raise RuntimeError("vivid_b200 codec runs on CUDA only")
        x = x.to(torch.float32).contiguous()
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        L.check(L.lib().vb_decode_u8(x.data_ptr(), out.data_ptr(), x.numel(),
                                     torch.cuda.current_stream(x.device).cuda_stream), "vb_decode_u8")
        return out
