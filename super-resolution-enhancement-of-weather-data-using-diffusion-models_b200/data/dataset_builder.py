"""The part of the reference's DataLoader collate that produces the condition (data/dataset_builder.py:372-382):
``output_batch["SR"] = interpolate(LR, scale_factor=4, mode="bicubic")``, on the device and for the whole batch at once
(the reference loops over samples on the CPU).  No CPU fallback: tensors must live on a CUDA device."""
import torch

from .. import _native as nat


def bicubic_sr(lr, scale_factor=4):
    """LR (B, C, h, w) fp32 CUDA tensor -> SR (B, C, h*scale, w*scale), same arithmetic as F.interpolate(mode='bicubic')."""
    if not lr.is_cuda:
        raise nat.WsrError("bicubic_sr runs on the CUDA path only (got a %s tensor)" % lr.device)
    x = lr.to(torch.float32).contiguous()
    b, c, h, w = x.shape
    s = int(scale_factor)
    if s != scale_factor or s < 1:
        raise NotImplementedError("integer scale factors only (the reference uses 4 everywhere)")
    out = torch.empty((b, c, h * s, w * s), device=x.device, dtype=torch.float32)
    nat.call("wsr_bicubic_upsample", x.data_ptr(), b * c, h, w, s, out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    return out


def collate_batch(lr, hr):
    """The dict the reference's collate returns (dataset_builder.py:372-382) from already-stacked LR / HR device tensors."""
    return {"HR": hr, "LR": lr, "SR": bicubic_sr(lr, 4)}
