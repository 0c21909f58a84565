"""Pre-training loss of the SimpleCNN prior -- ``image_compare_loss`` / ``fft_mse_loss`` / ``dwt_mse_loss`` with the reference's
signatures (models/simple_cnn/loss.py:9-76), computed by ONE kernel launch on the CUDA path (``wsr_image_compare_loss``):

    fft_mse_loss(x, y) = MSE(Re F(x), Re F(y)) + MSE(Im F(x), Im F(y)),  F = orthonormal 2-D FFT
                       = MSE(x, y)                                      (F is linear and unitary)
    dwt_mse_loss(x, y) = sum over J = 4 Haar levels and the 3 detail bands of MSE(band(x), band(y)) = the same of band(x - y)

The returned 0-dim tensors carry a ``grad_fn`` (the gradient w.r.t. ``x`` comes out of the same launch), so the reference's
``loss = criterion(outputs, targets); loss.backward()`` (pretrain.py:45-48) works unchanged."""
import torch

from ... import _native as nat


class _CompareLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, alpha, beta):
        if not x.is_cuda:
            raise nat.WsrError("image_compare_loss runs on the CUDA path only (got a %s tensor)" % x.device)
        xs = x.detach().to(torch.float32).contiguous()
        ys = y.detach().to(device=x.device, dtype=torch.float32).contiguous()
        assert xs.shape == ys.shape and xs.dim() == 4
        b, c, h, w = xs.shape
        acc = torch.zeros(1, dtype=torch.float64, device=x.device)
        grad = torch.empty_like(xs)
        nat.call("wsr_image_compare_loss", xs.data_ptr(), ys.data_ptr(), b * c, h, w, float(alpha), float(beta), acc.data_ptr(),
                 grad.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
        ctx.save_for_backward(grad)
        return acc.to(torch.float32)[0]

    @staticmethod
    def backward(ctx, go):
        (grad,) = ctx.saved_tensors
        return grad * go.to(grad.dtype), None, None, None


def fft_mse_loss(img1, img2):
    return _CompareLossFn.apply(img1, img2, 1.0, 0.0)


def dwt_mse_loss(x, y, J=4):
    if J != 4:
        raise NotImplementedError("the kernel implements the reference's J = 4 levels")
    return _CompareLossFn.apply(x, y, 0.0, 1.0)


def image_compare_loss(x, y, alpha=0.2, beta=0.1):
    return _CompareLossFn.apply(x, y, alpha, beta)
