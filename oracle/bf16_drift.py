"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Calibrates the bf16 tolerance of the training-step parity test ("set the tolerance at about 2x the
reference's own bf16 drift"): gradients of the REAL reference under ``torch.autocast(bfloat16)`` against its fp32
gradients on the ``resdiff_grad_small`` case.  Measured in the build container (CPU):
    whole-gradient rel-L2 drift 7.6e-2, median per-tensor 3.2e-2, worst 0.53 (fd_spliter.channel_transform.weight,
    a 2-element sum with heavy cancellation).
Run:  python -m oracle.bf16_drift
"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import ref_shims
from oracle.cases import CASES, LINEAR_1000, fields
from oracle.weights import fill_module, seeded_randn
from oracle.make_golden import _unet
ref = ref_shims.import_reference()
spec = CASES['resdiff_grad_small']; cfg = spec['cfg']; b = spec['batch']; t = spec['t']; seed = spec['seed']
def run(autocast):
    net = fill_module(_unet(ref, cfg), seed).train()
    diff = ref.ResDiffDiffusion(net, image_height=32, image_width=64, channels=1, conditional=True)
    diff.set_new_noise_schedule(LINEAR_1000, 'cpu'); diff.set_loss('cpu')
    _, sr, hr = fields('resdiff_grad_small', b, 1, 32, 64, seed)
    noise = seeded_randn('resdiff_grad_small.noise', sr.shape, seed)
    sap = diff.sqrt_alphas_cumprod_prev
    u = np.random.RandomState(seed).uniform(sap[t-1], sap[t], size=b)
    np.random.randint = lambda *a, **k: t
    np.random.uniform = lambda *a, **k: u
    with torch.autocast('cpu', dtype=torch.bfloat16, enabled=autocast):
        loss = diff.p_losses({'HR': hr, 'SR': sr}, noise=noise)
    (loss.float().sum() / hr.numel()).backward()
    return float(loss), {n: p.grad.double() for n, p in net.named_parameters() if p.grad is not None}
l0, g0 = run(False)
l1, g1 = run(True)
num = sum(float((g1[n]-g0[n]).norm()**2) for n in g0); den = sum(float(g0[n].norm()**2) for n in g0)
print('loss fp32 %.4f bf16-autocast %.4f ; whole-gradient rel-L2 drift %.3e' % (l0, l1, (num/den)**0.5))
rels = sorted(((float((g1[n]-g0[n]).norm()/g0[n].norm().clamp_min(1e-30)), n) for n in g0), reverse=True)
print('worst tensors:', rels[:8]); print('median', rels[len(rels)//2])
