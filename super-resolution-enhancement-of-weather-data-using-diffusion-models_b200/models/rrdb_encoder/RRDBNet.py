"""RRDB encoder -- drop-in for the reference's models/rrdb_encoder/RRDBNet.py:11-133 (same class names, constructor
arguments and state_dict keys).  ``forward(x, get_fea)`` returns ``(sr_image, [18 feature maps])`` like the reference;
all convolutions run in the CUDA engine:

  * every ResidualDenseBlock owns ONE NHWC buffer of nf + 4*gc channels; conv_k reads the channel prefix and writes its
    growth slice in place (LeakyReLU fused), so the reference's ``torch.cat`` (:106-110) never materialises;
  * ``x5 * 0.2 + x`` and the RRDB-level ``out * 0.2 + x`` (:111, :133) are epilogue scale/residual terms;
  * channel prefixes that are not a multiple of 64 (96, 160) are read as the next multiple of 64 with zero weights, so
    all 255 dense-block convolutions are eligible for the tcgen05 kernel.
"""
import functools

import torch
from torch import nn

from ... import _native as nat
from ...engine import Engine


def make_layer(block, n_layers, seq=False):
    layers = [block() for _ in range(n_layers)]
    return nn.Sequential(*layers) if seq else nn.ModuleList(layers)


class ResidualDenseBlock_5C(nn.Module):
    def __init__(self, nf=64, gc=32, bias=True):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, gc, 3, 1, 1, bias=bias)
        self.conv2 = nn.Conv2d(nf + gc, gc, 3, 1, 1, bias=bias)
        self.conv3 = nn.Conv2d(nf + 2 * gc, gc, 3, 1, 1, bias=bias)
        self.conv4 = nn.Conv2d(nf + 3 * gc, gc, 3, 1, 1, bias=bias)
        self.conv5 = nn.Conv2d(nf + 4 * gc, nf, 3, 1, 1, bias=bias)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)


class RRDB(nn.Module):
    def __init__(self, nf, gc=32):
        super().__init__()
        self.RDB1 = ResidualDenseBlock_5C(nf, gc)
        self.RDB2 = ResidualDenseBlock_5C(nf, gc)
        self.RDB3 = ResidualDenseBlock_5C(nf, gc)


class _RRDBPlan:
    """Buffers + packed weights for one (batch, h, w, precision)."""

    def __init__(self, net, B, h, w, device, precision):
        self.net, self.B, self.h, self.w = net, B, h, w
        e = self.eng = Engine(device, precision)
        nf, gc = net.nf, net.gc
        self.wide = nf + 4 * gc
        nb = len(net.RRDB_trunk)
        self.x0 = e.new_act(B, h, w, net.in_nc, dt=nat.F32)
        self.first = e.new_act(B, h, w, nf)
        # dense buffers: one per RDB (zero-initialised: padded reads hit zero weights, the data must stay finite)
        self.dense = [[e.new_act(B, h, w, self.wide, zero=True) for _ in range(3)] for _ in range(nb)]
        self.out_last = e.new_act(B, h, w, nf)          # output of the last RRDB
        self.trunk = e.new_act(B, h, w, nf)
        self.up1 = e.new_act(B, 2 * h, 2 * w, nf)
        self.up2 = e.new_act(B, 4 * h, 4 * w, nf)
        self.hr = e.new_act(B, 4 * h, 4 * w, nf)
        self.last = e.new_act(B, 4 * h, 4 * w, net.out_nc, dt=nat.F32)
        self._wver = None
        self.refresh()

    def refresh(self):
        v = tuple(p._version for p in self.net.parameters()) + tuple(p.data_ptr() for p in self.net.parameters())
        if v == self._wver:
            return
        self._wver = v
        e, net = self.eng, self.net
        bf = e.mode == "bf16"

        def pad64(c):
            return (c + 63) // 64 * 64 if bf else c

        with torch.no_grad():
            self.w_first = self._pack_f32(net.conv_first)
            self.blocks = []
            for rrdb in net.RRDB_trunk:
                packs = []
                for rdb in (rrdb.RDB1, rrdb.RDB2, rrdb.RDB3):
                    convs = [rdb.conv1, rdb.conv2, rdb.conv3, rdb.conv4, rdb.conv5]
                    packs.append([e.pack_conv(c.weight, c.bias, cin_pad=min(pad64(c.in_channels), self.wide),
                                              rows=64 if bf else None) for c in convs])
                self.blocks.append(packs)
            self.w_trunk = e.pack_conv(net.trunk_conv.weight, net.trunk_conv.bias)
            self.w_up1 = e.pack_conv(net.upconv1.weight, net.upconv1.bias)
            self.w_up2 = e.pack_conv(net.upconv2.weight, net.upconv2.bias)
            self.w_hr = e.pack_conv(net.HRconv.weight, net.HRconv.bias)
            self.w_last = e.pack_conv(net.conv_last.weight, net.conv_last.bias, rows=64 if bf else None)
        torch.cuda.current_stream(e.device).synchronize()
        e._keep.clear()

    def _pack_f32(self, conv):
        from ...engine import PackedConv
        e = self.eng
        w = e.f32(conv.weight)
        Cout, Cin, KH, KW = w.shape
        pc = PackedConv()
        pc.Cout, pc.Cin, pc.k, pc.Cin_pad, pc.rows = Cout, Cin, KH, Cin, Cout
        pc.bias = e.f32(conv.bias)
        pc.w = e.empty((KH * KW, Cout, Cin), torch.float32)
        nat.call("wsr_pack_conv_weight", w.data_ptr(), Cout, Cin, KH, KW, pc.w.data_ptr(), nat.F32, Cout, Cin, e.stream)
        e._keep.append(w)
        return pc

    def run(self, x):
        """x: LR (B, C, h, w) fp32.  Returns (list of nb+1 feature Acts, fea Act) -- features stay on the device."""
        e, net = self.eng, self.net
        nf, gc = net.nf, net.gc
        LR = nat.ACT_LRELU02
        xin = x.to(torch.float32).contiguous()
        ones = torch.ones_like(xin)
        x01 = torch.empty_like(xin)
        # x = (x + 1) / 2   (RRDBNet.py:40)
        nat.call("wsr_axpby", xin.data_ptr(), nat.F32, 1, 0.5, ones.data_ptr(), nat.F32, 1, 0.5, x01.data_ptr(), nat.F32, 1,
                 xin.numel(), 1, e.stream)
        e.nchw_to_act(x01, self.x0)
        nb = len(self.blocks)
        e.conv(self.x0, self.w_first, self.first, force_simt=True)
        # the trunk input lives in channel slot [0, nf) of the first dense buffer
        nat.call("wsr_axpby", self.first.ptr, self.first.dt, self.first.ld, 1.0, self.first.ptr, self.first.dt, self.first.ld, 0.0,
                 self.dense[0][0].ptr, self.dense[0][0].dt, self.dense[0][0].ld, self.B * self.h * self.w, nf, e.stream)
        feas = []
        for i, packs in enumerate(self.blocks):
            rrdb_in = self.dense[i][0].slice(0, nf)
            for j, pk in enumerate(packs):
                D = self.dense[i][j]
                for k in range(4):
                    e.conv(D.slice(0, pk[k].Cin_pad), pk[k], D.slice(nf + k * gc, gc), act=LR)
                xj = D.slice(0, nf)
                if j < 2:
                    e.conv(D.slice(0, pk[4].Cin_pad), pk[4], self.dense[i][j + 1].slice(0, nf), out_scale=0.2, res=xj)
                else:
                    dst = self.dense[i + 1][0].slice(0, nf) if i + 1 < nb else self.out_last
                    # (conv5*0.2 + x3)*0.2 + rrdb_in  (RRDBNet.py:111, 133)
                    e.conv(D.slice(0, pk[4].Cin_pad), pk[4], dst, out_scale=0.04, res=xj, res_scale=0.2, res2=rrdb_in)
                    feas.append(dst)
        e.conv(feas[-1], self.w_trunk, self.trunk, res=self.first)            # fea_first + trunk_conv(fea)
        feas.append(self.trunk)
        self.last_feas = feas
        return feas

    def sr_image(self):
        """The upsampling tail (RRDBNet.py:49-55); only needed when the encoder's own SR output is used."""
        e = self.eng
        LR = nat.ACT_LRELU02
        e.conv(self.trunk, self.w_up1, self.up1, upsample=True, act=LR)
        e.conv(self.up1, self.w_up2, self.up2, upsample=True, act=LR)
        e.conv(self.up2, self.w_hr, self.hr, act=LR)
        e.conv(self.hr, self.w_last, self.last)
        return self.last.to_nchw(e).clamp(0, 1) * 2 - 1

    def sr_image_raw(self):
        """conv_last output before ``clamp(0, 1) * 2 - 1`` (:55-57), NCHW fp32 -- the pre-training path applies the clamp with
        differentiable torch ops so that its mask reaches ``backward``."""
        self.sr_image()
        return self.last.to_nchw(self.eng)


    # ---- pre-training: backward pass (reference pretrain.py:45-48 with criterion = F.l1_loss) ------------------
    def _axpby(self, x, a, z, b, y):
        nat.call("wsr_axpby", x.ptr, x.dt, x.ld, float(a), z.ptr, z.dt, z.ld, float(b), y.ptr, y.dt, y.ld, y.N * y.H * y.W, y.C, self.eng.stream)

    def _mask(self, y, dy):
        nat.call("wsr_lrelu_mask", y.ptr, y.dt, y.ld, dy.ptr, dy.dt, dy.ld, y.N * y.H * y.W, y.C, 0.2, self.eng.stream)

    def _dgrad_pack(self, conv, up=False):
        from ... import taps as T
        e = self.eng
        w = conv.weight
        if up:
            return e.pack_conv(T.upsample_dgrad_weight(w), None, key=("updgrad", w.data_ptr(), w._version))
        return e.pack_conv(T.dgrad_weight(w).contiguous(), None, key=("dgrad", w.data_ptr(), w._version))

    def _wgrad(self, conv, x, dy, table, up=1, force_simt=False):
        for prm in (conv.weight, conv.bias):
            if prm.grad is None:
                prm.grad = torch.zeros_like(prm)
        co, ci, kh, kw = conv.weight.shape
        xs = x if x.C == ci else x.slice(0, ci)
        self.eng.wgrad(xs, dy, table, conv.weight.grad, (1, ci * kh * kw, kh * kw), conv.bias.grad, up, force_simt=force_simt)

    @torch.no_grad()
    def backward(self, d_raw=None, d_cond=None):
        """Parameter gradients (accumulated into ``p.grad``) from the gradient w.r.t. the RAW ``conv_last`` output (B, out_nc, 4h, 4w)
        and / or w.r.t. the condition cat(feas[2::3]) (B, 6 nf, h, w) that SRDiff's UNet consumes (joint training,
        srdiff_diffusion.py:212-214), for the forward pass whose activations this plan still holds.  Every convolution: weight / bias gradient (tcgen05 where the
        shapes allow, SIMT otherwise) + data gradient = the forward kernel on transposed flipped weights; the dense blocks run on ONE
        (nf + 4 gc)-channel gradient buffer that mirrors the forward concat buffer (growth slices accumulate in place), LeakyReLU
        masks come from the kept outputs."""
        from ... import taps as T
        e, net = self.eng, self.net
        B, h, w, nf, gc = self.B, self.h, self.w, net.nf, net.gc
        new = e.new_act
        t1, t2, t4 = T.forward_taps(3, 1, h, w), T.forward_taps(3, 1, 2 * h, 2 * w), T.forward_taps(3, 1, 4 * h, 4 * w)
        nb = len(net.RRDB_trunk)
        fgrads = {}
        if d_cond is not None:
            picked = list(range(nb + 1))[2::3]                   # feature k < nb = output of RRDB k, feature nb = fea_first + trunk
            dc = e.nchw_to_act(d_cond.to(torch.float32).contiguous(), new(B, h, w, nf * len(picked)))
            fgrads = {k: dc.slice(n_ * nf, nf) for n_, k in enumerate(picked)}
        if d_raw is not None:
            # ---- tail: conv_last <- lrelu <- HRconv <- lrelu <- upconv2 <- lrelu <- upconv1 (RRDBNet.py:49-55) ----
            g_last = e.nchw_to_act(d_raw.to(torch.float32).contiguous(), new(B, 4 * h, 4 * w, net.out_nc))
            self._wgrad(net.conv_last, self.hr, g_last, t4, force_simt=True)
            g_hr = e.conv(g_last, self._dgrad_pack(net.conv_last), new(B, 4 * h, 4 * w, nf), bias=False, force_simt=True)
            self._mask(self.hr, g_hr)
            self._wgrad(net.HRconv, self.up2, g_hr, t4)
            g_up2 = e.conv(g_hr, self._dgrad_pack(net.HRconv), new(B, 4 * h, 4 * w, nf), bias=False)
            self._mask(self.up2, g_up2)
            self._wgrad(net.upconv2, self.up1, g_up2, T.forward_upsample_taps(2 * h, 2 * w), up=2)
            g_up1 = e.conv(g_up2, self._dgrad_pack(net.upconv2, up=True), new(B, 2 * h, 2 * w, nf), taps=T.dgrad_upsample_taps(2 * h, 2 * w), bias=False)
            self._mask(self.up1, g_up1)
            self._wgrad(net.upconv1, self.trunk, g_up1, T.forward_upsample_taps(h, w), up=2)
            g_fea = e.conv(g_up1, self._dgrad_pack(net.upconv1, up=True), new(B, h, w, nf), taps=T.dgrad_upsample_taps(h, w), bias=False)
            if nb in fgrads:
                self._axpby(g_fea, 1.0, fgrads[nb], 1.0, g_fea)
        else:
            assert nb in fgrads, "nothing to differentiate"
            g_fea = fgrads[nb]
        # ---- fea = fea_first + trunk_conv(last RRDB output) (:46-47) ----
        last_in = self.out_last
        self._wgrad(net.trunk_conv, last_in, g_fea, t1)
        g = e.conv(g_fea, self._dgrad_pack(net.trunk_conv), new(B, h, w, nf), bias=False)       # d / d(last RRDB output)
        # ---- the trunk, last block first ----
        Dg = new(B, h, w, self.wide)
        go, g5, g_sum = new(B, h, w, nf), new(B, h, w, nf), new(B, h, w, nf)
        for i in range(nb - 1, -1, -1):
            rrdb = net.RRDB_trunk[i]
            if i in fgrads:
                self._axpby(g, 1.0, fgrads[i], 1.0, g)            # this block's output is one of the condition features
            self._axpby(g, 0.2, g, 0.0, go)                       # out = RDB3(...) * 0.2 + x  (:133): gradient into the RDB chain
            for j, rdb in ((2, rrdb.RDB3), (1, rrdb.RDB2), (0, rrdb.RDB1)):
                D = self.dense[i][j]
                self._axpby(go, 0.2, go, 0.0, g5)                 # x5 * 0.2 + x  (:111)
                self._wgrad(rdb.conv5, D, g5, t1)
                e.conv(g5, self._dgrad_pack(rdb.conv5), Dg, bias=False)
                self._axpby(Dg.slice(0, nf), 1.0, go, 1.0, Dg.slice(0, nf))
                for k in (4, 3, 2, 1):
                    conv = (rdb.conv1, rdb.conv2, rdb.conv3, rdb.conv4)[k - 1]
                    cin = nf + (k - 1) * gc
                    sl = Dg.slice(cin, gc)
                    self._mask(D.slice(cin, gc), sl)
                    self._wgrad(conv, D.slice(0, cin), sl, t1)
                    e.conv(sl, self._dgrad_pack(conv), Dg.slice(0, cin), bias=False, res=Dg.slice(0, cin))
                self._axpby(Dg.slice(0, nf), 1.0, Dg.slice(0, nf), 0.0, go)          # gradient w.r.t. this RDB's input
            self._axpby(g, 1.0, go, 1.0, g_sum)                   # direct path + RDB chain
            g, g_sum = g_sum, g
        # ---- conv_first: fea_first feeds the trunk and the skip (:42,47) ----
        self._axpby(g, 1.0, g_fea, 1.0, g)
        self._wgrad(net.conv_first, self.x0, g, t1, force_simt=True)
        e._keep.clear()


class _RRDBPretrainFn(torch.autograd.Function):
    """raw conv_last output with a grad_fn: backward() runs ``_RRDBPlan.backward`` (parameter gradients only; the input is data)."""

    @staticmethod
    def forward(ctx, anchor, net, x):
        pl = net.plan(x)
        pl.run(x)
        raw = pl.sr_image_raw()
        ctx.pl = pl
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        ctx.pl.backward(d_raw)
        return None, None, None


class _RRDBJointFn(torch.autograd.Function):
    """(raw conv_last output, cat(feas[2::3])) with grad_fns: ONE backward call receives both gradients (joint training of the
    encoder inside SRDiff, srdiff_diffusion.py:176-214)."""

    @staticmethod
    def forward(ctx, anchor, net, x):
        pl = net.plan(x)
        feas = pl.run(x)
        raw = pl.sr_image_raw()
        cond = torch.cat([f.to_nchw(pl.eng) for f in feas[2::3]], dim=1)
        ctx.pl = pl
        ctx.set_materialize_grads(False)
        return raw, cond

    @staticmethod
    def backward(ctx, d_raw, d_cond):
        if d_raw is not None or d_cond is not None:
            ctx.pl.backward(d_raw, d_cond)
        return None, None, None


class RRDBNet(nn.Module):
    def __init__(self, in_nc, out_nc, nf, nb, gc=32, precision="bf16"):
        super().__init__()
        self.in_nc, self.out_nc, self.nf, self.gc = in_nc, out_nc, nf, gc
        block = functools.partial(RRDB, nf=nf, gc=gc)
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1, bias=True)
        self.RRDB_trunk = make_layer(block, nb)
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.HRconv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.conv_last = nn.Conv2d(nf, out_nc, 3, 1, 1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)
        self.precision = precision
        self._plans = {}

    def plan(self, x):
        B, _, h, w = x.shape
        key = (B, h, w, str(x.device), self.precision)
        pl = self._plans.get(key)
        if pl is None:
            pl = _RRDBPlan(self, B, h, w, x.device, self.precision)
            self._plans[key] = pl
        pl.refresh()
        return pl

    def forward(self, x, get_fea=False):
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            # pre-training (reference pretrain.py:37-48): the SR image carries the gradient; features come out detached
            raw = _RRDBPretrainFn.apply(self.conv_first.weight, self, x)
            out = raw.clamp(0, 1) * 2 - 1
            if get_fea:
                pl = self.plan(x)
                return out, [f.to_nchw(pl.eng) for f in pl.last_feas]
            return out
        with torch.no_grad():
            return self._forward_eval(x, get_fea)

    def forward_joint(self, x):
        """(sr_image, condition cat(feas[2::3])) -- both differentiable w.r.t. the encoder's parameters."""
        raw, cond = _RRDBJointFn.apply(self.conv_first.weight, self, x)
        return raw.clamp(0, 1) * 2 - 1, cond

    def _forward_eval(self, x, get_fea=False):
        pl = self.plan(x)
        feas = pl.run(x)
        out = pl.sr_image()
        if get_fea:
            return out, [f.to_nchw(pl.eng) for f in feas]
        return out

    @torch.no_grad()
    def features_device(self, x):
        """Fast path used by SRDiffDiffusion: the 18 feature maps as device-resident NHWC Acts (no NCHW round trip, no
        SR tail -- the reference discards the SR image on this path, srdiff_diffusion.py:107-108)."""
        pl = self.plan(x)
        return pl, pl.run(x)
