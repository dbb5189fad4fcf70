"""Parameter holders and shared forward building blocks of the engine.

The holders create parameters with the SAME names, shapes, default initialisers and construction ORDER as
the reference modules (so `torch.manual_seed(s)` followed by construction gives bit-identical parameters,
and reference checkpoints load with `load_state_dict`).  torch.nn layer classes are used only as parameter
containers; their forward() is never called -- all compute goes through genesis_b200.ops (CUDA kernels)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


# ------------------------------------------------------------------------------------------- holders
class GatedLayer(nn.Module):
    """Holder for sylvester GatedConv2d / GatedConvTranspose2d (reference third_party/sylvester/layers.py:11-41,
    57-87): `conv` to 2*cout channels, optional `h_norm` / `g_norm`."""

    def __init__(self, cin, cout, k, stride, pad, transposed=False, out_pad=0, norm=None):
        super().__init__()
        self.k, self.stride, self.pad, self.transposed, self.norm = k, stride, pad, transposed, norm
        if transposed:
            self.conv = nn.ConvTranspose2d(cin, 2 * cout, k, stride, pad, out_pad)
        else:
            self.conv = nn.Conv2d(cin, 2 * cout, k, stride, pad)
        self.h_norm, self.g_norm = None, None
        if norm == 'in':
            self.h_norm = nn.InstanceNorm2d(cout, affine=True)
            self.g_norm = nn.InstanceNorm2d(cout, affine=True)
        elif norm == 'bn':
            self.h_norm = nn.BatchNorm2d(cout)
            self.g_norm = nn.BatchNorm2d(cout)
        elif norm is not None and norm != 'none':
            raise ValueError("Normalisation option not recognised.")


class ToVarHolder(nn.Module):
    """Placeholder for blocks.ToVar (parameter-free) so that `q_z_var.0` keeps its reference name."""


def sylvester_strides(img_size):
    """reference third_party/sylvester/VAE.py:56-69."""
    table = {32: (8, [1, 2, 1, 2, 1]), 64: (16, [1, 2, 1, 2, 1]),
             128: (16, [2, 2, 2, 1, 1]), 256: (16, [2, 2, 2, 2, 1])}
    if img_size not in table:
        raise ValueError('Invalid input size.')
    return table[img_size]


class SylvesterVAE(nn.Module):
    """Holder for third_party/sylvester/VAE.py:36-124 (gated conv encoder / decoder of the attention core)."""

    def __init__(self, z_size, input_size, nout, enc_norm=None, dec_norm=None):
        super().__init__()
        self.z_size, self.input_size = z_size, input_size
        self.nout = nout if nout is not None else input_size[0]
        self.enc_norm, self.dec_norm = enc_norm, dec_norm
        self.last_kernel_size, self.strides = sylvester_strides(input_size[1])
        self.q_z_nn_output_dim = 256
        cin, cout = [input_size[0], 32, 32, 64, 64], [32, 32, 64, 64, 64]
        enc = [GatedLayer(i, o, 5, s, 2, norm=enc_norm) for i, o, s in zip(cin, cout, self.strides)]
        enc.append(GatedLayer(cout[-1], self.q_z_nn_output_dim, self.last_kernel_size, 1, 0))
        self.q_z_nn = nn.Sequential(*enc)
        self.q_z_mean = nn.Linear(256, z_size)
        self.q_z_var = nn.Sequential(nn.Linear(256, z_size), ToVarHolder())
        cin, cout = [64, 64, 32, 32, 32], [64, 32, 32, 32, 32]
        rstrides = list(reversed(self.strides))
        dec = [GatedLayer(z_size, cin[0], self.last_kernel_size, 1, 0, transposed=True)]
        dec += [GatedLayer(i, o, 5, s, 2, transposed=True, out_pad=s - 1, norm=dec_norm)
                for i, o, s in zip(cin, cout, rstrides)]
        self.p_x_nn = nn.Sequential(*dec)
        self.p_x_mean = nn.Conv2d(cout[-1], self.nout, 1, 1, 0)


class CompEncoderHolder(nn.Module):
    """Holder for modules/encoders.py:22-37 (MONetCompEncoder); parameters live at module.{0,2,4,6,9,11}."""

    def __init__(self, cfg):
        super().__init__()
        nin = cfg.input_channels if hasattr(cfg, 'input_channels') else 3
        c = cfg.comp_enc_channels
        self.ldim = cfg.comp_ldim
        nin_mlp = 2 * c * (cfg.img_size // 16) ** 2
        nhid = max(256, 2 * self.ldim)
        idt = nn.Identity
        self.module = nn.Sequential(
            nn.Conv2d(nin + 1, c, 3, 2, 1), idt(), nn.Conv2d(c, c, 3, 2, 1), idt(),
            nn.Conv2d(c, 2 * c, 3, 2, 1), idt(), nn.Conv2d(2 * c, 2 * c, 3, 2, 1), idt(),
            idt(), nn.Linear(nin_mlp, nhid), idt(), nn.Linear(nhid, 2 * self.ldim))


class BroadcastDecoderHolder(nn.Module):
    """Holder for modules/decoders.py:21-32; parameters live at seq.{1,3,..}."""

    def __init__(self, in_chnls, out_chnls, h_chnls, num_layers, img_dim):
        super().__init__()
        self.num_layers, self.img_dim, self.in_chnls = num_layers, img_dim, in_chnls
        mods = [nn.Identity(), nn.Conv2d(in_chnls + 2, h_chnls, 3), nn.Identity()]
        for _ in range(num_layers - 1):
            mods.extend([nn.Conv2d(h_chnls, h_chnls, 3), nn.Identity()])
        mods.append(nn.Conv2d(h_chnls, out_chnls, 1))
        self.seq = nn.Sequential(*mods)


class ComponentVAEHolder(nn.Module):
    """Holder for modules/component_vae.py:27-43."""

    def __init__(self, nout, cfg):
        super().__init__()
        self.ldim = cfg.comp_ldim
        self.pixel_bound = cfg.pixel_bound
        self.nout = nout
        self.encoder_module = CompEncoderHolder(cfg)
        self.decoder_module = BroadcastDecoderHolder(self.ldim, nout, cfg.comp_dec_channels,
                                                     cfg.comp_dec_layers, cfg.img_size)


class _GatedCore(object):
    """Adapter that lets sylvester_encode / sylvester_decode run on free-standing gated-conv stacks (the comp_symmetric
    component VAE): q_z_nn / p_x_nn are nn.Sequential of GatedLayer, p_x_mean a 1x1 Conv2d."""

    def __init__(self, q_z_nn=None, p_x_nn=None, p_x_mean=None):
        self.q_z_nn, self.p_x_nn, self.p_x_mean = q_z_nn, p_x_nn, p_x_mean


def make_symmetric_component_vae(comp_vae, nin, cfg, last_kernel):
    """Replace the MONet-style encoder / broadcast decoder of a ComponentVAEHolder by the gated conv stacks of reference
    genesis_config.py:101-120 (comp_symmetric=True).  Called AFTER the default modules were constructed, as the reference
    does, so the seeded initialisation consumes the generator in the same order.  state_dict names:
    encoder_module.0.{0-5}.(conv|h_norm|g_norm), decoder_module.1.{0-5}.*, decoder_module.2.(weight|bias)."""
    strides = [1, 2, 1, 2, 1]
    cin, cout = [nin + 1, 32, 32, 64, 64], [32, 32, 64, 64, 64]
    enc = [GatedLayer(i, o, 5, s, 2, norm=cfg.enc_norm) for i, o, s in zip(cin, cout, strides)]
    enc.append(GatedLayer(cout[-1], 2 * cfg.comp_ldim, last_kernel, 1, 0))
    comp_vae.encoder_module = nn.Sequential(nn.Sequential(*enc), nn.Identity())
    cin, cout = [64, 64, 32, 32, 32], [64, 32, 32, 32, 32]
    dec = [GatedLayer(cfg.comp_ldim, cin[0], last_kernel, 1, 0, transposed=True)]
    dec += [GatedLayer(i, o, 5, s, 2, transposed=True, out_pad=s - 1, norm=cfg.dec_norm) for i, o, s in zip(cin, cout, strides)]
    comp_vae.decoder_module = nn.Sequential(nn.Identity(), nn.Sequential(*dec), nn.Conv2d(32, nin, 1))
    return comp_vae


def symmetric_comp_encode(comp_vae, packed, training):
    """packed NHWC [K*B,H,W,cp] (channel 0 = log-mask, 1..3 = image) -> [K*B, 2*ldim]."""
    return sylvester_encode(_GatedCore(q_z_nn=comp_vae.encoder_module[0]), packed, training)


def symmetric_comp_decode(comp_vae, z, training, nsig):
    """z [K*B, ldim] -> NCHW [K*B, nin, H, W] with the pixel-bound sigmoid on the first nsig channels."""
    dm = comp_vae.decoder_module
    return sylvester_decode(_GatedCore(p_x_nn=dm[1], p_x_mean=dm[2]), z, training, nsig)


def _conv_block(nin, nout, norm):
    """modules/blocks.py:144-165: Conv3x3 (no bias with a norm) + IN(affine) / GN(8) [+ ReLU]."""
    if norm == 'in':
        return nn.Sequential(nn.Conv2d(nin, nout, 3, 1, 1, bias=False), nn.InstanceNorm2d(nout, affine=True))
    if norm == 'gn':
        return nn.Sequential(nn.Conv2d(nin, nout, 3, 1, 1, bias=False), nn.GroupNorm(8, nout))
    return nn.Sequential(nn.Conv2d(nin, nout, 3, 1, 1))


class UNetHolder(nn.Module):
    """Holder for modules/unet.py:23-67."""

    def __init__(self, num_blocks, img_size=64, filter_start=32, in_chnls=4, out_chnls=1, norm='in'):
        super().__init__()
        c = filter_start
        self.norm, self.num_blocks = norm, num_blocks
        if num_blocks == 4:
            enc_in, enc_out = [in_chnls, c, 2 * c, 2 * c], [c, 2 * c, 2 * c, 2 * c]
            dec_in, dec_out = [4 * c, 4 * c, 4 * c, 2 * c], [2 * c, 2 * c, c, c]
        elif num_blocks == 5:
            enc_in, enc_out = [in_chnls, c, c, 2 * c, 2 * c], [c, c, 2 * c, 2 * c, 2 * c]
            dec_in, dec_out = [4 * c, 4 * c, 4 * c, 2 * c, 2 * c], [2 * c, 2 * c, c, c, c]
        elif num_blocks == 6:
            enc_in, enc_out = [in_chnls, c, c, c, 2 * c, 2 * c], [c, c, c, 2 * c, 2 * c, 2 * c]
            dec_in, dec_out = [4 * c, 4 * c, 4 * c, 2 * c, 2 * c, 2 * c], [2 * c, 2 * c, c, c, c, c]
        else:
            raise ValueError('unsupported number of UNet blocks')
        self.down = nn.ModuleList([_conv_block(i, o, norm) for i, o in zip(enc_in, enc_out)])
        self.up = nn.ModuleList([_conv_block(i, o, norm) for i, o in zip(dec_in, dec_out)])
        self.featuremap_size = img_size // 2 ** (num_blocks - 1)
        f = 2 * c * self.featuremap_size ** 2
        self.mlp = nn.Sequential(nn.Identity(), nn.Linear(f, 128), nn.Identity(), nn.Linear(128, 128),
                                 nn.Identity(), nn.Linear(128, f), nn.Identity())
        self.final_conv = nn.Conv2d(c, out_chnls, 1)
        self.out_chnls = out_chnls


# ------------------------------------------------------------------------------------------- scalar maps
def to_sigma(x):
    """reference modules/blocks.py:22-23."""
    return F.softplus(x + 0.5) + 1e-8


def to_prior_sigma(x):
    """reference modules/blocks.py:28-34."""
    return torch.sigmoid(x + 4.0) + 1e-4


LOG_2PI = math.log(2.0 * math.pi)


def normal_log_prob(z, mu, sigma):
    return -((z - mu) ** 2) / (2.0 * sigma ** 2) - torch.log(sigma) - 0.5 * LOG_2PI


def mc_kl(z, mu, sigma, pmu=None, psigma=None):
    """sum_d [log q(z) - log p(z)]; p = N(0,1) when pmu is None (reference genesis_config.py:328-336)."""
    if ops.fused_latent() and z.is_cuda:
        return ops.mc_kl(z, mu, sigma, pmu, psigma)
    log_q = normal_log_prob(z, mu, sigma).sum(dim=1)
    if pmu is None:
        log_p = (-0.5 * z ** 2 - 0.5 * LOG_2PI).sum(dim=1)
    else:
        log_p = normal_log_prob(z, pmu, psigma).sum(dim=1)
    return log_q - log_p


# ------------------------------------------------------------------------------------------- forward blocks
def pad_cin(w, cin):
    """zero-pad a conv weight [Co,Ci,R,S] along Ci (matches activations zero-padded to `cin` channels)."""
    return w if w.shape[1] == cin else F.pad(w, (0, 0, 0, 0, 0, cin - w.shape[1]))


def gated_forward(layer, x, training):
    """GatedConv2d / GatedConvTranspose2d forward on NHWC x (reference layers.py:42-54, 88-101)."""
    conv = layer.conv
    # the conv's bias gradient (column sums of the gradient w.r.t. y) is accumulated by the gate / norm backward pass
    if layer.transposed:
        y = ops.conv_transpose2d(x, conv.weight, conv.bias, layer.stride, layer.pad, bias_grad=False)
    else:
        y = ops.conv2d(x, conv.weight, conv.bias, layer.stride, layer.pad, bias_grad=False)
    return gate_norm(layer, y, training, conv_bias=conv.bias)


# BatchNorm's num_batches_tracked counters: one multi-tensor add per encoder / decoder stack instead of one tiny kernel per
# counter (20 launches per GENESIS step).  A layer used twice before the flush appears twice and is incremented twice.
_NBT_PENDING = []


def flush_batch_counters():
    if _NBT_PENDING:
        torch._foreach_add_(list(_NBT_PENDING), 1)
        del _NBT_PENDING[:]


def gate_norm(layer, y, training, conv_bias=None):
    hn, gn = layer.h_norm, layer.g_norm
    if hn is None:
        return ops.norm_post(y, mode=ops.NORM_NONE, post=ops.POST_GATE, conv_bias=conv_bias)
    if layer.norm == 'bn':
        out = ops.norm_post(y, hn.weight, hn.bias, gn.weight, gn.bias, hn.running_mean, hn.running_var,
                            gn.running_mean, gn.running_var, mode=ops.NORM_BATCH, post=ops.POST_GATE,
                            training=training, eps=hn.eps, momentum=hn.momentum, conv_bias=conv_bias)
        if training:
            _NBT_PENDING.extend((hn.num_batches_tracked, gn.num_batches_tracked))
        return out
    return ops.norm_post(y, hn.weight, hn.bias, gn.weight, gn.bias, mode=ops.NORM_INSTANCE,
                         post=ops.POST_GATE, eps=hn.eps, conv_bias=conv_bias)


def sylvester_encode(core, x_nhwc, training):
    """core.q_z_nn(x) -> [B,256] (reference VAE.py:92-110; attention.py:85-87).  The last, full-map gated
    conv is a GEMM over the NHWC-flattened feature map."""
    h = x_nhwc
    n_layers = len(core.q_z_nn)
    for i in range(n_layers - 1):
        h = gated_forward(core.q_z_nn[i], h, training)
    last = core.q_z_nn[n_layers - 1]
    B = h.shape[0]
    w = last.conv.weight                                   # [512, 64, k, k]
    wm = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)     # columns in (h, w, c) order
    y = ops.linear(h.reshape(B, -1), wm, last.conv.bias)
    out = gate_norm(last, y.view(B, 1, 1, -1), training).view(B, -1)
    flush_batch_counters()
    return out


def sylvester_decode(core, z, training, nsig=0):
    """core.decode(z) (reference VAE.py:143-153) -> NCHW [N, nout, H, W] (mask logits for GENESIS; the BaselineVAE applies the
    pixel-bound sigmoid to its `nsig` image channels in the same kernel)."""
    N = z.shape[0]
    first = core.p_x_nn[0]
    w = first.conv.weight                                  # [z, 2C, k, k]
    k = w.shape[2]
    wm = w.permute(2, 3, 1, 0).reshape(-1, w.shape[0])     # rows in (h, w, c) order
    bias = first.conv.bias.repeat(k * k)
    y = ops.linear(z, wm, bias).view(N, k, k, w.shape[1])
    h = gate_norm(first, y, training)
    for i in range(1, len(core.p_x_nn)):
        h = gated_forward(core.p_x_nn[i], h, training)
    flush_batch_counters()
    return ops.out1x1(h, core.p_x_mean.weight, core.p_x_mean.bias, nsig)


def lstm_step(x, state, lstm):
    """One step of nn.LSTM (1 layer; gates i,f,g,o) with our GEMM; x [B,in]."""
    gates = ops.linear(x, lstm.weight_ih_l0, lstm.bias_ih_l0)
    if state is None:
        # h_0 = 0.  The recurrent term still goes through ops.linear (it returns exactly bias_hh) so that EVERY use of
        # bias_hh_l0 / weight_hh_l0 takes the same gradient path: in direct-gradient mode the kernels accumulate into
        # param.grad on the gradient side stream, and a second, autograd-side `grad +=` of the same tensor on the main
        # stream (what `gates + lstm.bias_hh_l0` produced) could interleave with them.
        h_prev = x.new_zeros(x.shape[0], lstm.weight_hh_l0.shape[1])
        c_prev = None
    else:
        h_prev, c_prev = state
    gh = ops.linear(h_prev, lstm.weight_hh_l0, lstm.bias_hh_l0)
    if ops.fused_latent():
        h, c = ops.lstm_cell(gates, gh, c_prev)
        return h, (h, c)
    gates = gates + gh
    i, f, g, o = torch.chunk(gates, 4, dim=1)
    c = torch.sigmoid(i) * torch.tanh(g)
    if c_prev is not None:
        c = c + torch.sigmoid(f) * c_prev
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, (h, c)


def autoreg_prior(z_k, lstm, lin):
    """Teacher-forced autoregressive prior (reference genesis_config.py:297-320).  Returns lists of length
    K-1 (prior of steps 1..K-1)."""
    pmu, psig = [], []
    state = None
    for z in z_k[:-1]:
        out, state = lstm_step(z, state, lstm)
        lo = ops.linear(out, lin.weight, lin.bias)
        m, s_ = prior_head(lo)
        pmu.append(m)
        psig.append(s_)
    return pmu, psig


def prior_head(lo, use_tanh=True):
    """lo [B,2D] -> (tanh(lo[:, :D]), to_prior_sigma(lo[:, D:])) (reference genesis_config.py:306-314, 234-239)."""
    if ops.fused_latent():
        return ops.prior_head(lo, use_tanh)
    a, b = torch.chunk(lo, 2, dim=1)
    return (torch.tanh(a) if use_tanh else a), to_prior_sigma(b)


def gauss_head(lo, eps):
    """lo [B,2D] = (mu | raw) -> z = mu + to_sigma(raw) * eps, mu, sigma (reference attention.py:98-103, component_vae.py:64-69)."""
    if ops.fused_latent():
        return ops.gauss_head(lo, eps)
    mu, raw = torch.chunk(lo, 2, dim=1)
    sigma = to_sigma(raw)
    return mu + sigma * eps, mu, sigma


_COORDS = {}


def coords_nhwc(dim, device):
    """[1,dim,dim,2] coordinate planes, channel 0 along rows (reference blocks.py:119-130)."""
    key = (dim, str(device))
    if key not in _COORDS:
        lin = torch.linspace(-1, 1, dim)
        g1 = lin.view(dim, 1).expand(dim, dim)
        g2 = lin.view(1, dim).expand(dim, dim)
        _COORDS[key] = torch.stack([g1, g2], dim=2).unsqueeze(0).contiguous().to(device)
    return _COORDS[key]


def comp_encode(enc, packed, act):
    """MONetCompEncoder on the packed NHWC4 input (reference encoders.py:31-40) -> [N, 2*ldim]."""
    m = enc.module
    h = packed
    for i in (0, 2, 4, 6):
        h = ops.conv2d(h, m[i].weight, m[i].bias, 2, 1, act)
    N, fh, fw, c = h.shape
    w = m[9].weight                                        # [256, c*fh*fw] in NCHW flatten order
    wm = w.view(w.shape[0], c, fh, fw).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    h = ops.linear(h.reshape(N, -1), wm, m[9].bias, act)
    return ops.linear(h, m[11].weight, m[11].bias)


def broadcast_decode(dec, z, act, nsig, head_act=None):
    """BroadcastDecoder (reference decoders.py:21-35) without materialising the broadcast: the first VALID
    3x3 conv over [z tiled | coords] splits into a per-sample vector (sum of the z-taps) plus a
    sample-independent coordinate map.  Returns NCHW [N, nout, D, D]; with `head_act` the final 1x1
    conv is a feature layer instead (BaselineVAE broadcast decoder, vae_config.py:53-60): NHWC [N, D, D, nout] after head_act."""
    L, D = dec.num_layers, dec.img_dim
    d = D + 2 * L
    c1 = dec.seq[1]
    ld = dec.in_chnls
    wz = c1.weight[:, :ld].sum(dim=(2, 3))                                  # [h, ldim]
    zb = ops.linear(z, wz, c1.bias)                                          # [N, h]
    cmap = ops.conv2d(coords_nhwc(d, z.device), c1.weight[:, ld:].contiguous(), None, 1, 0)   # [1,d-2,d-2,h]
    hc = cmap.shape[3]
    h = ops.bcast_add_act(zb, cmap.view(-1, hc), act).view(z.shape[0], d - 2, d - 2, hc)
    for i in range(1, L):
        c = dec.seq[1 + 2 * i]
        h = ops.conv2d(h, c.weight, c.bias, 1, 0, act)
    last = dec.seq[1 + 2 * L]
    if head_act is not None:
        return ops.conv2d(h, last.weight, last.bias, 1, 0, head_act)
    return ops.out1x1(h, last.weight, last.bias, nsig)


# ------------------------------------------------------------------------------------------- UNet
def conv_norm_relu(block, h, norm):
    """ConvINReLU / ConvGNReLU (reference modules/blocks.py:151-165) on NHWC h: 3x3 p1 conv without bias, then the
    per-sample norm fused with ReLU."""
    conv, nrm = block[0], block[1]
    y = ops.conv2d(h, conv.weight, None, 1, 1)
    if norm == 'in':
        return ops.norm_post(y, nrm.weight, nrm.bias, mode=ops.NORM_INSTANCE, post=ops.POST_RELU, eps=nrm.eps)
    return ops.norm_post(y, nrm.weight, nrm.bias, mode=ops.NORM_GROUP, groups=nrm.num_groups, post=ops.POST_RELU,
                         eps=nrm.eps)


def unet_forward(unet, h):
    """UNet.forward without final_conv (reference modules/unet.py:69-90) on NHWC h (channels may be zero-padded).
    Skips are taken before the nearest x0.5; the bottleneck MLP sees the NCHW flattening of the reference (its
    weights are re-indexed to the NHWC order instead of transposing the activations)."""
    nb, norm = unet.num_blocks, unet.norm
    skip = []
    for i in range(nb):
        h = conv_norm_relu(unet.down[i], h, norm)
        skip.append(h)
        if i < nb - 1:
            h = ops.down2(h)
    N, f, _, c = h.shape
    m = unet.mlp
    w1 = m[1].weight.view(-1, c, f, f).permute(0, 2, 3, 1).reshape(m[1].weight.shape[0], -1)
    u = ops.linear(h.reshape(N, -1), w1, m[1].bias, 'relu')
    u = ops.linear(u, m[3].weight, m[3].bias, 'relu')
    w5 = m[5].weight.view(c, f, f, -1).permute(1, 2, 0, 3).reshape(c * f * f, -1)
    b5 = m[5].bias.view(c, f, f).permute(1, 2, 0).reshape(-1)
    u = ops.linear(u, w5, b5, 'relu').view(N, f, f, c)
    for i in range(nb):
        u = conv_norm_relu(unet.up[i], torch.cat([u, skip[-1 - i]], dim=3), norm)
        if i < nb - 1:
            u = ops.up2(u)
    return u


def layer_norm(x, ln):
    """nn.LayerNorm over the last dim of x [N,C] = a one-group, one-pixel GroupNorm."""
    N, C = x.shape
    return ops.norm_post(x.view(N, 1, 1, C), ln.weight, ln.bias, mode=ops.NORM_GROUP, groups=1, post=ops.POST_NONE,
                         eps=ln.eps).view(N, C)
