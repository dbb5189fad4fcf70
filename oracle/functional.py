"""ORACLE (test infrastructure, NOT product code) -- primitive operators.

CPU restatement, in plain functional torch fp32/fp64, of the operators the reference's hot path
is made of.  Every function cites the reference lines it restates (paths relative to the
reference checkout).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; genesis_b200/ never does.

Parameters are passed as a flat dict `P` keyed by the reference's state_dict names (SURVEY.md
section 8b), so a reference checkpoint can be fed to the oracle unchanged.
"""
import math

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------- noise tape
class NoiseTape(object):
    """Deterministic source of the eps / u draws of one forward pass.

    The reference draws noise through torch.distributions.Normal.rsample (-> _standard_normal)
    and Tensor.uniform_ in a fixed order per model ("RNG parity").  The tape
    produces the same sequence from a seeded CPU generator and records it, so the reference, the
    oracle and the CUDA engine can all consume identical noise."""

    def __init__(self, seed=None, record=None):
        self.gen = None
        if seed is not None:
            self.gen = torch.Generator(device='cpu')
            self.gen.manual_seed(int(seed))
        self.record = [] if record is None else list(record)
        self.replay = record is not None
        self.pos = 0

    def _next(self, kind, shape, dtype):
        shape = tuple(int(s) for s in shape)
        if self.replay:
            k, t = self.record[self.pos]
            assert k == kind and tuple(t.shape) == shape, (k, kind, tuple(t.shape), shape)
            self.pos += 1
            return t.to(dtype)
        if kind == 'normal':
            t = torch.randn(shape, generator=self.gen, dtype=torch.float32)
        else:
            t = torch.rand(shape, generator=self.gen, dtype=torch.float32)
        self.record.append((kind, t))
        return t.to(dtype)

    def normal(self, shape, dtype=torch.float32):
        return self._next('normal', shape, dtype)

    def uniform(self, shape, dtype=torch.float32):
        return self._next('uniform', shape, dtype)

    def rewound(self):
        return NoiseTape(record=self.record)


# ----------------------------------------------------------------------------- scalar maps
def to_sigma(x):
    """modules/blocks.py:22-23 -- softplus(x + 0.5) + 1e-8."""
    return F.softplus(x + 0.5) + 1e-8


def to_var(x):
    """modules/blocks.py:25-26."""
    return to_sigma(x) ** 2


def to_prior_sigma(x):
    """modules/blocks.py:28-34 -- sigmoid(x + 4) + 1e-4."""
    return torch.sigmoid(x + 4.0) + 1e-4


def normal_log_prob(value, mu, sigma):
    """torch.distributions.Normal.log_prob as used by genesis_config.py:242-243,275-276,330-332."""
    if not torch.is_tensor(sigma):
        sigma = torch.tensor(float(sigma), dtype=value.dtype)
    return -((value - mu) ** 2) / (2.0 * sigma ** 2) - torch.log(sigma) - 0.5 * LOG_2PI


def clamp_ste(x, lo, hi):
    """modules/blocks.py:18-20 -- clamp in value, identity in gradient."""
    return x + (x.clamp(lo, hi) - x).detach()


def pixel_coords(dim, dtype=torch.float32):
    """modules/blocks.py:42-47,119-130: meshgrid 'ij' of linspace(-1,1,dim); channel 0 varies along
    rows (H), channel 1 along columns (W)."""
    lin = torch.linspace(-1, 1, dim, dtype=torch.float32).to(dtype)
    g1 = lin.view(dim, 1).expand(dim, dim)
    g2 = lin.view(1, dim).expand(dim, dim)
    return torch.stack([g1, g2], 0).unsqueeze(0)  # [1,2,dim,dim]


# ----------------------------------------------------------------------------- norms
def batch_norm(y, P, name, training, updates=None, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d as configured by third_party/sylvester/layers.py:27,36,74,83: batch statistics
    (biased variance) in training, running statistics in eval; running_var is updated with the
    UNBIASED variance.  `updates` collects the new buffer values (the oracle never mutates P)."""
    w, b = P[name + '.weight'], P[name + '.bias']
    if training:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if updates is not None:
            n = y.numel() // y.shape[1]
            rm, rv = P[name + '.running_mean'], P[name + '.running_var']
            updates[name + '.running_mean'] = ((1 - momentum) * rm + momentum * mean).detach()
            updates[name + '.running_var'] = ((1 - momentum) * rv
                                              + momentum * var * n / max(n - 1, 1)).detach()
            updates[name + '.num_batches_tracked'] = P[name + '.num_batches_tracked'] + 1
    else:
        mean, var = P[name + '.running_mean'], P[name + '.running_var']
    yhat = (y - mean.view(1, -1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1) + eps)
    return yhat * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def instance_norm(y, w, b, eps=1e-5):
    """nn.InstanceNorm2d(affine=True) (modules/blocks.py:155): per (n,c) over HxW, biased var."""
    mean = y.mean(dim=(2, 3), keepdim=True)
    var = y.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (y - mean) * torch.rsqrt(var + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def group_norm(y, groups, w, b, eps=1e-5):
    """nn.GroupNorm(8, C) (modules/blocks.py:163; genesisv2_config.py:92-98): contiguous channel
    groups, per sample, biased var, per-channel affine."""
    n, c = y.shape[:2]
    yg = y.reshape(n, groups, -1)
    mean = yg.mean(dim=2, keepdim=True)
    var = yg.var(dim=2, unbiased=False, keepdim=True)
    yh = ((yg - mean) * torch.rsqrt(var + eps)).reshape(y.shape)
    return yh * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm(128) in z_head (genesisv2_config.py:83)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * w + b


# ----------------------------------------------------------------------------- layers
def linear(x, P, name):
    return x @ P[name + '.weight'].t() + P[name + '.bias']


def gated_conv(x, P, name, stride, pad, norm, training, updates, transpose=False):
    """third_party/sylvester/layers.py:42-54 (GatedConv2d) and :88-101 (GatedConvTranspose2d):
    one conv to 2C channels, chunk -> h (first half, no activation), g (second half);
    optional norm on each; out = h * sigmoid(g).  Transposed layers use output_padding=stride-1
    (VAE.py:31-32)."""
    w, b = P[name + '.conv.weight'], P[name + '.conv.bias']
    if transpose:
        y = F.conv_transpose2d(x, w, b, stride=stride, padding=pad, output_padding=stride - 1)
    else:
        y = F.conv2d(x, w, b, stride=stride, padding=pad)
    h, g = torch.chunk(y, 2, dim=1)
    if norm == 'bn':
        h = batch_norm(h, P, name + '.h_norm', training, updates)
        g = batch_norm(g, P, name + '.g_norm', training, updates)
    elif norm == 'in':
        h = instance_norm(h, P[name + '.h_norm.weight'], P[name + '.h_norm.bias'])
        g = instance_norm(g, P[name + '.g_norm.weight'], P[name + '.g_norm.bias'])
    return h * torch.sigmoid(g)


def sylvester_strides(img_size):
    """third_party/sylvester/VAE.py:56-69."""
    table = {32: (8, [1, 2, 1, 2, 1]), 64: (16, [1, 2, 1, 2, 1]),
             128: (16, [2, 2, 2, 1, 1]), 256: (16, [2, 2, 2, 2, 1])}
    if img_size not in table:
        raise ValueError('Invalid input size.')
    return table[img_size]


def sylvester_q_z_nn(x, P, prefix, img_size, norm, training, updates):
    """VAE.create_encoder / build_gc_encoder (VAE.py:18-24,92-110): five gated 5x5 convs (p=2) with
    norm, then one un-normed gated conv with a full-map kernel -> [B,256,1,1]."""
    _, strides = sylvester_strides(img_size)
    h = x
    for i, s in enumerate(strides):
        h = gated_conv(h, P, '%s.%d' % (prefix, i), s, 2, norm, training, updates)
    return gated_conv(h, P, '%s.%d' % (prefix, len(strides)), 1, 0, None, training, updates)


def sylvester_decode(z, P, prefix, img_size, norm, training, updates, nn_prefix=None, mean_prefix=None):
    """VAE.decode (VAE.py:143-153) with build_gc_decoder (VAE.py:27-33): un-normed gated
    conv-transpose 1x1 -> kz x kz, five gated 5x5 conv-transposes (p=2, op=s-1) with norm, 1x1 conv.
    nn_prefix / mean_prefix name the stack and the 1x1 conv when they are not `<prefix>.p_x_nn` / `<prefix>.p_x_mean`
    (the comp_symmetric component decoder, genesis_config.py:110-120)."""
    _, strides = sylvester_strides(img_size)
    strides = list(reversed(strides))
    nn_prefix = nn_prefix or prefix + '.p_x_nn'
    mean_prefix = mean_prefix or prefix + '.p_x_mean'
    h = z.view(z.shape[0], -1, 1, 1)
    h = gated_conv(h, P, nn_prefix + '.0', 1, 0, None, training, updates, transpose=True)
    for i, s in enumerate(strides):
        h = gated_conv(h, P, '%s.%d' % (nn_prefix, i + 1), s, 2, norm, training, updates, transpose=True)
    return F.conv2d(h, P[mean_prefix + '.weight'], P[mean_prefix + '.bias'])


def lstm_cell(x, state, P, name):
    """One step of a single-layer torch nn.LSTM (gate order i,f,g,o; two bias vectors), as stepped by
    modules/attention.py:94-96 and run over a sequence by genesis_config.py:301-305."""
    hid = P[name + '.weight_hh_l0'].shape[1]
    if state is None:
        h = x.new_zeros(x.shape[0], hid)
        c = x.new_zeros(x.shape[0], hid)
    else:
        h, c = state
    gates = (x @ P[name + '.weight_ih_l0'].t() + P[name + '.bias_ih_l0']
             + h @ P[name + '.weight_hh_l0'].t() + P[name + '.bias_hh_l0'])
    i, f, g, o = torch.chunk(gates, 4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, (h, c)


def act_fn(name):
    return {'elu': F.elu, 'relu': F.relu}[name]


def monet_comp_encoder(x4, P, prefix, act):
    """modules/encoders.py:31-40: 4x [Conv3x3 s2 p1 + act], Flatten (NCHW order), Linear+act, Linear."""
    h = x4
    for i in (0, 2, 4, 6):
        n = '%s.module.%d' % (prefix, i)
        h = act(F.conv2d(h, P[n + '.weight'], P[n + '.bias'], stride=2, padding=1))
    h = h.flatten(1)
    h = act(linear(h, P, prefix + '.module.9'))
    return linear(h, P, prefix + '.module.11')


def broadcast_decoder(z, P, prefix, img_size, num_layers, act):
    """modules/decoders.py:21-35 + blocks.BroadcastLayer/PixelCoords (blocks.py:104-130): tile z over a
    (D+2L)^2 grid, append the two coordinate channels, L VALID 3x3 convs + act, one 1x1 conv."""
    d = img_size + 2 * num_layers
    n = z.shape[0]
    h = z.view(n, -1, 1, 1).expand(-1, -1, d, d)
    h = torch.cat([h, pixel_coords(d, z.dtype).expand(n, -1, -1, -1)], dim=1)
    for i in range(num_layers):
        name = '%s.seq.%d' % (prefix, 1 + 2 * i)
        h = act(F.conv2d(h, P[name + '.weight'], P[name + '.bias']))
    name = '%s.seq.%d' % (prefix, 1 + 2 * num_layers)
    return F.conv2d(h, P[name + '.weight'], P[name + '.bias'])


def stick_breaking(logits_k, append_scope=True):
    """modules/attention.py:40-50,114-130: log m_k = log s_k + logsigmoid(a_k);
    log s_{k+1} = log s_k + logsigmoid(-a_k); log s_0 = 0; the final scope is appended as the last
    mask."""
    log_s = [torch.zeros_like(logits_k[0])]
    log_m = []
    for a in logits_k:
        log_m.append(log_s[-1] + F.logsigmoid(a))
        log_s.append(log_s[-1] + F.logsigmoid(-a))
    if append_scope:
        log_m.append(log_s[-1])
    return log_m, log_s


def mixture_nll(x, log_m_k, x_r_k, std):
    """Genesis.x_loss (models/genesis_config.py:273-286): err_b = -sum_{c,h,w} log sum_k
    exp(log m_k + log N(x; x_r_k, std_k)).  `std` is a python float or a [K] tensor (the reference's
    `std` buffer is [1,1,1,1,K])."""
    xr = torch.stack(x_r_k, dim=4)
    if torch.is_tensor(std):
        std = std.reshape(1, 1, 1, 1, -1).to(x.dtype)
    lp = normal_log_prob(x.unsqueeze(4), xr, std)
    log_mx = torch.stack(log_m_k, dim=4) + lp
    return -torch.log(log_mx.exp().sum(dim=4)).sum(dim=(1, 2, 3))


def autoreg_prior(z_k, P, lstm='prior_lstm', lin='prior_linear'):
    """Genesis.mask_latent_loss prior part (models/genesis_config.py:297-320): teacher-forced LSTM over
    z_0..z_{K-2} -> Linear -> (tanh mean, to_prior_sigma); the first step's prior is N(0,1)."""
    pmu, psig = [None], [None]
    state = None
    for z in z_k[:-1]:
        out, state = lstm_cell(z, state, P, lstm)
        lo = linear(out, P, lin)
        a, b = torch.chunk(lo, 2, dim=1)
        pmu.append(torch.tanh(a))
        psig.append(to_prior_sigma(b))
    return pmu, psig


def mc_kl(z, mu, sigma, pmu=None, psigma=None):
    """Monte-Carlo KL: sum_d [log q(z) - log p(z)] (genesis_config.py:328-336; utils/misc.py:254-255);
    p = N(0,1) when pmu is None."""
    log_q = normal_log_prob(z, mu, sigma).sum(dim=1)
    if pmu is None:
        log_p = (-0.5 * z ** 2 - 0.5 * LOG_2PI).sum(dim=1)
    else:
        log_p = normal_log_prob(z, pmu, psigma).sum(dim=1)
    return log_q - log_p


def unet(x, P, prefix, num_blocks, norm, groups=8):
    """modules/unet.py:69-90 (without final_conv): down blocks (Conv3x3 p1 no bias + IN/GN + ReLU),
    skip taken before nearest x0.5; Flatten-MLP bottleneck; up blocks on cat([x_up, skip]) with
    nearest x2 between blocks."""
    def block(h, name):
        h = F.conv2d(h, P[name + '.0.weight'], None, padding=1)
        if norm == 'in':
            h = instance_norm(h, P[name + '.1.weight'], P[name + '.1.bias'])
        else:
            h = group_norm(h, groups, P[name + '.1.weight'], P[name + '.1.bias'])
        return F.relu(h)

    skip = []
    h = x
    for i in range(num_blocks):
        h = block(h, '%s.down.%d' % (prefix, i))
        skip.append(h)
        if i < num_blocks - 1:
            h = h[:, :, ::2, ::2]          # F.interpolate(scale_factor=0.5, 'nearest')
    n, c, f, _ = h.shape
    u = h.flatten(1)
    for j in (1, 3, 5):
        u = F.relu(linear(u, P, '%s.mlp.%d' % (prefix, j)))
    u = u.view(n, c, f, f)
    for i in range(num_blocks):
        u = block(torch.cat([u, skip[-1 - i]], dim=1), '%s.up.%d' % (prefix, i))
        if i < num_blocks - 1:
            u = u.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)   # nearest x2
    return u


def categorical_kl(q_probs, p_probs):
    """torch.distributions kl_divergence(Categorical(q), Categorical(p)) as called by
    models/monet_config.py:167-170: probabilities are renormalised, logits = log(clamp(probs, eps,
    1-eps))."""
    eps = torch.finfo(q_probs.dtype).eps
    q = q_probs / q_probs.sum(-1, keepdim=True)
    p = p_probs / p_probs.sum(-1, keepdim=True)
    lq = torch.log(q.clamp(eps, 1 - eps))
    lp = torch.log(p.clamp(eps, 1 - eps))
    return (q * (lq - lp)).sum(-1)
