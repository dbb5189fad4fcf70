"""ORACLE (test infrastructure, NOT product code) -- whole-model forward passes.

Functional CPU restatement of Genesis.forward, GenesisV2.forward and MONet.forward.  Each takes
the reference's state_dict `P` (name -> tensor; tensors that require grad are differentiated by
torch autograd), an image batch and a NoiseTape, and returns a dict with the same quantities the
reference returns (recon, losses, stats, att_stats, comp_stats) plus `bn_updates`.

Pinned against the real reference (imported from /root/reference with shims) by
oracle/make_golden.py -> tests/golden/*.npz; see tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F

from . import functional as O


class Cfg(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    __setattr__ = dict.__setitem__


GENESIS_DEFAULTS = dict(  # models/genesis_config.py:33-52
    two_stage=True, autoreg_prior=True, comp_prior=True, attention_latents=64, enc_norm='bn',
    dec_norm='bn', comp_enc_channels=32, comp_ldim=16, comp_dec_channels=32, comp_dec_layers=4,
    comp_symmetric=False, pixel_bound=True, pixel_std1=0.7, pixel_std2=0.7, montecarlo_kl=True)
GENESISV2_DEFAULTS = dict(  # models/genesisv2_config.py:35-42 (+ the genesis/monet flags it imports)
    GENESIS_DEFAULTS, feat_dim=64, kernel='gaussian', semiconv=True, dynamic_K=False,
    klm_loss=False, detach_mr_in_klm=True, filter_start=32, prior_mode='softmax')
MONET_DEFAULTS = dict(GENESIS_DEFAULTS, filter_start=32, prior_mode='softmax')  # monet_config.py:36-37


def make_cfg(model, **over):
    base = {'genesis': GENESIS_DEFAULTS, 'genesisv2': GENESISV2_DEFAULTS, 'monet': MONET_DEFAULTS, 'vae': VAE_DEFAULTS}[model]
    cfg = Cfg(base, debug=False, multi_gpu=False, img_size=64, K_steps=5)
    cfg.update(over)
    return cfg


def _stds(P, cfg, K, dtype):
    if 'std' in P:
        return P['std'].reshape(-1).to(dtype)
    s = torch.full((K,), cfg.pixel_std2, dtype=dtype)
    s[0] = cfg.pixel_std1
    return s


# ============================================================================ GENESIS (V1)
def genesis_forward(P, x, tape, cfg, training=True):
    """models/genesis_config.py:145-271 with modules/attention.py:84-133 (LatentSBP) and
    modules/component_vae.py:45-81, default two-stage configuration."""
    K, img = cfg.K_steps, cfg.img_size
    B = x.shape[0]
    dt = x.dtype
    upd = {}
    core = 'att_process.core'
    # --- LatentSBP.forward (attention.py:85-103): encode, K posterior samples through the LSTM
    h = O.sylvester_q_z_nn(x, P, core + '.q_z_nn', img, cfg.enc_norm, training, upd).flatten(1)
    mu = O.linear(h, P, core + '.q_z_mean')
    sigma = O.to_var(O.linear(h, P, core + '.q_z_var.0')).sqrt()       # VAE.py:126
    mu_k, sigma_k, z_k = [mu], [sigma], [mu + sigma * tape.normal(mu.shape, dt)]
    state = None
    for _ in range(1, K):
        out, state = O.lstm_cell(torch.cat([h, z_k[-1]], dim=1), state, P, 'att_process.lstm')
        lo = O.linear(out, P, 'att_process.linear')
        a, b = torch.chunk(lo, 2, dim=1)
        s = O.to_var(b).sqrt()
        mu_k.append(a)
        sigma_k.append(s)
        z_k.append(a + s * tape.normal(a.shape, dt))
    # --- batched mask decode + stick breaking (attention.py:110-130)
    logits = O.sylvester_decode(torch.cat(z_k, 0), P, core, img, cfg.dec_norm, training, upd)
    x_k = list(torch.chunk(logits, K, dim=0))
    log_m_k, log_s_k = O.stick_breaking([a[:, :1] for a in x_k])
    # genesis_config.py:169-171: drop mask K, last kept mask := its scope
    del log_m_k[-1]
    log_m_k[K - 1] = log_s_k[K - 1]
    act = O.act_fn('elu')
    if not cfg.two_stage:
        # one stage (genesis_config.py:121-126, 178-185): appearances decoded from the mask latents, no component VAE / KL
        x_r = O.broadcast_decoder(torch.cat(z_k, 0), P, 'decoder', img, cfg.comp_dec_layers, act)
        if cfg.pixel_bound:
            x_r = torch.sigmoid(x_r)
        x_r_k = list(torch.chunk(x_r, K, 0))
        recon = sum(m.exp() * xr for m, xr in zip(log_m_k, x_r_k))
        err = O.mixture_nll(x, log_m_k, x_r_k, _stds(P, cfg, K, dt))
        pmu_k, psig_k = O.autoreg_prior(z_k, P) if cfg.autoreg_prior else ([None] * K, [None] * K)
        kl_m_k = [O.mc_kl(z_k[k], mu_k[k], sigma_k[k], pmu_k[k], psig_k[k]) for k in range(K)]
        return dict(recon=recon, err=err, kl_m_k=kl_m_k, log_m_k=log_m_k, log_s_k=log_s_k, x_r_k=x_r_k,
                    att=dict(x_k=x_k, mu_k=mu_k, sigma_k=sigma_k, z_k=z_k, pmu_k=pmu_k[1:], psigma_k=psig_k[1:]), comp=None,
                    bn_updates=upd)
    # --- component VAE (component_vae.py:55-81)
    enc_in = torch.cat([torch.cat(log_m_k, 0), x.repeat(K, 1, 1, 1)], dim=1)
    sym = cfg.get('comp_symmetric', False)
    if sym:     # genesis_config.py:101-120: the component VAE uses the attention core's gated conv stacks
        enc = O.sylvester_q_z_nn(enc_in, P, 'comp_vae.encoder_module.0', img, cfg.enc_norm, training, upd).flatten(1)
    else:
        enc = O.monet_comp_encoder(enc_in, P, 'comp_vae.encoder_module', act)
    cmu, cps = torch.chunk(enc, 2, dim=1)
    csig = O.to_sigma(cps)
    cz = cmu + csig * tape.normal(cmu.shape, dt)
    if sym:
        x_r = O.sylvester_decode(cz, P, None, img, cfg.dec_norm, training, upd, nn_prefix='comp_vae.decoder_module.1',
                                 mean_prefix='comp_vae.decoder_module.2')
    else:
        x_r = O.broadcast_decoder(cz, P, 'comp_vae.decoder_module', img, cfg.comp_dec_layers, act)
    if cfg.pixel_bound:
        x_r = torch.sigmoid(x_r)
    x_r_k = list(torch.chunk(x_r, K, 0))
    cmu_k, csig_k, cz_k = (list(torch.chunk(t, K, 0)) for t in (cmu, csig, cz))
    # --- recon + losses (genesis_config.py:188-259)
    recon = sum(m.exp() * xr for m, xr in zip(log_m_k, x_r_k))
    err = O.mixture_nll(x, log_m_k, x_r_k, _stds(P, cfg, K, dt))
    if cfg.autoreg_prior:
        pmu_k, psig_k = O.autoreg_prior(z_k, P)
    else:
        pmu_k, psig_k = [None] * K, [None] * K
    kl_m_k = [O.mc_kl(z_k[k], mu_k[k], sigma_k[k], pmu_k[k], psig_k[k]) for k in range(K)]
    kl_l_k, cpmu_k, cpsig_k = [], [], []
    for k in range(K):
        if cfg.comp_prior:
            t = z_k[k]
            t = F.elu(O.linear(t, P, 'prior_mlp.0'))
            t = F.elu(O.linear(t, P, 'prior_mlp.2'))
            a, b = torch.chunk(O.linear(t, P, 'prior_mlp.4'), 2, dim=1)
            cpmu_k.append(torch.tanh(a))
            cpsig_k.append(O.to_prior_sigma(b))
            kl_l_k.append(O.mc_kl(cz_k[k], cmu_k[k], csig_k[k], cpmu_k[-1], cpsig_k[-1]))
        else:
            kl_l_k.append(O.mc_kl(cz_k[k], cmu_k[k], csig_k[k]))
    return dict(
        recon=recon, err=err, kl_m_k=kl_m_k, kl_l_k=kl_l_k,
        log_m_k=log_m_k, log_s_k=log_s_k, x_r_k=x_r_k,
        att=dict(x_k=x_k, mu_k=mu_k, sigma_k=sigma_k, z_k=z_k, pmu_k=pmu_k[1:], psigma_k=psig_k[1:]),
        comp=dict(mu_k=cmu_k, sigma_k=csig_k, z_k=cz_k, pmu_k=cpmu_k, psigma_k=cpsig_k),
        bn_updates=upd)


# ============================================================================ MONet
def mask_recon_log_softmax(logits_k):
    """MONet.get_mask_recon_stack, prior_mode='softmax', log=True (monet_config.py:136-140)."""
    ls = F.log_softmax(torch.stack(logits_k, dim=4), dim=4)
    return [ls[..., k] for k in range(len(logits_k))]


def mask_recon_log_scope(logits_k):
    """MONet.get_mask_recon_stack, prior_mode='scope', log=True (monet_config.py:141-153): stick-breaking over the mask logits,
    the last mask takes the remaining scope."""
    log_s = torch.zeros_like(logits_k[0])
    out = []
    for step, a in enumerate(logits_k):
        if step == len(logits_k) - 1:
            out.append(log_s)
        else:
            out.append(log_s + F.logsigmoid(a))
            log_s = log_s + F.logsigmoid(-a)
    return out


def monet_kl_m(log_m_k, log_m_r_k):
    """MONet.kl_m_loss (monet_config.py:157-170)."""
    B = log_m_k[0].shape[0]
    K = len(log_m_k)
    q = torch.stack(log_m_k, dim=4).exp().clamp_min(1e-5).reshape(-1, K)
    p = torch.stack(log_m_r_k, dim=4).exp().clamp_min(1e-5).reshape(-1, K)
    return O.categorical_kl(q, p).view(B, -1).sum(dim=1)


def monet_forward(P, x, tape, cfg, training=True):
    """models/monet_config.py:74-128 with modules/attention.py:31-51 (SimpleSBP over a UNet with
    InstanceNorm) and the component VAE with nout=4, ReLU, no pixel bound inside the VAE."""
    K, img = cfg.K_steps, cfg.img_size
    dt = x.dtype
    nb = int(math.log2(img) - 1)
    core = 'att_process.core'
    log_s_k = [torch.zeros_like(x[:, :1])]
    log_m_k = []
    for k in range(K - 1):
        u = O.unet(torch.cat([x, log_s_k[k]], dim=1), P, core, nb, 'in')
        a = F.conv2d(u, P[core + '.final_conv.weight'], P[core + '.final_conv.bias'])[:, :1]
        log_m_k.append(log_s_k[k] + F.logsigmoid(a))
        log_s_k.append(log_s_k[k] + F.logsigmoid(-a))
    log_m_k.append(log_s_k[-1])
    act = O.act_fn('relu')
    enc_in = torch.cat([torch.cat(log_m_k, 0), x.repeat(K, 1, 1, 1)], dim=1)
    enc = O.monet_comp_encoder(enc_in, P, 'comp_vae.encoder_module', act)
    cmu, cps = torch.chunk(enc, 2, dim=1)
    csig = O.to_sigma(cps)
    cz = cmu + csig * tape.normal(cmu.shape, dt)
    dec = O.broadcast_decoder(cz, P, 'comp_vae.decoder_module', img, cfg.comp_dec_layers, act)
    dec_k = list(torch.chunk(dec, K, 0))
    x_r_k = [d[:, :3] for d in dec_k]
    if cfg.pixel_bound:
        x_r_k = [torch.sigmoid(t) for t in x_r_k]
    if cfg.get('prior_mode', 'softmax') == 'scope':
        log_m_r_k = mask_recon_log_scope([d[:, 3:] for d in dec_k])
    else:
        log_m_r_k = mask_recon_log_softmax([d[:, 3:] for d in dec_k])
    recon = sum(m.exp() * xr for m, xr in zip(log_m_k, x_r_k))
    err = O.mixture_nll(x, log_m_k, x_r_k, _stds(P, cfg, K, dt))
    kl_m = monet_kl_m(log_m_k, log_m_r_k)
    cmu_k, csig_k, cz_k = (list(torch.chunk(t, K, 0)) for t in (cmu, csig, cz))
    kl_l_k = [O.mc_kl(cz_k[k], cmu_k[k], csig_k[k]) for k in range(K)]
    return dict(recon=recon, err=err, kl_m=kl_m, kl_l_k=kl_l_k, log_m_k=log_m_k, log_s_k=log_s_k,
                x_r_k=x_r_k, log_m_r_k=log_m_r_k,
                comp=dict(mu_k=cmu_k, sigma_k=csig_k, z_k=cz_k), bn_updates={})


# ============================================================================ GENESIS-V2
def icsbp(colour, u, log_sigma, steps, kernel='gaussian', dynamic_K=False):
    """InstanceColouringSBP.forward (modules/attention.py:177-223), dynamic_K=False; `kernel` as :195-205.
    colour [B,C,H,W] (after SemiConv), u [B,1,H,W] uniform draws, `steps` = K-1.  The bilinear resize
    of the scope at :185-186 is the identity because colour and scope share img_size."""
    B, C, H, W = colour.shape
    log_s_k = [torch.zeros(B, 1, H, W, dtype=colour.dtype)]
    log_m_k, seeds, idxs = [], [], []
    flat = colour.flatten(2)
    for k in range(steps):
        probs = u * log_s_k[k].exp()
        idx = probs.flatten(2).argmax(2).flatten()                           # :187-188
        seed = flat[torch.arange(B), :, idx]                                  # :190-192, keeps grad
        dist = ((colour - seed.view(B, C, 1, 1)) ** 2).sum(1)                 # blocks.py:63-71
        if kernel == 'gaussian':
            alpha = torch.exp(-dist / log_sigma.exp())                        # :198-200
        elif kernel == 'laplacian':                                           # :195-197, blocks.py:49-61
            alpha = torch.exp(-O.clamp_ste(dist, 1e-10, 1e10).sqrt() / log_sigma.exp())
        elif kernel == 'epanechnikov':                                        # :201-203
            alpha = (1 - dist / log_sigma.exp()).relu()
        else:
            raise ValueError("No valid kernel.")
        alpha = alpha.unsqueeze(1)
        alpha = O.clamp_ste(alpha, 0.01, 0.99)                                # :213
        seeds.append(seed)                                                    # :193 (before the early exit)
        idxs.append(idx)
        log_m = log_s_k[k] + torch.log(alpha)
        if dynamic_K and log_m.exp().sum() < 20:                              # :218-219 (batch of one, :168-169)
            assert B == 1
            break
        log_m_k.append(log_m)
        log_s_k.append(log_s_k[k] + torch.log(1 - alpha))
    log_m_k.append(log_s_k[-1])
    return log_m_k, log_s_k, seeds, idxs


def v2_decoder(z, P, img_size):
    """GenesisV2.decoder_module (genesisv2_config.py:88-99): broadcast to (img/16)^2 + coords, four
    ConvTranspose 5x5 s2 p2 op1 + GroupNorm(8) + ReLU, 1x1 conv to 4 channels."""
    d = img_size // 16
    n = z.shape[0]
    h = z.view(n, -1, 1, 1).expand(-1, -1, d, d)
    h = torch.cat([h, O.pixel_coords(d, z.dtype).expand(n, -1, -1, -1)], dim=1)
    for ci, ni in ((1, 2), (4, 5), (7, 8), (10, 11)):
        c = 'decoder_module.%d' % ci
        g = 'decoder_module.%d' % ni
        h = F.conv_transpose2d(h, P[c + '.weight'], P[c + '.bias'], stride=2, padding=2,
                               output_padding=1)
        h = F.relu(O.group_norm(h, 8, P[g + '.weight'], P[g + '.bias']))
    return F.conv2d(h, P['decoder_module.13.weight'], P['decoder_module.13.bias'])


def _seed_kink_margin(pre, idxs):
    """Test conditioning guard (not part of the reference): the smallest |pre-ReLU| value of seg_head over all channels of the
    IC-SBP seed pixels.  Every pixel's distance to a seed depends on the seed pixel's embedding, so its gradient is
    concentrated there; with a channel within rounding of the ReLU kink the gradient is discontinuous and two correct
    implementations can disagree by > 10 % on seg_head.0.weight / seg_head.1.bias."""
    if not idxs:
        return float('inf')
    B, C = pre.shape[0], pre.shape[1]
    flat = pre.detach().reshape(B, C, -1)
    m = float('inf')
    for idx in idxs:
        idx = idx.reshape(-1).long()
        if idx.numel() != B:
            continue
        v = flat[torch.arange(B), :, idx]
        m = min(m, v.abs().min().item())
    return m


def genesisv2_forward(P, x, tape, cfg, training=True):
    """models/genesisv2_config.py:110-203 (dynamic_K=False, klm_loss=False defaults)."""
    K, img = cfg.K_steps, cfg.img_size
    B = x.shape[0]
    dt = x.dtype
    nb = int(math.log2(img) - 1)

    pre_relu = {}

    def conv_gn_relu(h, name):
        h = F.conv2d(h, P[name + '.0.weight'], None, padding=1)
        pre_relu[name] = O.group_norm(h, 8, P[name + '.1.weight'], P[name + '.1.bias'])
        return F.relu(pre_relu[name])

    enc_feat = F.relu(O.unet(x, P, 'encoder', nb, 'gn'))                       # :114-115
    seg = conv_gn_relu(enc_feat, 'seg_head')
    # SemiConv (blocks.py:167-178)
    if getattr(cfg, 'semiconv', True):
        cw, cb = P['att_process.colour_head.conv.weight'], P['att_process.colour_head.conv.bias']
        out = P['att_process.colour_head.gate.gate'] * F.conv2d(seg, cw, cb)
        delta = out[:, -2:]
        uv = torch.cat([torch.zeros(1, out.shape[1] - 2, img, img, dtype=dt), O.pixel_coords(img, dt)], 1)
        colour = out + uv
    else:                                                                      # attention.py:159-160: plain 1x1 conv
        colour = F.conv2d(seg, P['att_process.colour_head.weight'], P['att_process.colour_head.bias'])
        delta = None
    kern = getattr(cfg, 'kernel', 'gaussian')
    if getattr(cfg, 'dynamic_K', False):                                       # genesisv2_config.py:118-137
        if B > 1:       # image by image (one uniform draw each), padded with -1e10 masks; log_s_k / att_stats are None there
            per = [icsbp(colour[b:b + 1], tape.uniform((1, 1, img, img), dt), P['att_process.log_sigma'], K - 1, kern, True)
                   for b in range(B)]
            pad = torch.full((1, 1, img, img), -1e10, dtype=dt)
            log_m_k = [torch.cat([r[0][k] if k < len(r[0]) else pad for r in per], 0) for k in range(K)]
            log_s_k, seeds, idxs = None, None, None
            n_masks = [len(r[0]) for r in per]
        else:
            log_m_k, log_s_k, seeds, idxs = icsbp(colour, tape.uniform((1, 1, img, img), dt), P['att_process.log_sigma'], K - 1,
                                                  kern, True)
            n_masks = [len(log_m_k)]
            K = len(log_m_k)
    else:
        u = tape.uniform((B, 1, img, img), dt)                                 # attention.py:177-178
        log_m_k, log_s_k, seeds, idxs = icsbp(colour, u, P['att_process.log_sigma'], K - 1, kern)
        n_masks = [K] * B
    # slot latents (:145-161); feat_head evaluated once -- identical values to the K recomputations
    f = conv_gn_relu(enc_feat, 'feat_head.0')
    f = F.conv2d(f, P['feat_head.1.weight'], P['feat_head.1.bias'])
    mu_k, sigma_k, z_k = [], [], []
    for log_m in log_m_k:
        m = log_m.exp()
        obj = (m * f).sum((2, 3)) / (m.sum((2, 3)) + 1e-5)
        t = O.layer_norm(obj, P['z_head.0.weight'], P['z_head.0.bias'])
        t = F.relu(O.linear(t, P, 'z_head.1'))
        a, b = torch.chunk(O.linear(t, P, 'z_head.3'), 2, dim=1)
        s = O.to_sigma(b)
        mu_k.append(a)
        sigma_k.append(s)
        z_k.append(a + s * tape.normal(a.shape, dt))
    # decode (:205-225)
    dec = v2_decoder(torch.cat(z_k, 0), P, img)
    dec_k = list(torch.chunk(dec, K, 0))
    x_r_k = [d[:, :3] for d in dec_k]
    if cfg.pixel_bound:
        x_r_k = [torch.sigmoid(t) for t in x_r_k]
    log_m_r_k = mask_recon_log_softmax([d[:, 3:] for d in dec_k])
    recon = sum(m.exp() * xr for m, xr in zip(log_m_r_k, x_r_k))
    err = O.mixture_nll(x, log_m_r_k, x_r_k, cfg.pixel_std1)                   # :169
    if cfg.autoreg_prior:
        pmu_k, psig_k = O.autoreg_prior(z_k, P)
    else:
        pmu_k, psig_k = [None] * K, [None] * K
    kl_l_k = [O.mc_kl(z_k[k], mu_k[k], sigma_k[k], pmu_k[k], psig_k[k]) for k in range(K)]
    out = dict(recon=recon, err=err, kl_l_k=kl_l_k, log_m_k=log_m_k, log_s_k=log_s_k, x_r_k=x_r_k,
               log_m_r_k=log_m_r_k,
               att=dict(colour=colour, delta=delta, seeds=seeds, seed_idx=idxs,
                        seed_kink_margin=_seed_kink_margin(pre_relu['seg_head'], idxs)), n_masks=n_masks,
               comp=dict(mu_k=mu_k, sigma_k=sigma_k, z_k=z_k, pmu_k=pmu_k[1:], psigma_k=psig_k[1:]),
               bn_updates={})
    if cfg.get('klm_loss', False):          # genesisv2_config.py:171-176: KL(masks || reconstructed masks), the latter detached by default
        lmr = [m.detach() for m in log_m_r_k] if cfg.get('detach_mr_in_klm', True) else log_m_r_k
        out['kl_m'] = monet_kl_m(log_m_k, lmr)
    return out


# ============================================================================ BaselineVAE (config c1)
VAE_DEFAULTS = dict(latent_dimension=64, broadcast_decoder=False, pixel_bound=True, pixel_std=0.7)  # models/vae_config.py:26-31


def vae_forward(P, x, tape, cfg, training=True):
    """BaselineVAE.forward (models/vae_config.py:63-89) over third_party/sylvester/VAE.forward (VAE.py:155-168): gated conv
    encoder without norms -> (mu, sqrt(to_var)) -> rsample -> gated conv-transpose decoder -> sigmoid; err = -sum log N(x;
    recon, pixel_std); kl = sum_d log q(z) - log N(z; 0, 1)."""
    img, dt = cfg.img_size, x.dtype
    upd = {}
    h = O.sylvester_q_z_nn(x, P, 'vae.q_z_nn', img, None, training, upd).flatten(1)
    mu = O.linear(h, P, 'vae.q_z_mean')
    sigma = O.to_var(O.linear(h, P, 'vae.q_z_var.0')).sqrt()
    z = mu + sigma * tape.normal(mu.shape, dt)
    if cfg.get('broadcast_decoder', False):        # vae_config.py:53-61
        hdec = F.elu(O.broadcast_decoder(z, P, 'vae.p_x_nn.1', img, 4, O.act_fn('elu')))
        recon = F.conv2d(hdec, P['vae.p_x_mean.weight'], P['vae.p_x_mean.bias'])
    else:
        recon = O.sylvester_decode(z, P, 'vae', img, None, training, upd)
    if cfg.pixel_bound:
        recon = torch.sigmoid(recon)
    err = -O.normal_log_prob(x, recon, torch.tensor(cfg.pixel_std, dtype=dt)).sum(dim=(1, 2, 3))
    kl = O.mc_kl(z, mu, sigma)
    return dict(recon=recon, err=err, kl_l=kl, mu=mu, sigma=sigma, z=z, bn_updates=upd)


# ============================================================================ sample() (ancestral sampling)
def genesis_sample(P, batch_size, tape, cfg, training=False):
    """Genesis.sample (models/genesis_config.py:345-425) with LatentSBP.masks_from_zm_k (modules/attention.py:53-74):
    z_m,1 ~ N(0,1); z_m,k ~ N(mu, sigma) from the prior LSTM (no tanh on mu in sample(), :359); every z_m decoded
    SEPARATELY (:61), K+1 masks cut back to K (:371-373); z_c,k ~ N(tanh, to_prior_sigma) from prior_mlp(z_m,k), one draw
    per slot (:392-397); batched component decode; composite."""
    K, img = cfg.K_steps, cfg.img_size
    dt = torch.float32
    ldim = cfg.attention_latents
    upd = {}
    zm_k = [tape.normal((batch_size, ldim), dt)]
    if cfg.autoreg_prior:
        state = None
        for _ in range(1, K):
            out, state = O.lstm_cell(zm_k[-1], state, P, 'prior_lstm')
            lo = O.linear(out, P, 'prior_linear')
            mu, sigma = lo[:, :ldim], O.to_prior_sigma(lo[:, ldim:])
            zm_k.append(mu + sigma * tape.normal(mu.shape, dt))
    else:
        zm_k += [tape.normal((batch_size, ldim), dt) for _ in range(1, K)]
    logits_k = [O.sylvester_decode(zm, P, 'att_process.core', img, cfg.dec_norm, training, upd)[:, :1] for zm in zm_k]
    log_m_k, log_s_k = O.stick_breaking(logits_k)
    del log_m_k[-1]
    log_m_k[K - 1] = log_s_k[K - 1]
    if not cfg.two_stage:               # genesis_config.py:404-409
        x = O.broadcast_decoder(torch.cat(zm_k, 0), P, 'decoder', img, cfg.comp_dec_layers, O.act_fn('elu'))
        if cfg.pixel_bound:
            x = torch.sigmoid(x)
        x_k = list(torch.chunk(x, K, 0))
        image = sum(m.exp() * xk for m, xk in zip(log_m_k, x_k))
        return dict(image=image, x_k=x_k, log_m_k=log_m_k, log_s_k=log_s_k)
    zc_k = []
    for zm in zm_k:
        if cfg.comp_prior:
            t = F.elu(O.linear(zm, P, 'prior_mlp.0'))
            t = F.elu(O.linear(t, P, 'prior_mlp.2'))
            a, b = torch.chunk(O.linear(t, P, 'prior_mlp.4'), 2, dim=1)
            mu, sigma = torch.tanh(a), O.to_prior_sigma(b)
            zc_k.append(mu + sigma * tape.normal(mu.shape, dt))
        else:
            zc_k.append(tape.normal((batch_size, cfg.comp_ldim), dt))
    if cfg.get('comp_symmetric', False):
        x = O.sylvester_decode(torch.cat(zc_k, 0), P, None, img, cfg.dec_norm, training, upd, nn_prefix='comp_vae.decoder_module.1',
                               mean_prefix='comp_vae.decoder_module.2')
    else:
        x = O.broadcast_decoder(torch.cat(zc_k, 0), P, 'comp_vae.decoder_module', img, cfg.comp_dec_layers, O.act_fn('elu'))
    if cfg.pixel_bound:
        x = torch.sigmoid(x)
    x_k = list(torch.chunk(x, K, 0))
    image = sum(m.exp() * xk for m, xk in zip(log_m_k, x_k))
    return dict(image=image, x_k=x_k, log_m_k=log_m_k, log_s_k=log_s_k)


def genesisv2_sample(P, batch_size, tape, cfg, training=False):
    """GenesisV2.sample (models/genesisv2_config.py:227-256): z_1 ~ N(0,1), z_k ~ N(tanh(.), to_prior_sigma(.)) from the
    prior LSTM, then decode_latents (:205-225)."""
    K, img, fd = cfg.K_steps, cfg.img_size, cfg.feat_dim
    dt = torch.float32
    z_k = [tape.normal((batch_size, fd), dt)]
    if cfg.autoreg_prior:
        state = None
        for _ in range(1, K):
            out, state = O.lstm_cell(z_k[-1], state, P, 'prior_lstm')
            a, b = torch.chunk(O.linear(out, P, 'prior_linear'), 2, dim=1)
            mu, sigma = torch.tanh(a), O.to_prior_sigma(b)
            z_k.append(mu + sigma * tape.normal(mu.shape, dt))
    else:
        z_k += [tape.normal((batch_size, fd), dt) for _ in range(1, K)]
    dec_k = [v2_decoder(z, P, img) for z in z_k]
    x_k = [torch.sigmoid(d[:, :3]) if cfg.pixel_bound else d[:, :3] for d in dec_k]
    log_m_k = mask_recon_log_softmax([d[:, 3:] for d in dec_k])
    image = sum(m.exp() * xk for m, xk in zip(log_m_k, x_k))
    return dict(image=image, x_k=x_k, log_m_k=log_m_k)


def monet_sample(P, batch_size, tape, cfg, training=False):
    """MONet.sample (models/monet_config.py:172-198): one N(0,1) draw for all K*B slots, broadcast-decode, softmax masks."""
    K, img = cfg.K_steps, cfg.img_size
    z = tape.normal((batch_size * K, cfg.comp_ldim), torch.float32)
    dec = O.broadcast_decoder(z, P, 'comp_vae.decoder_module', img, cfg.comp_dec_layers, O.act_fn('relu'))
    x = torch.sigmoid(dec[:, :3]) if cfg.pixel_bound else dec[:, :3]
    x_k = list(torch.chunk(x, K, 0))
    if cfg.get('prior_mode', 'softmax') == 'scope':
        log_m_k = mask_recon_log_scope(list(torch.chunk(dec[:, 3:], K, 0)))
    else:
        log_m_k = mask_recon_log_softmax(list(torch.chunk(dec[:, 3:], K, 0)))
    image = sum(m.exp() * xk for m, xk in zip(log_m_k, x_k))
    return dict(image=image, x_k=x_k, log_m_k=log_m_k)


SAMPLE = {'genesis': genesis_sample, 'genesisv2': genesisv2_sample, 'monet': monet_sample}
FORWARD = {'genesis': genesis_forward, 'genesisv2': genesisv2_forward, 'monet': monet_forward, 'vae': vae_forward}


def total_loss(out, beta=1.0):
    """Scalar the caller differentiates: train.py:227-259 with GECO's beta held fixed --
    err.mean(0) + beta * (sum_k mean_b kl_l_k + sum_k mean_b kl_m_k [+ mean kl_m])."""
    loss = out['err'].mean(0)
    kl = 0.0
    for key in ('kl_l_k', 'kl_m_k'):
        if key in out:
            kl = kl + torch.stack(out[key], dim=1).mean(0).sum()
    if 'kl_m' in out:
        kl = kl + out['kl_m'].mean(0)
    if 'kl_l' in out:                       # BaselineVAE (train.py:228-229)
        kl = kl + out['kl_l'].mean(0)
    return loss + beta * kl
