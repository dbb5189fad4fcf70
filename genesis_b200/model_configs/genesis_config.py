"""GENESIS (V1) plug-in: drop-in for the reference's models/genesis_config.py.

Same Forge contract -- importing this file registers the model flags, `load(cfg)` returns an nn.Module whose
`forward(x)` returns `(recon, losses, stats, att_stats, comp_stats)`, plus `sample()` / `get_features()` --
and the same state_dict names, so `train.py --model_config .../genesis_config.py` runs unchanged
(reference models/genesis_config.py:33-56, 145-271).  All convolutions, conv-transposes, linears, norms,
the stick-breaking scan and the mixture likelihood run in hand-written sm_100a kernels (genesis_b200.ops).
"""
import os
import sys

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import genesis_b200  # noqa: E402

try:
    from forge import flags
except ImportError:  # reference Forge not importable: use the bundled stand-in
    genesis_b200.enable_compat()
    from forge import flags
try:
    from attrdict import AttrDict
except ImportError:
    genesis_b200.enable_compat()
    from attrdict import AttrDict

from genesis_b200 import holders as H  # noqa: E402
from genesis_b200 import ops  # noqa: E402
from genesis_b200.noise import NoiseMixin  # noqa: E402,F401  (re-exported: the other plug-ins use genesis_config.NoiseMixin)

# Flag names / defaults: reference models/genesis_config.py:33-52
flags.DEFINE_boolean('two_stage', True, 'Use two stages if two, else only one.')
flags.DEFINE_boolean('autoreg_prior', True, 'Autoregressive prior.')
flags.DEFINE_boolean('comp_prior', True, 'Component prior.')
flags.DEFINE_integer('attention_latents', 64, 'Latent dimension.')
flags.DEFINE_string('enc_norm', 'bn', '{bn, in} - norm type in encoder.')
flags.DEFINE_string('dec_norm', 'bn', '{bn, in} - norm type in decoder.')
flags.DEFINE_integer('comp_enc_channels', 32, 'Starting number of channels.')
flags.DEFINE_integer('comp_ldim', 16, 'Latent dimension of the VAE.')
flags.DEFINE_integer('comp_dec_channels', 32, 'Num channels in Broadcast Decoder.')
flags.DEFINE_integer('comp_dec_layers', 4, 'Num layers in Broadcast Decoder.')
flags.DEFINE_boolean('comp_symmetric', False, 'Use same encoder/decoder as in attention VAE.')
flags.DEFINE_boolean('pixel_bound', True, 'Bound pixel values to [0, 1].')
flags.DEFINE_float('pixel_std1', 0.7, 'StdDev of reconstructed pixels.')
flags.DEFINE_float('pixel_std2', 0.7, 'StdDev of reconstructed pixels.')
flags.DEFINE_boolean('montecarlo_kl', True, 'Evaluate KL via MC samples.')


def load(cfg):
    return Genesis(cfg)


class LatentSBPHolder(nn.Module):
    """Holder for modules/attention.py:77-82 (core + posterior LSTM + Linear)."""

    def __init__(self, core):
        super().__init__()
        self.core = core
        self.lstm = nn.LSTM(core.z_size + 256, 2 * core.z_size)
        self.linear = nn.Linear(2 * core.z_size, 2 * core.z_size)


class Genesis(nn.Module, NoiseMixin):

    def __init__(self, cfg):
        super().__init__()
        self.K_steps = cfg.K_steps
        self.img_size = cfg.img_size
        self.two_stage = cfg.two_stage
        self.autoreg_prior = cfg.autoreg_prior
        self.comp_prior = False
        if self.two_stage and self.K_steps > 1:
            self.comp_prior = cfg.comp_prior
        self.ldim = cfg.attention_latents
        self.pixel_bound = cfg.pixel_bound
        if not hasattr(cfg, 'comp_symmetric'):
            cfg.comp_symmetric = False
        self.debug = cfg.debug
        self.side_stream = True         # overlap the prior / KL branch with the decoders (see forward)
        assert cfg.montecarlo_kl == True  # noqa: E712  (reference genesis_config.py:80)
        self.comp_symmetric = bool(cfg.comp_symmetric) and self.two_stage
        input_channels = cfg.input_channels if hasattr(cfg, 'input_channels') else 3
        # construction order == reference (genesis_config.py:86-138) so seeded init is identical
        core = H.SylvesterVAE(self.ldim, [input_channels, cfg.img_size, cfg.img_size], 1,
                              cfg.enc_norm, cfg.dec_norm)
        if self.K_steps > 1:
            self.att_steps = self.K_steps
            self.att_process = LatentSBPHolder(core)
        else:
            # K_steps == 1 (reference genesis_config.py:94-96, 161-166): no attention process -- the core was built (it consumes
            # the seeded generator exactly as in the reference) but is not part of the model; a single all-ones mask.
            assert self.two_stage       # reference :122
            self.autoreg_prior = False
        if self.two_stage:
            self.comp_vae = H.ComponentVAEHolder(nout=input_channels, cfg=cfg)
            if self.comp_symmetric:     # reference genesis_config.py:101-120
                H.make_symmetric_component_vae(self.comp_vae, input_channels, cfg, core.last_kernel_size)
        else:       # one stage: the mask latents are decoded into appearances directly (reference genesis_config.py:121-126)
            self.decoder = H.BroadcastDecoderHolder(self.ldim, input_channels, cfg.comp_dec_channels, cfg.comp_dec_layers,
                                                    cfg.img_size)
        if self.autoreg_prior:
            self.prior_lstm = nn.LSTM(self.ldim, 256)
            self.prior_linear = nn.Linear(256, 2 * self.ldim)
        if self.comp_prior:
            self.prior_mlp = nn.Sequential(
                nn.Linear(self.ldim, 256), nn.Identity(), nn.Linear(256, 256), nn.Identity(),
                nn.Linear(256, 2 * cfg.comp_ldim))
        std = cfg.pixel_std2 * torch.ones(1, 1, 1, 1, self.K_steps)
        std[0, 0, 0, 0, 0] = cfg.pixel_std1
        self.register_buffer('std', std)

    # --------------------------------------------------------------------------------------------
    def _masks(self, x):
        """LatentSBP.forward (reference attention.py:84-133) + the K-th mask fix-up (genesis_config.py:169-171)."""
        K, B = self.K_steps, x.shape[0]
        ap, core = self.att_process, self.att_process.core
        tc = ops.get_precision() == 'tf32'      # tensor-core kernels want 32-channel k-blocks: zero-pad the image
        xh = ops.to_nhwc_padded(x, 32) if tc else ops.to_nhwc(x)
        h = H.sylvester_encode(core, xh, self.training)                                 # [B,256]
        wmv = torch.cat([core.q_z_mean.weight, core.q_z_var[0].weight], 0)
        bmv = torch.cat([core.q_z_mean.bias, core.q_z_var[0].bias], 0)
        # sigma = to_sigma(raw): sqrt(to_var(.)) == to_sigma(.) (VAE.py:126)
        z, mu, sigma = H.gauss_head(ops.linear(h, wmv, bmv), self._normal((B, self.ldim), x))
        mu_k, sigma_k, z_k = [mu], [sigma], [z]
        state = None
        for _ in range(1, K):
            out, state = H.lstm_step(torch.cat([h, z_k[-1]], dim=1), state, ap.lstm)
            z, a, s = H.gauss_head(ops.linear(out, ap.linear.weight, ap.linear.bias), self._normal((B, self.ldim), x))
            mu_k.append(a)
            sigma_k.append(s)
            z_k.append(z)
        logits = H.sylvester_decode(core, torch.cat(z_k, 0), self.training)             # [K*B,1,H,W]
        logits = logits.view(K, B, 1, self.img_size, self.img_size)
        log_m, log_s = ops.sbp_scan(logits, K)
        att_stats = AttrDict(x_k=list(logits.unbind(0)), mu_k=mu_k, sigma_k=sigma_k, z_k=z_k)
        return log_m, log_s, att_stats

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('genesis_b200 runs on CUDA (sm_100a) only; there is no CPU path')
        K, B = self.K_steps, x.shape[0]
        x = x.contiguous().float()
        if K == 1:
            return self._forward_single_slot(x)
        log_m, log_s, att_stats = self._masks(x)                     # [K,B,1,H,W], [K+1,B,1,H,W]
        z_k = att_stats.z_k
        if not self.two_stage:
            return self._forward_one_stage(x, log_m, log_s, att_stats)
        # The prior / KL branch (LSTM prior over z_k, prior MLP, MC-KL sums: ~250 tiny latency-bound kernels on [B,64]
        # tensors) does not feed the decoders, so it runs on a side stream beside the large decoder kernels.  Autograd
        # replays each node's backward on its forward stream, so the overlap carries to the backward pass; under CUDA
        # graph capture this becomes a parallel branch of the graph.
        cur = torch.cuda.current_stream()
        side = ops.side_stream(x.device) if (self.side_stream and ops.side_streams_enabled() and torch.is_grad_enabled()) else cur
        if side is not cur:
            side.wait_stream(cur)
            for t_ in z_k + att_stats.mu_k + att_stats.sigma_k:      # allocated on `cur`, read (fwd and bwd) on `side`
                t_.record_stream(side)
        with torch.cuda.stream(side):
            # --- KL terms of the mask latents (reference genesis_config.py:198-228, 288-343)
            if self.autoreg_prior:
                pmu, psig = H.autoreg_prior(z_k, self.prior_lstm, self.prior_linear)
            else:
                pmu, psig = [], []
            kl_m_k = [H.mc_kl(z_k[0], att_stats.mu_k[0], att_stats.sigma_k[0])]
            for k in range(1, K):
                if self.autoreg_prior:
                    kl_m_k.append(H.mc_kl(z_k[k], att_stats.mu_k[k], att_stats.sigma_k[k], pmu[k - 1], psig[k - 1]))
                else:
                    kl_m_k.append(H.mc_kl(z_k[k], att_stats.mu_k[k], att_stats.sigma_k[k]))
            if self.comp_prior:
                pm = self.prior_mlp
                t = ops.linear(torch.cat(z_k, 0), pm[0].weight, pm[0].bias, 'elu')
                t = ops.linear(t, pm[2].weight, pm[2].bias, 'elu')
                cpmu, cpsig = H.prior_head(ops.linear(t, pm[4].weight, pm[4].bias))
        # --- component VAE (reference component_vae.py:45-81), K slots batched k-major
        cv = self.comp_vae
        packed = ops.comp_pack(x, log_m, 32 if ops.get_precision() == 'tf32' else 4)
        if self.comp_symmetric:
            enc = H.symmetric_comp_encode(cv, packed, self.training)
        else:
            enc = H.comp_encode(cv.encoder_module, packed, 'elu')
        cz, cmu, csig = H.gauss_head(enc, self._normal((enc.shape[0], enc.shape[1] // 2), x))
        if side is not cur:
            side.wait_stream(cur)           # cz, cmu, csig are ready
            for t_ in (cz, cmu, csig, enc):
                t_.record_stream(side)
        with torch.cuda.stream(side):
            # --- KL of the component latents (reference genesis_config.py:230-259)
            kl = H.mc_kl(cz, cmu, csig, cpmu, cpsig) if self.comp_prior else H.mc_kl(cz, cmu, csig)
        if self.comp_symmetric:
            x_r = H.symmetric_comp_decode(cv, cz, self.training, 3 if self.pixel_bound else 0)
        else:
            x_r = H.broadcast_decode(cv.decoder_module, cz, 'elu', 3 if self.pixel_bound else 0)
        x_r = x_r.view(K, B, x.shape[1], self.img_size, self.img_size)
        # --- reconstruction + mixture likelihood (reference genesis_config.py:188-196)
        err, recon, _ = ops.mixture_nll(x, x_r, log_m, self.std.reshape(-1), False)
        if side is not cur:
            cur.wait_stream(side)           # join: the KL terms are consumed by the caller on the current stream
            for t_ in kl_m_k + [kl] + pmu + psig + ([cpmu, cpsig] if self.comp_prior else []):
                t_.record_stream(cur)
        losses = AttrDict()
        losses['err'] = err
        losses['kl_m_k'] = kl_m_k
        att_stats['pmu_k'], att_stats['psigma_k'] = pmu, psig
        comp_stats = AttrDict(mu_k=list(torch.chunk(cmu, K, 0)), sigma_k=list(torch.chunk(csig, K, 0)),
                              z_k=list(torch.chunk(cz, K, 0)))
        if self.comp_prior:
            comp_stats['pmu_k'] = list(torch.chunk(cpmu, K, 0))
            comp_stats['psigma_k'] = list(torch.chunk(cpsig, K, 0))
        losses['kl_l_k'] = list(torch.chunk(kl, K, 0))
        # --- tracking (reference genesis_config.py:262-264)
        log_m_k = list(log_m.unbind(0))
        x_r_k = list(x_r.unbind(0))
        with torch.no_grad():
            mx = x_r * log_m.exp()
        stats = AttrDict(recon=recon, log_m_k=log_m_k, log_s_k=list(log_s.unbind(0)), x_r_k=x_r_k,
                         mx_r_k=list(mx.unbind(0)))
        if self.debug or not self.training:
            assert len(log_m_k) == self.K_steps
            check_log_masks(log_m_k)
        return recon, losses, stats, att_stats, comp_stats

    def _forward_single_slot(self, x):
        """K_steps == 1 (reference genesis_config.py:161-166, 226-228, 249-255): the only mask is all ones (log m = 0, log s =
        -1e10), the component VAE sees cat(0, x), the component prior is N(0,1), `kl_m` is the constant 0 and att_stats None."""
        B = x.shape[0]
        cv = self.comp_vae
        log_m = x.new_zeros(1, B, 1, self.img_size, self.img_size)
        packed = ops.comp_pack(x, log_m, 32 if ops.get_precision() == 'tf32' else 4)
        if self.comp_symmetric:
            enc = H.symmetric_comp_encode(cv, packed, self.training)
        else:
            enc = H.comp_encode(cv.encoder_module, packed, 'elu')
        cz, cmu, csig = H.gauss_head(enc, self._normal((enc.shape[0], enc.shape[1] // 2), x))
        kl = H.mc_kl(cz, cmu, csig)
        if self.comp_symmetric:
            x_r = H.symmetric_comp_decode(cv, cz, self.training, 3 if self.pixel_bound else 0)
        else:
            x_r = H.broadcast_decode(cv.decoder_module, cz, 'elu', 3 if self.pixel_bound else 0)
        x_r = x_r.view(1, B, x.shape[1], self.img_size, self.img_size)
        err, recon, _ = ops.mixture_nll(x, x_r, log_m, self.std.reshape(-1), False)
        losses = AttrDict(err=err, kl_m=x.new_zeros(()), kl_l_k=[kl])
        comp_stats = AttrDict(mu_k=[cmu], sigma_k=[csig], z_k=[cz])
        with torch.no_grad():
            mx = x_r * log_m.exp()
        stats = AttrDict(recon=recon, log_m_k=list(log_m.unbind(0)), log_s_k=[torch.full_like(log_m[0], -1e10)],
                         x_r_k=list(x_r.unbind(0)), mx_r_k=list(mx.unbind(0)))
        return recon, losses, stats, None, comp_stats

    def _forward_one_stage(self, x, log_m, log_s, att_stats):
        """two_stage=False (reference genesis_config.py:178-185, 198-228): appearances = BroadcastDecoder(z_m,k), no component
        VAE and no component KL; losses = {err, kl_m_k}; comp_stats is None."""
        K, B = self.K_steps, x.shape[0]
        z_k = att_stats.z_k
        x_r = H.broadcast_decode(self.decoder, torch.cat(z_k, 0), 'elu', 3 if self.pixel_bound else 0)
        x_r = x_r.view(K, B, x.shape[1], self.img_size, self.img_size)
        err, recon, _ = ops.mixture_nll(x, x_r, log_m, self.std.reshape(-1), False)
        if self.autoreg_prior:
            pmu, psig = H.autoreg_prior(z_k, self.prior_lstm, self.prior_linear)
        else:
            pmu, psig = [], []
        kl_m_k = [H.mc_kl(z_k[0], att_stats.mu_k[0], att_stats.sigma_k[0])]
        for k in range(1, K):
            if self.autoreg_prior:
                kl_m_k.append(H.mc_kl(z_k[k], att_stats.mu_k[k], att_stats.sigma_k[k], pmu[k - 1], psig[k - 1]))
            else:
                kl_m_k.append(H.mc_kl(z_k[k], att_stats.mu_k[k], att_stats.sigma_k[k]))
        losses = AttrDict(err=err, kl_m_k=kl_m_k)
        att_stats['pmu_k'], att_stats['psigma_k'] = pmu, psig
        log_m_k = list(log_m.unbind(0))
        x_r_k = list(x_r.unbind(0))
        with torch.no_grad():
            mx = x_r * log_m.exp()
        stats = AttrDict(recon=recon, log_m_k=log_m_k, log_s_k=list(log_s.unbind(0)), x_r_k=x_r_k, mx_r_k=list(mx.unbind(0)))
        if self.debug or not self.training:
            check_log_masks(log_m_k)
        return recon, losses, stats, att_stats, None

    @staticmethod
    def x_loss(x, log_m_k, x_r_k, std, pixel_wise=False):
        """Genesis.x_loss (reference genesis_config.py:273-286) on the fused kernel."""
        K = len(log_m_k)
        if not torch.is_tensor(std):
            std = torch.full((K,), float(std), device=x.device)
        std = std.reshape(-1).to(x.device).float()
        if std.numel() == 1:
            std = std.expand(K).contiguous()
        if pixel_wise:      # per pixel and channel, [B,3,H,W] (reference :283-284): the kernel's saved log-sum-exp, negated
            return ops.mixture_nll_pixelwise(x, torch.stack(list(x_r_k), 0), torch.stack(list(log_m_k), 0), std)
        err, _, _ = ops.mixture_nll(x, torch.stack(list(x_r_k), 0), torch.stack(list(log_m_k), 0), std, False)
        return err

    def sample(self, batch_size, K_steps=None):
        """Ancestral sampling (reference genesis_config.py:345-425) on the engine's decoders."""
        if self.K_steps == 1:
            raise NotImplementedError       # reference :346-347
        K = self.K_steps if K_steps is None else K_steps
        assert K == self.K_steps
        dev = self.std.device
        like = self.std
        with torch.no_grad():
            z_k = [self._normal((batch_size, self.ldim), like)]
            if self.autoreg_prior:
                state = None
                for _ in range(1, self.att_steps):
                    out, state = H.lstm_step(z_k[-1], state, self.prior_lstm)
                    lo = ops.linear(out, self.prior_linear.weight, self.prior_linear.bias)
                    mu = lo[:, :self.ldim]                      # reference :359 -- no tanh in sample()
                    sig = H.to_prior_sigma(lo[:, self.ldim:])
                    z_k.append(mu + sig * self._normal(mu.shape, like))
            else:
                z_k += [self._normal((batch_size, self.ldim), like) for _ in range(1, self.att_steps)]
            core = self.att_process.core
            if self.training and core.dec_norm == 'bn':
                # the reference decodes every z_m on its own (attention.py:61): BatchNorm batch statistics are per slot
                logits = torch.cat([H.sylvester_decode(core, z, True) for z in z_k], 0)
            else:
                logits = H.sylvester_decode(core, torch.cat(z_k, 0), self.training)
            logits = logits.view(K, batch_size, 1, self.img_size, self.img_size)
            log_m, log_s = ops.sbp_scan(logits, K)
            if not self.two_stage:                     # reference :404-409
                x_k = H.broadcast_decode(self.decoder, torch.cat(z_k, 0), 'elu', 3 if self.pixel_bound else 0)
                x_k = x_k.view(K, batch_size, -1, self.img_size, self.img_size)
                mx = x_k * log_m.exp()
                stats = AttrDict(x_k=list(x_k.unbind(0)), log_m_k=list(log_m.unbind(0)), log_s_k=list(log_s.unbind(0)),
                                 mx_k=list(mx.unbind(0)))
                return mx.sum(0), stats
            if self.comp_prior:
                pm = self.prior_mlp
                t = ops.linear(torch.cat(z_k, 0), pm[0].weight, pm[0].bias, 'elu')
                t = ops.linear(t, pm[2].weight, pm[2].bias, 'elu')
                a, b = torch.chunk(ops.linear(t, pm[4].weight, pm[4].bias), 2, dim=1)
                mu, sig = torch.tanh(a), H.to_prior_sigma(b)
                # one draw per slot, in slot order (reference :392-397)
                eps = torch.cat([self._normal((batch_size, mu.shape[1]), like) for _ in range(K)], 0)
                zc = mu + sig * eps
            else:
                zc = torch.cat([self._normal((batch_size, self.comp_vae.ldim), like) for _ in range(K)], 0)
            if self.comp_symmetric:
                x_k = H.symmetric_comp_decode(self.comp_vae, zc, self.training, 3 if self.pixel_bound else 0)
            else:
                x_k = H.broadcast_decode(self.comp_vae.decoder_module, zc, 'elu', 3 if self.pixel_bound else 0)
            x_k = x_k.view(K, batch_size, -1, self.img_size, self.img_size)
            mx = x_k * log_m.exp()
            img = mx.sum(0)
        stats = AttrDict(x_k=list(x_k.unbind(0)), log_m_k=list(log_m.unbind(0)),
                         log_s_k=list(log_s.unbind(0)), mx_k=list(mx.unbind(0)))
        return img, stats

    def get_features(self, image_batch):
        with torch.no_grad():
            _, _, _, att_stats, comp_stats = self.forward(image_batch)
        if att_stats is None:           # K_steps == 1: no mask latents
            return torch.cat(list(comp_stats['z_k']), dim=1)
        if comp_stats is None:          # one stage: no component latents
            return torch.cat(att_stats['z_k'][:self.K_steps - 1], dim=1)
        return torch.cat([*att_stats['z_k'][:self.K_steps - 1], *comp_stats['z_k']], dim=1)


def check_log_masks(log_m_k):
    """Invariant of reference utils/misc.py:258-270: masks sum to one within 1e-3, no NaNs."""
    summed = torch.stack(list(log_m_k), dim=4).exp().sum(dim=4)
    bad = torch.isnan(summed).any() | ((summed - 1.0).max() > 1e-3)
    if bool(bad):
        raise ValueError("Masks do not sum to 1.0. Not close enough.")
