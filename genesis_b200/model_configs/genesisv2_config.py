"""GENESIS-V2 plug-in: drop-in for the reference's models/genesisv2_config.py.

Same Forge contract (flags registered at import, `load(cfg)` -> nn.Module, `forward(x)` returning
`(recon, losses, stats, att_stats, comp_stats)`, `sample()`), same state_dict names (reference
models/genesisv2_config.py:35-46, 51-256).  UNet(GN) backbone, instance-colouring stick-breaking (one CTA per
image looping the K-1 steps), masked feature pooling, the conv-transpose decoder batched over all K*B slots and the
softmax-mask mixture likelihood run in hand-written sm_100a kernels (genesis_b200.ops).  `feat_head(enc_feat)` is
evaluated once instead of K times (identical values and summed gradient)."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import genesis_b200  # noqa: E402

try:
    from forge import flags
except ImportError:
    genesis_b200.enable_compat()
    from forge import flags
try:
    from attrdict import AttrDict
except ImportError:
    genesis_b200.enable_compat()
    from attrdict import AttrDict

from genesis_b200 import holders as H  # noqa: E402
from genesis_b200 import ops  # noqa: E402
from genesis_b200.model_configs import genesis_config as _g  # noqa: E402
from genesis_b200.model_configs import monet_config as _m  # noqa: E402,F401  (registers filter_start / prior_mode)

# reference models/genesisv2_config.py:35-42
flags.DEFINE_integer('feat_dim', 64, 'Number of features and latents.')
flags.DEFINE_string('kernel', 'gaussian', '{laplacian, gaussian, epanechnikov')
flags.DEFINE_boolean('semiconv', True, 'Use semi-convolutional embeddings.')
flags.DEFINE_boolean('dynamic_K', False, 'Dynamic K.')
flags.DEFINE_boolean('klm_loss', False, 'KL mask regulariser.')
flags.DEFINE_boolean('detach_mr_in_klm', True, 'Detach reconstructed masks.')


def load(cfg):
    return GenesisV2(cfg)


class ScalarGateHolder(nn.Module):
    """modules/blocks.py:85-90."""

    def __init__(self, init=0.0):
        super().__init__()
        self.gate = nn.Parameter(torch.tensor(init))


class SemiConvHolder(nn.Module):
    """modules/blocks.py:167-178 (conv 1x1 + scalar gate; the uv buffer is not persistent in the reference)."""

    def __init__(self, nin, nout):
        super().__init__()
        self.conv = nn.Conv2d(nin, nout, 1)
        self.gate = ScalarGateHolder()


class ICSBPHolder(nn.Module):
    """modules/attention.py:136-160."""

    def __init__(self, K_steps, feat_dim, colour_dim=8, kernel='gaussian', semiconv=True):
        super().__init__()
        self.colour_dim = colour_dim
        self.kernel = kernel
        # numpy float64 -> float64 parameter, exactly as the reference (modules/attention.py:146-155)
        if kernel == 'laplacian':
            sigma_init = 1.0 / (np.sqrt(K_steps) * np.log(2))
        elif kernel == 'gaussian':
            sigma_init = 1.0 / (K_steps * np.log(2))
        elif kernel == 'epanechnikov':
            sigma_init = 2.0 / K_steps
        else:
            raise ValueError("No valid kernel.")
        self.log_sigma = nn.Parameter(torch.tensor(sigma_init).log())
        # modules/attention.py:157-160: SemiConv (1x1 conv, scalar gate, pixel coordinates added to the last two channels)
        # or a plain 1x1 conv
        self.colour_head = SemiConvHolder(feat_dim, colour_dim) if semiconv else nn.Conv2d(feat_dim, colour_dim, 1)


class GenesisV2(nn.Module, _g.NoiseMixin):

    def __init__(self, cfg):
        super().__init__()
        self.K_steps = cfg.K_steps
        self.pixel_bound = cfg.pixel_bound
        self.feat_dim = cfg.feat_dim
        self.klm_loss = cfg.klm_loss
        self.detach_mr_in_klm = cfg.detach_mr_in_klm
        self.dynamic_K = cfg.dynamic_K
        self.debug = cfg.debug
        self.multi_gpu = cfg.multi_gpu
        self.img_size = cfg.img_size
        self.semiconv = bool(cfg.semiconv)
        self.precise_backbone = True    # 3xTF32 in the UNet backbone + seg / feat heads (fp32-level accuracy on the tensor cores)
        self.precise_decoder = False    # the same for the conv-transpose decoder (experiments)
        if cfg.feat_dim != 64:
            raise NotImplementedError('engine is tiled for feat_dim=64')
        c = cfg.feat_dim
        # construction order == reference (genesisv2_config.py:63-105) so seeded init is identical
        self.encoder = H.UNetHolder(int(math.log2(cfg.img_size) - 1), cfg.img_size, min(c, 64), 3, c, norm='gn')
        self.encoder.final_conv = nn.Identity()
        self.att_process = ICSBPHolder(self.K_steps, c, kernel=cfg.kernel, semiconv=self.semiconv)
        self.seg_head = H._conv_block(c, c, 'gn')
        self.feat_head = nn.Sequential(H._conv_block(c, c, 'gn'), nn.Conv2d(c, 2 * c, 1))
        self.z_head = nn.Sequential(nn.LayerNorm(2 * c), nn.Linear(2 * c, 2 * c), nn.Identity(), nn.Linear(2 * c, 2 * c))
        cm = min(c, 64)
        self.decoder_module = nn.Sequential(
            nn.Identity(),
            nn.ConvTranspose2d(c + 2, c, 5, 2, 2, 1), nn.GroupNorm(8, c), nn.Identity(),
            nn.ConvTranspose2d(c, c, 5, 2, 2, 1), nn.GroupNorm(8, c), nn.Identity(),
            nn.ConvTranspose2d(c, cm, 5, 2, 2, 1), nn.GroupNorm(8, cm), nn.Identity(),
            nn.ConvTranspose2d(cm, cm, 5, 2, 2, 1), nn.GroupNorm(8, cm), nn.Identity(),
            nn.Conv2d(cm, 4, 1))
        self.autoreg_prior = cfg.autoreg_prior
        self.prior_lstm, self.prior_linear = None, None
        if self.autoreg_prior and self.K_steps > 1:
            self.prior_lstm = nn.LSTM(c, 4 * c)
            self.prior_linear = nn.Linear(4 * c, 2 * c)
        assert cfg.pixel_std1 == cfg.pixel_std2
        self.std = cfg.pixel_std1

    # --------------------------------------------------------------------------------------------
    def _decode(self, z):
        """decoder_module on all K*B slot latents at once (reference genesisv2_config.py:88-99, 205-211: GroupNorm is
        per sample, so batching the K sequential calls is exact) -> [K*B,4,H,W] NCHW, RGB planes after the sigmoid."""
        dm = self.decoder_module
        N = z.shape[0]
        d = self.img_size // 16
        cpad = 128
        h = torch.cat([z.view(N, 1, 1, -1).expand(N, d, d, z.shape[1]),
                       H.coords_nhwc(d, z.device).expand(N, d, d, 2),
                       z.new_zeros(N, d, d, cpad - z.shape[1] - 2)], dim=3).contiguous()
        for ci in (1, 4, 7, 10):
            conv, gn = dm[ci], dm[ci + 1]
            # the first layer (4x4 -> 8x8 maps, a few MFLOP) is always a precise layer with the backbone: its weight gradient is
            # the one decoder tensor that plain TF32 leaves near the 2e-2 bar; `precise_decoder` extends that to all four
            with ops.precise(self.precise_decoder or (ci == 1 and self.precise_backbone)):
                y = ops.conv_transpose2d(h, conv.weight, conv.bias, 2, 2, bias_grad=False)     # the op zero-pads the weight to h's channels
            h = ops.norm_post(y, gn.weight, gn.bias, mode=ops.NORM_GROUP, groups=gn.num_groups, post=ops.POST_RELU, eps=gn.eps,
                              conv_bias=conv.bias)
        return ops.out1x1(h, dm[13].weight, dm[13].bias, 3 if self.pixel_bound else 0)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('genesis_b200 runs on CUDA (sm_100a) only; there is no CPU path')
        K, B = self.K_steps, x.shape[0]
        x = x.contiguous().float()
        S = self.img_size
        tc = ops.get_precision() == 'tf32'
        xh = ops.to_nhwc_padded(x, 32) if tc else ops.to_nhwc(x)
        # the GroupNorm UNet backbone and the two heads that read it run as 3xTF32 (ops.precise): plain TF32 operand rounding is
        # amplified by the per-sample norms into 5-15 % errors on their gradients (profiles/r01_parity_report.txt)
        with ops.precise(self.precise_backbone):
            enc_feat = H.unet_forward(self.encoder, xh)        # ends in ReLU: F.relu(enc_feat) (:115) is the identity
            # --- attention masks (reference attention.py:162-226)
            seg = H.conv_norm_relu(self.seg_head, enc_feat, 'gn')
        ap = self.att_process
        ch = ap.colour_head
        cd = ap.colour_dim
        if self.semiconv:
            out = ops.conv2d(seg, ch.conv.weight, ch.conv.bias, 1, 0) * ch.gate.gate         # [B,H,W,8] NHWC
            uv = torch.cat([out.new_zeros(1, S, S, cd - 2), H.coords_nhwc(S, x.device)], dim=3)
            colour = out + uv
        else:                                   # plain 1x1 colour head: no coordinate offsets, delta = None (attention.py:174-178)
            out = None
            colour = ops.conv2d(seg, ch.weight, ch.bias, 1, 0)
        n_seeds = K - 1
        if self.dynamic_K:
            # reference :118-137: the attention runs image by image (one [1,1,H,W] uniform draw each) and stops early once a
            # mask would hold fewer than 20 pixels; batches are padded with -1e10 masks up to K_steps, a single image keeps
            # only the masks it produced (the number of slots then decides every shape downstream: one host sync)
            u = torch.cat([self._uniform((1, 1, S, S), x) for _ in range(B)], 0)
            log_m, log_s, seed_idx, n_masks = ops.icsbp_dynamic(colour, u, ap.log_sigma, K, ap.kernel)
            if B == 1:
                k_eff = int(n_masks.item())
                n_seeds = min(k_eff, K - 1)          # the seed of the step that stopped the loop is still recorded (attention.py:193)
                K = k_eff
                log_m, log_s = log_m[:K], log_s[:K]
        else:
            u = self._uniform((B, 1, S, S), x)
            log_m, log_s, seed_idx = ops.icsbp(colour, u, ap.log_sigma, K, ap.kernel)        # [K,B,1,H,W]
        # --- slot latents (reference genesisv2_config.py:145-161), feat_head evaluated once
        with ops.precise(self.precise_backbone):
            f = H.conv_norm_relu(self.feat_head[0], enc_feat, 'gn')
            f = ops.conv2d(f, self.feat_head[1].weight, self.feat_head[1].bias, 1, 0)        # [B,H,W,128]
        num, msum = ops.masked_pool(f, log_m)                                                # [K,B,128], [K,B]
        obj = (num / (msum.unsqueeze(2) + 1e-5)).reshape(K * B, -1)
        zh = self.z_head
        # z_head: two [K*B,128] x [128,128] products whose BACKWARD carries the whole attention gradient (masks -> pooled features
        # -> latents -> decoder): 3xTF32 in both directions (measured: colour-head gate gradient 5.8e-2 -> see DESIGN.md section 5)
        with ops.precise(self.precise_backbone, backward=True):
            t = H.layer_norm(obj, zh[0])
            t = ops.linear(t, zh[1].weight, zh[1].bias, 'relu')
            lo = ops.linear(t, zh[3].weight, zh[3].bias)
        # the reference draws one [B,64] normal per slot, in slot order (:156-157)
        eps = torch.cat([self._normal((B, lo.shape[1] // 2), x) for _ in range(K)], 0)
        z, mu, sigma = H.gauss_head(lo, eps)
        z_k = list(torch.chunk(z, K, 0))
        mu_k, sigma_k = list(torch.chunk(mu, K, 0)), list(torch.chunk(sigma, K, 0))
        # --- KL (reference genesis_config.py:288-343): the prior LSTM + MC-KL sums are ~200 tiny kernels that do not feed the
        # decoder, so they run on a side stream beside it (same scheme as genesis_config.py; backward follows automatically)
        cur = torch.cuda.current_stream()
        side = ops.side_stream(x.device) if (ops.side_streams_enabled() and torch.is_grad_enabled()) else cur
        if side is not cur:
            side.wait_stream(cur)
            for t_ in (z, mu, sigma):
                t_.record_stream(side)
        with torch.cuda.stream(side):
            if self.autoreg_prior and K > 1:
                pmu, psig = H.autoreg_prior(z_k, self.prior_lstm, self.prior_linear)
            else:
                pmu, psig = [], []
            kl = [H.mc_kl(z_k[0], mu_k[0], sigma_k[0])]
            for k in range(1, K):
                if pmu:
                    kl.append(H.mc_kl(z_k[k], mu_k[k], sigma_k[k], pmu[k - 1], psig[k - 1]))
                else:
                    kl.append(H.mc_kl(z_k[k], mu_k[k], sigma_k[k]))
        # --- decode + softmax-mask mixture likelihood (reference :164-169, 205-225)
        dec = self._decode(z).view(K, B, 4, S, S)
        std = torch.full((K,), float(self.std), device=x.device)
        err, recon, log_m_r = ops.mixture_nll_packed(x, dec, None, std, True)
        if side is not cur:
            cur.wait_stream(side)
            for t_ in kl + pmu + psig:
                t_.record_stream(cur)
        losses = AttrDict()
        losses['err'] = err
        if self.klm_loss:           # reference :171-176: KL(masks || reconstructed masks)
            losses['kl_m'] = ops.mask_kl(log_m, dec, self.detach_mr_in_klm)
        losses['kl_l_k'] = kl
        # --- tracking
        log_m_k = list(log_m.unbind(0))
        log_m_r_k = list(log_m_r.unbind(0))
        x_r_k = [dec[k, :, :3] for k in range(K)]
        with torch.no_grad():
            mx_r_k = [x_r_k[k] * log_m_r_k[k].exp() for k in range(K)]
            inst = torch.argmax(log_m.squeeze(2).permute(1, 0, 2, 3), dim=1)
            inst_r = torch.argmax(log_m_r.squeeze(2).permute(1, 0, 2, 3), dim=1)
            colour_nchw = colour.permute(0, 3, 1, 2)
            flat = colour.reshape(B, S * S, cd)
            seeds = [flat[torch.arange(B, device=x.device), seed_idx[k].long().clamp_min(0)] for k in range(n_seeds)]
        batched_dynamic = self.dynamic_K and B > 1          # reference :121-122: log_s_k and att_stats are None on that path
        stats = AttrDict(recon=recon, log_m_k=log_m_k, log_s_k=(None if batched_dynamic else list(log_s.unbind(0))), x_r_k=x_r_k,
                         log_m_r_k=log_m_r_k, mx_r_k=mx_r_k, instance_seg=inst, instance_seg_r=inst_r)
        att_stats = None if batched_dynamic else AttrDict(
            colour=colour_nchw, delta=(out.permute(0, 3, 1, 2)[:, -2:] if out is not None else None), seeds=seeds,
            seed_idx=seed_idx[:n_seeds])
        comp_stats = AttrDict(mu_k=mu_k, sigma_k=sigma_k, z_k=z_k, kl_l_k=[], pmu_k=pmu, psigma_k=psig)
        if self.debug:
            _g.check_log_masks(log_m_k)
            _g.check_log_masks(log_m_r_k)
        return recon, losses, stats, att_stats, comp_stats

    def sample(self, batch_size, K_steps=None):
        """reference genesisv2_config.py:227-256."""
        K = self.K_steps if K_steps is None else K_steps
        like = self.att_process.log_sigma
        S = self.img_size
        with torch.no_grad():
            z_k = [self._normal((batch_size, self.feat_dim), like)]
            if self.autoreg_prior and self.prior_lstm is not None:
                state = None
                for _ in range(1, K):
                    out, state = H.lstm_step(z_k[-1], state, self.prior_lstm)
                    lo = ops.linear(out, self.prior_linear.weight, self.prior_linear.bias)
                    mu = torch.tanh(lo[:, :self.feat_dim])
                    sig = H.to_prior_sigma(lo[:, self.feat_dim:])
                    z_k.append(mu + sig * self._normal(mu.shape, like))
            else:
                z_k += [self._normal((batch_size, self.feat_dim), like) for _ in range(1, K)]
            dec = self._decode(torch.cat(z_k, 0)).view(K, batch_size, 4, S, S)
            log_m = torch.log_softmax(dec[:, :, 3:], dim=0)
            x_k = dec[:, :, :3]
            mx = x_k * log_m.exp()
            recon = mx.sum(0)
        stats = AttrDict(x_k=list(x_k.unbind(0)), log_m_k=list(log_m.unbind(0)), mx_k=list(mx.unbind(0)))
        return recon, stats
