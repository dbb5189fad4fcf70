"""MONet plug-in: drop-in for the reference's models/monet_config.py.

Same Forge contract (flags registered at import, `load(cfg)` -> nn.Module, `forward(x)` returning
`(recon, losses, stats, att_stats, comp_stats)`, `sample()`), same state_dict names (reference
models/monet_config.py:36-41, 46-128).  The K-1 autoregressive UNet passes (modules/attention.py:31-51), the
component VAE and the loss head run in hand-written sm_100a kernels (genesis_b200.ops)."""
import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

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
from genesis_b200.model_configs import genesis_config as _g  # noqa: E402,F401  (registers the shared flags)

# reference models/monet_config.py:36-37
flags.DEFINE_integer('filter_start', 32, 'Starting number of channels in UNet.')
flags.DEFINE_string('prior_mode', 'softmax', '{scope, softmax}')


def load(cfg):
    return MONet(cfg)


class SimpleSBPHolder(nn.Module):
    """Holder for modules/attention.py:25-29."""

    def __init__(self, core):
        super().__init__()
        self.core = core


class MONet(nn.Module, _g.NoiseMixin):

    def __init__(self, cfg):
        super().__init__()
        self.K_steps = cfg.K_steps
        self.prior_mode = cfg.prior_mode
        self.mckl = cfg.montecarlo_kl
        self.debug = cfg.debug
        self.pixel_bound = cfg.pixel_bound
        self.img_size = cfg.img_size
        if self.prior_mode not in ('softmax', 'scope'):
            raise ValueError("No valid prior mode.")         # reference monet_config.py:155
        if not hasattr(cfg, 'filter_start'):
            cfg['filter_start'] = 32
        core = H.UNetHolder(int(math.log2(cfg.img_size) - 1), cfg.img_size, cfg.filter_start, 4, 1, norm='in')
        self.att_process = SimpleSBPHolder(core)
        self.comp_vae = H.ComponentVAEHolder(nout=4, cfg=cfg)
        self.comp_vae.pixel_bound = False
        std = cfg.pixel_std2 * torch.ones(1, 1, 1, 1, self.K_steps)
        std[0, 0, 0, 0, 0] = cfg.pixel_std1
        self.register_buffer('std', std)
        self.precise_unet = True        # 3xTF32 in the attention UNet (fp32-level accuracy on the tensor cores); False = plain TF32
        self.precise_comp_encoder = True

    def _attention(self, x):
        """SimpleSBP.forward (reference attention.py:31-51): K-1 sequential UNet passes on cat(x, log_s_k)."""
        K, B = self.K_steps, x.shape[0]
        core = self.att_process.core
        HW = (self.img_size, self.img_size)
        cp = 32 if ops.get_precision() == 'tf32' else 4
        w0 = core.down[0][0].weight
        log_s = torch.zeros(B, 1, *HW, device=x.device)
        log_m_k, log_s_k = [], [log_s]
        # the packed input carries log_s in channel 0 and x in 1..3: present the first conv's weight in that order
        core_w0 = torch.cat([w0[:, 3:4], w0[:, :3]], dim=1)
        for _ in range(K - 1):
            h = ops.comp_pack(x, log_s_k[-1].view(1, B, 1, *HW), cp)
            # K-1 recurrent passes through per-sample norms amplify operand rounding: the UNet runs as 3xTF32 (ops.precise)
            with ops.precise(self.precise_unet):
                h = H.unet_forward(_FirstWeight(core, core_w0), h)
            a = ops.out1x1(h, core.final_conv.weight[:1], core.final_conv.bias[:1], 0)       # core_out[:, :1]
            log_m_k.append(log_s_k[-1] + F.logsigmoid(a))
            log_s_k.append(log_s_k[-1] + F.logsigmoid(-a))
        log_m_k.append(log_s_k[-1])
        return log_m_k, log_s_k

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('genesis_b200 runs on CUDA (sm_100a) only; there is no CPU path')
        K, B = self.K_steps, x.shape[0]
        x = x.contiguous().float()
        log_m_k, log_s_k = self._attention(x)
        log_m = torch.stack(log_m_k, 0)                                     # [K,B,1,H,W]
        cv = self.comp_vae
        with ops.precise(self.precise_comp_encoder):    # log-mask inputs down to -9: TF32 rounding of the operand alone is 4e-3 absolute
            enc = H.comp_encode(cv.encoder_module, ops.comp_pack(x, log_m, 32 if ops.get_precision() == 'tf32' else 4), 'relu')
        cz, cmu, csig = H.gauss_head(enc, self._normal((enc.shape[0], enc.shape[1] // 2), x))
        dec = H.broadcast_decode(cv.decoder_module, cz, 'relu', 3 if self.pixel_bound else 0)
        dec = dec.view(K, B, 4, self.img_size, self.img_size)
        if self.prior_mode == 'scope':
            # reference monet_config.py:141-153: reconstructed masks by stick-breaking over the decoder's mask logits
            err, recon, _ = ops.mixture_nll_packed(x, dec, log_m, self.std.reshape(-1), False)
            log_m_r, _ = ops.sbp_scan(dec[:, :, 3:4].contiguous(), K)
            kl_m = ops.mask_kl(log_m, log_m_r, False)
            log_m_r = log_m_r.detach()
        else:
            err, kl_m, recon, log_m_r = ops.monet_loss(x, dec, log_m, self.std.reshape(-1))
        losses = AttrDict()
        losses['err'] = err
        losses['kl_m'] = kl_m
        kl = H.mc_kl(cz, cmu, csig)                                         # vs N(0,1), reference monet_config.py:109-113
        losses['kl_l_k'] = list(torch.chunk(kl, K, 0))
        x_r_k = [dec[k, :, :3] for k in range(K)]
        log_m_r_k = list(log_m_r.unbind(0))
        with torch.no_grad():
            mx_r_k = [x_r_k[k] * log_m_k[k].exp() for k in range(K)]
        stats = AttrDict(recon=recon, log_m_k=log_m_k, log_s_k=log_s_k, x_r_k=x_r_k, log_m_r_k=log_m_r_k, mx_r_k=mx_r_k)
        comp_stats = AttrDict(mu_k=list(torch.chunk(cmu, K, 0)), sigma_k=list(torch.chunk(csig, K, 0)),
                              z_k=list(torch.chunk(cz, K, 0)))
        if self.debug:
            assert len(log_m_k) == self.K_steps
            _g.check_log_masks(log_m_k)
            _g.check_log_masks(log_m_r_k)
        return recon, losses, stats, {}, comp_stats

    def get_features(self, image_batch):
        with torch.no_grad():
            _, _, _, _, comp_stats = self.forward(image_batch)
            return torch.cat(comp_stats.z_k, dim=1)

    def sample(self, batch_size, K_steps=None):
        """reference monet_config.py:172-198: z ~ N(0,1) for all K slots, decode, softmax masks, composite."""
        K = self.K_steps if K_steps is None else K_steps
        like = self.std
        with torch.no_grad():
            z = self._normal((batch_size * K, self.comp_vae.ldim), like)
            dec = H.broadcast_decode(self.comp_vae.decoder_module, z, 'relu', 3 if self.pixel_bound else 0)
            dec = dec.view(K, batch_size, 4, self.img_size, self.img_size)
            if self.prior_mode == 'scope':
                log_m, _ = ops.sbp_scan(dec[:, :, 3:4].contiguous(), K)
            else:
                log_m = F.log_softmax(dec[:, :, 3:], dim=0)
            x_k = dec[:, :, :3]
            mx = x_k * log_m.exp()
            img = mx.sum(0)
        stats = AttrDict(gen_image=img, x_k=list(x_k.unbind(0)), log_m_k=list(log_m.unbind(0)), mx_k=list(mx.unbind(0)))
        return img, stats


class _FirstWeight(object):
    """View of a UNet holder whose first down-conv weight is replaced (channel-reordered) for this pass."""

    def __init__(self, unet, w0):
        self._u = unet
        first = _Block(w0, unet.down[0][1])
        self.down = [first] + [unet.down[i] for i in range(1, unet.num_blocks)]
        self.up, self.mlp, self.num_blocks, self.norm = unet.up, unet.mlp, unet.num_blocks, unet.norm


class _Block(object):
    def __init__(self, w, norm):
        self._items = (_W(w), norm)

    def __getitem__(self, i):
        return self._items[i]


class _W(object):
    def __init__(self, w):
        self.weight = w
