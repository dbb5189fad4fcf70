"""BaselineVAE plug-in (config c1): drop-in for the reference's models/vae_config.py.

Same Forge contract (flags of vae_config.py:26-31, `load(cfg)`, forward -> (recon, losses{err, kl_l}, stats{x, mu, sigma, z},
None, None), sample(), get_features()) and the same state_dict names (`vae.q_z_nn.*`, `vae.q_z_mean`, `vae.q_z_var.0`,
`vae.p_x_nn.*`, `vae.p_x_mean`), on the engine's kernels: the gated conv encoder / conv-transpose decoder are the ones
GENESIS uses for its attention core (holders.sylvester_encode / sylvester_decode), the Gaussian pixel likelihood is the
mixture kernel with one slot and a zero log-mask.  Both decoder variants (`broadcast_decoder` False / True, vae_config.py:53-61).
Validated on a B200 in round 2 (tests/test_extras_gpu.py, tests/test_train_py_dropin.py)."""
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
from genesis_b200.noise import NoiseMixin  # noqa: E402  (not via genesis_config: that would register its flags, e.g. a second `pixel_bound`)

# Flag names / defaults: reference models/vae_config.py:26-31
flags.DEFINE_integer('latent_dimension', 64, 'Latent channels.')
flags.DEFINE_boolean('broadcast_decoder', False, 'Use broadcast decoder instead of deconv.')
flags.DEFINE_boolean('pixel_bound', True, 'Bound pixel values to [0, 1].')
flags.DEFINE_float('pixel_std', 0.7, 'StdDev of reconstructed pixels.')


def load(cfg):
    return BaselineVAE(cfg)


class BaselineVAE(nn.Module, NoiseMixin):

    def __init__(self, cfg):
        super().__init__()
        cfg.K_steps = None                                   # reference vae_config.py:44
        # train.py:463 reads `model.K_steps` for every model; the reference's BaselineVAE does not define it, so its own
        # train.py dies in visualise_outputs() at iteration 0.  The plug-in carries the attribute (one slot) so the caller runs.
        self.K_steps = 1
        self.ldim = cfg.latent_dimension
        self.pixel_std = cfg.pixel_std
        self.pixel_bound = cfg.pixel_bound
        self.debug = cfg.debug
        self.img_size = cfg.img_size
        self.broadcast_decoder = bool(getattr(cfg, 'broadcast_decoder', False))
        nin = cfg.input_channels if hasattr(cfg, 'input_channels') else 3
        self.vae = H.SylvesterVAE(self.ldim, [nin, cfg.img_size, cfg.img_size], nin)
        if self.broadcast_decoder:          # reference vae_config.py:53-61 (replaces the modules AFTER the full VAE was built)
            self.vae.p_x_nn = nn.Sequential(nn.Identity(), H.BroadcastDecoderHolder(self.ldim, 64, 64, 4, cfg.img_size),
                                            nn.Identity())
            self.vae.p_x_mean = nn.Conv2d(64, nin, 1, 1, 0)
        self.register_buffer('_std', torch.full((1,), float(cfg.pixel_std)), persistent=False)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('genesis_b200 runs on CUDA (sm_100a) only; there is no CPU path')
        x = x.contiguous().float()
        B = x.shape[0]
        core = self.vae
        tc = ops.get_precision() == 'tf32'
        xh = ops.to_nhwc_padded(x, 32) if tc else ops.to_nhwc(x)
        h = H.sylvester_encode(core, xh, self.training)                                   # [B,256]
        wmv = torch.cat([core.q_z_mean.weight, core.q_z_var[0].weight], 0)
        bmv = torch.cat([core.q_z_mean.bias, core.q_z_var[0].bias], 0)
        z, mu, sigma = H.gauss_head(ops.linear(h, wmv, bmv), self._normal((B, self.ldim), x))
        recon = self._decode(z)                                                            # NCHW [B,nin,H,W]
        # -log N(x; recon, std) summed over pixels = the mixture likelihood with one slot and log m = 0
        log_m = torch.zeros(1, B, 1, self.img_size, self.img_size, device=x.device)
        err, _, _ = ops.mixture_nll(x, recon.unsqueeze(0), log_m, self._std, False)
        kl = H.mc_kl(z, mu, sigma)
        losses = AttrDict(err=err, kl_l=kl)
        stats = AttrDict(x=recon, mu=mu, sigma=sigma, z=z)
        return recon, losses, stats, None, None

    def _decode(self, z):
        nsig = 3 if self.pixel_bound else 0
        if self.broadcast_decoder:
            h = H.broadcast_decode(self.vae.p_x_nn[1], z, 'elu', 0, head_act='elu')       # NHWC [B,H,W,64]
            return ops.out1x1(h, self.vae.p_x_mean.weight, self.vae.p_x_mean.bias, nsig)
        return H.sylvester_decode(self.vae, z, self.training, nsig)

    def sample(self, batch_size, *args, **kwargs):
        with torch.no_grad():
            z = self._normal((batch_size, self.ldim), self._std)
            x = self._decode(z)
        return x, AttrDict(z=z)

    def get_features(self, image_batch):
        with torch.no_grad():
            _, _, stats, _, _ = self.forward(image_batch)
        return stats.z
