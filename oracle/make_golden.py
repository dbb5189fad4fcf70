"""ORACLE / test infrastructure -- generate tests/golden/*.npz FROM THE REAL REFERENCE.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
For each case: seed the reference model (torch.manual_seed(0) -> default init), build a synthetic
batch (oracle/synth.py, seed 1), replay a NoiseTape (seed 2) through the reference's own forward,
differentiate the train.py loss (beta=1) and store inputs, noise, outputs, per-parameter checksums
and per-parameter gradient summaries.  Parameters themselves (up to 50 MB) are NOT stored: the
engine re-creates them from the same seed, and the checksums prove they are the same tensors."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import functional as O, models as M, ref_loader, synth  # noqa: E402

CASES = [  # name, model, K, img, B, generator
    ('genesis_k5_b2', 'genesis', 5, 64, 2, 'multid'),
    ('genesis_k3_b3', 'genesis', 3, 64, 3, 'rooms'),
    ('genesisv2_k7_b2', 'genesisv2', 7, 64, 2, 'stacks'),
    ('genesisv2_k4_b3', 'genesisv2', 4, 64, 3, 'rooms'),
    ('monet_k7_b2', 'monet', 7, 64, 2, 'multid'),
    ('monet128_k3_b1', 'monet', 3, 128, 1, 'multid'),
]
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def direction(n, idx):
    """Fixed pseudo-random probe vector for gradient projections."""
    i = torch.arange(n, dtype=torch.float64)
    return torch.cos(0.37 * i + 1.3 * idx + 0.1)


def param_checksums(sd):
    names = sorted(sd.keys())
    sums = np.array([[sd[k].double().sum().item(), sd[k].double().abs().sum().item()] for k in names])
    return names, sums


def ref_total_loss(losses):
    """train.py:227-259 with beta = 1."""
    tot = losses['err'].mean(0)
    for key in ('kl_l_k', 'kl_m_k'):
        if key in losses:
            tot = tot + torch.stack(list(losses[key]), 1).mean(0).sum()
    if 'kl_m' in losses and torch.is_tensor(losses['kl_m']):
        tot = tot + losses['kl_m'].mean(0)
    return tot


def run_case(name, model, K, img, B, gen, param_add=None, **over):
    """param_add: {parameter name: value added to the seeded initial value} -- used where the initial parameters never reach
    a branch (dynamic_K's early exit needs a wider IC-SBP kernel than the initial log_sigma); recorded in the golden."""
    cfg = M.make_cfg(model, K_steps=K, img_size=img, **over)
    ref = ref_loader.load_reference(model, cfg, seed=0)
    ref.train()
    if param_add:
        sd_ = dict(ref.named_parameters())
        with torch.no_grad():
            for k_, v_ in param_add.items():
                sd_[k_].add_(v_)
    names, sums = param_checksums(ref.state_dict())
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 1)[0])
    tape = O.NoiseTape(seed=2)
    with ref_loader.replay_noise(tape):
        recon, losses, stats, att, comp = ref(x)
    ref_total_loss(losses).backward()
    g = {}
    pnames = [k for k, _ in ref.named_parameters()]
    gsum = np.zeros((len(pnames), 2))
    for i, (k, p) in enumerate(ref.named_parameters()):
        if p.grad is not None:
            gd = p.grad.double().flatten()
            gsum[i] = [gd.norm().item(), (gd * direction(gd.numel(), i)).sum().item()]
    g['x'] = x.numpy()
    g['noise_kinds'] = np.array([k for k, _ in tape.record])
    for i, (_, t) in enumerate(tape.record):
        g['noise_%d' % i] = t.numpy()
    g['param_names'] = np.array(names)
    g['param_sums'] = sums
    g['grad_names'] = np.array(pnames)
    g['grad_sums'] = gsum
    g['recon'] = recon.detach().numpy()
    g['err'] = losses['err'].detach().numpy()
    for key in ('kl_l_k', 'kl_m_k'):
        if key in losses:
            g[key] = torch.stack(list(losses[key]), 0).detach().numpy()
    if 'kl_m' in losses and torch.is_tensor(losses['kl_m']):
        g['kl_m'] = losses['kl_m'].detach().numpy()
    g['log_m_k'] = torch.stack(list(stats['log_m_k']), 0).detach().numpy()
    if 'log_m_r_k' in stats:
        g['log_m_r_k'] = torch.stack(list(stats['log_m_r_k']), 0).detach().numpy()
    g['x_r_k'] = torch.stack(list(stats['x_r_k']), 0).detach().numpy().astype(np.float16)
    sd = ref.state_dict()
    bn = sorted(k for k in sd if k.endswith('running_mean') or k.endswith('running_var'))
    if bn:
        g['bn_names'] = np.array(bn)
        g['bn_sums'] = np.array([[sd[k].double().sum().item(), sd[k].double().abs().sum().item()] for k in bn])
    g['meta'] = np.array([model, str(K), str(img), str(B), gen])
    if over:
        g['overrides'] = np.array(['%s=%s' % kv for kv in sorted(over.items())])
    if param_add:
        g['param_add_names'] = np.array(sorted(param_add))
        g['param_add_values'] = np.array([param_add[k_] for k_ in sorted(param_add)], dtype=np.float64)
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **g)
    print(name, 'err', g['err'], '->', path, os.path.getsize(path) // 1024, 'KiB')


SAMPLE_CASES = [  # name, forward case that precedes it (same model instance: BatchNorm running stats), sample batch
    ('sample_genesis_k5', 'genesis_k5_b2', 3),
    ('sample_genesisv2_k7', 'genesisv2_k7_b2', 3),
    ('sample_monet_k7', 'monet_k7_b2', 2),
]


def run_sample_case(name, fwd_case, sb):
    """sample() of the real reference as its callers use it (train.py:425-472, scripts/compute_fid.py:104-125): one
    training-mode forward first (moves the BatchNorm running statistics off their initial values), then model.eval() and
    model.sample(batch) with a recorded noise tape (seed 7)."""
    _, model, K, img, B, gen = next(c for c in CASES if c[0] == fwd_case)
    cfg = M.make_cfg(model, K_steps=K, img_size=img)
    ref = ref_loader.load_reference(model, cfg, seed=0)
    ref.train()
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 1)[0])
    with ref_loader.replay_noise(O.NoiseTape(seed=2)):
        ref(x)
    ref.eval()
    tape = O.NoiseTape(seed=7)
    with torch.no_grad(), ref_loader.replay_noise(tape):
        image, stats = ref.sample(sb, K)
    g = {'meta': np.array([model, str(K), str(img), str(sb), fwd_case])}
    g['noise_kinds'] = np.array([k for k, _ in tape.record])
    for i, (_, t) in enumerate(tape.record):
        g['noise_%d' % i] = t.numpy()
    g['image'] = image.numpy()
    g['log_m_k'] = torch.stack(list(stats['log_m_k']), 0).numpy()
    g['x_k'] = torch.stack(list(stats['x_k']), 0).numpy().astype(np.float16)
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **g)
    print(name, 'image mean', float(image.mean()), '->', path, os.path.getsize(path) // 1024, 'KiB')


def run_vae_case(name='vae_b4', B=4, img=64, gen='multid', **over):
    """BaselineVAE (config c1, models/vae_config.py): forward + backward of err.mean + kl_l.mean (train.py:227-239)."""
    cfg = M.make_cfg('vae', K_steps=1, img_size=img, **over)
    ref = ref_loader.load_reference('vae', cfg, seed=0)
    ref.train()
    names, sums = param_checksums(ref.state_dict())
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 1)[0])
    tape = O.NoiseTape(seed=2)
    with ref_loader.replay_noise(tape):
        recon, losses, stats, _, _ = ref(x)
    (losses['err'].mean(0) + losses['kl_l'].mean(0)).backward()
    pnames = [k for k, _ in ref.named_parameters()]
    gsum = np.zeros((len(pnames), 2))
    for i, (k, p) in enumerate(ref.named_parameters()):
        if p.grad is not None:
            gd = p.grad.double().flatten()
            gsum[i] = [gd.norm().item(), (gd * direction(gd.numel(), i)).sum().item()]
    g = {'x': x.numpy(), 'meta': np.array(['vae', '1', str(img), str(B), gen])}
    g['noise_kinds'] = np.array([k for k, _ in tape.record])
    for i, (_, t) in enumerate(tape.record):
        g['noise_%d' % i] = t.numpy()
    g['param_names'], g['param_sums'] = np.array(names), sums
    g['grad_names'], g['grad_sums'] = np.array(pnames), gsum
    g['recon'] = recon.detach().numpy()
    g['err'] = losses['err'].detach().numpy()
    g['kl_l'] = losses['kl_l'].detach().numpy()
    g['z'] = stats['z'].detach().numpy()
    if over:
        g['overrides'] = np.array(['%s=%s' % kv for kv in sorted(over.items())])
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **g)
    print(name, 'err', g['err'], '->', path, os.path.getsize(path) // 1024, 'KiB')


EVAL_CASES = [('eval_genesis_k5', 'genesis_k5_b2'), ('eval_genesisv2_k7', 'genesisv2_k7_b2'), ('eval_monet_k7', 'monet_k7_b2')]


def run_eval_case(name, fwd_case):
    """model.eval() forward as in the reference's evaluation loop (train.py:479-546): one training-mode forward first (BatchNorm
    running statistics), then eval() + no_grad forward on a second batch with a recorded noise tape (seed 9)."""
    _, model, K, img, B, gen = next(c for c in CASES if c[0] == fwd_case)
    cfg = M.make_cfg(model, K_steps=K, img_size=img)
    ref = ref_loader.load_reference(model, cfg, seed=0)
    ref.train()
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 1)[0])
    with ref_loader.replay_noise(O.NoiseTape(seed=2)):
        ref(x)
    ref.eval()
    x2 = torch.from_numpy(synth.GENERATORS[gen](B, img, 3)[0])
    tape = O.NoiseTape(seed=9)
    with torch.no_grad(), ref_loader.replay_noise(tape):
        recon, losses, stats, att, comp = ref(x2)
    g = {'meta': np.array([model, str(K), str(img), str(B), fwd_case, gen]), 'x': x2.numpy()}
    g['noise_kinds'] = np.array([k for k, _ in tape.record])
    for i, (_, t) in enumerate(tape.record):
        g['noise_%d' % i] = t.numpy()
    g['recon'] = recon.numpy()
    g['err'] = losses['err'].numpy()
    for key in ('kl_l_k', 'kl_m_k'):
        if key in losses and len(losses[key]):
            g[key] = torch.stack(list(losses[key]), 0).numpy()
    if 'kl_m' in losses and torch.is_tensor(losses['kl_m']):
        g['kl_m'] = losses['kl_m'].numpy()
    g['log_m_k'] = torch.stack(list(stats['log_m_k']), 0).numpy()
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **g)
    print(name, 'err', g['err'], '->', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    if '--variants-dynamic' in sys.argv:    # GENESIS-V2 dynamic_K (genesisv2_config.py:118-137): batch (padded) and single image (truncated)
        import math
        run_case('variant_genesisv2_k6_dynamic_b3', 'genesisv2', 6, 64, 3, 'multid', dynamic_K=True,
                 param_add={'att_process.log_sigma': math.log(16.0)})
        run_case('variant_genesisv2_k8_dynamic_b1', 'genesisv2', 8, 64, 1, 'rooms', dynamic_K=True,
                 param_add={'att_process.log_sigma': math.log(32.0)})
        sys.exit(0)
    if '--variants-icsbp' in sys.argv:      # GENESIS-V2 attention options (modules/attention.py:138-160)
        run_case('variant_genesisv2_k4_laplacian', 'genesisv2', 4, 64, 2, 'multid', kernel='laplacian')
        run_case('variant_genesisv2_k4_epanechnikov', 'genesisv2', 4, 64, 2, 'rooms', kernel='epanechnikov')
        run_case('variant_genesisv2_k4_nosemiconv', 'genesisv2', 4, 64, 2, 'stacks', semiconv=False)
        sys.exit(0)
    if '--variants' in sys.argv:     # non-default model variants
        run_case('variant_genesis_k3_in', 'genesis', 3, 64, 3, 'multid', enc_norm='in', dec_norm='in')
        run_case('variant_genesis_k3_onestage', 'genesis', 3, 64, 2, 'multid', two_stage=False)
        run_case('variant_genesis_k3_nocompprior', 'genesis', 3, 64, 2, 'rooms', comp_prior=False)     # (autoreg_prior=False crashes in the reference itself: genesis_config.py:212 uses self.prior_lstm unconditionally)
        run_case('variant_genesis_k3_symmetric', 'genesis', 3, 64, 2, 'multid', comp_symmetric=True)
        run_case('variant_genesisv2_k4_klm', 'genesisv2', 4, 64, 2, 'stacks', klm_loss=True)
        run_case('variant_genesisv2_k4_klm_nodetach', 'genesisv2', 4, 64, 2, 'rooms', klm_loss=True, detach_mr_in_klm=False)
        run_case('variant_monet_k4_scope', 'monet', 4, 64, 2, 'multid', prior_mode='scope')
        run_case('variant_genesisv2_k4_noprior', 'genesisv2', 4, 64, 2, 'stacks', autoreg_prior=False)
        sys.exit(0)
    if '--vae' in sys.argv:
        run_vae_case()
        run_vae_case('vae_b2_broadcast', B=2, broadcast_decoder=True)
        sys.exit(0)
    if '--evals' in sys.argv:
        for case in EVAL_CASES:
            run_eval_case(*case)
        sys.exit(0)
    if not ref_loader.available():
        sys.exit('reference checkout not found at %s' % ref_loader.REF_ROOT)
    only_samples = '--samples' in sys.argv
    if not only_samples:
        for case in CASES:
            run_case(*case)
    for case in SAMPLE_CASES:
        run_sample_case(*case)
