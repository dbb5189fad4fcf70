"""Size-independent properties of the hot path, usable at BASELINE.json's full sizes where the CPU oracle would take minutes:

* identities that tie the engine's outputs to each other in float64 -- masks sum to one over the K slots; recon is the
  mask-weighted sum of the slot reconstructions; err is the mixture likelihood of (x, masks, slot reconstructions); every KL
  term is the Monte-Carlo KL of the returned (z, mu, sigma, prior) -- so a fused kernel cannot drift from the quantities it
  reports next to its loss;
* batch-partition invariance: with per-sample norms (GENESIS-V2, MONet) or BatchNorm in eval mode
  (GENESIS) the result for image b does not depend on the other images of the batch, given the same noise for image b.

The helpers are exercised on the CPU at small sizes through the op-contract stand-ins (tests/test_full_size_props_cpu.py);
the full-size runs are GPU tests."""
import torch

from oracle import functional as O


class SubsetTape(object):
    """Noise source whose draws for image b are the same whatever the batch: every draw is generated for `full_B` images and
    the first `sub_B` are returned.  Leading dimension B (per-step latents, IC-SBP uniforms) or K*B in k-major order
    (component latents)."""

    def __init__(self, seed, full_B, sub_B, K):
        self.gen = torch.Generator(device='cpu')
        self.gen.manual_seed(int(seed))
        self.full_B, self.sub_B, self.K = full_B, sub_B, K

    def _draw(self, fn, shape):
        shape = tuple(int(s) for s in shape)
        if shape[0] == self.sub_B:
            return fn((self.full_B,) + shape[1:])[:self.sub_B].clone()
        if shape[0] == self.K * self.sub_B:
            return fn((self.K, self.full_B) + shape[1:])[:, :self.sub_B].reshape(shape).clone()
        raise AssertionError('unexpected noise shape %r for B=%d K=%d' % (shape, self.sub_B, self.K))

    def normal(self, shape, dtype=torch.float32):
        return self._draw(lambda s: torch.randn(s, generator=self.gen, dtype=torch.float32), shape).to(dtype)

    def uniform(self, shape, dtype=torch.float32):
        return self._draw(lambda s: torch.rand(s, generator=self.gen, dtype=torch.float32), shape).to(dtype)


class RowTape(SubsetTape):
    """As SubsetTape, for an arbitrary subset of rows of the full batch (a data-parallel rank's shard)."""

    def __init__(self, seed, full_B, rows, K):
        super().__init__(seed, full_B, len(rows), K)
        self.rows = torch.as_tensor(list(rows), dtype=torch.long)

    def _draw(self, fn, shape):
        shape = tuple(int(s) for s in shape)
        if shape[0] == self.sub_B:
            return fn((self.full_B,) + shape[1:])[self.rows].clone()
        if shape[0] == self.K * self.sub_B:
            return fn((self.K, self.full_B) + shape[1:])[:, self.rows].reshape(shape).clone()
        raise AssertionError('unexpected noise shape %r for B=%d K=%d' % (shape, self.sub_B, self.K))


def _d(t):
    return t.detach().double().cpu()


def _stack(ts):
    return torch.stack([_d(t) for t in ts], 0)


def check_identities(model_name, x, out, std, tol=1.0):
    """out = (recon, losses, stats, att_stats, comp_stats) of one forward; std: python float or [K] tensor.
    `tol` scales the tolerances (1.0 = fp32 storage of float64-exact identities)."""
    recon, losses, stats, att, comp = out
    x = _d(x)
    log_m = _stack(stats['log_m_k'])                                     # [K,B,1,H,W]
    x_r = _stack(stats['x_r_k'])                                         # [K,B,3,H,W]
    K = log_m.shape[0]
    # masks are a partition of unity (utils/misc.py:check_log_masks uses 1e-3)
    assert (log_m.exp().sum(0) - 1).abs().max().item() < 1e-4 * tol
    used = log_m
    if 'log_m_r_k' in stats and model_name == 'genesisv2':               # V2 composes with the RECONSTRUCTED masks (:164-169, 221-223)
        used = _stack(stats['log_m_r_k'])
    if 'log_m_r_k' in stats:
        assert (_stack(stats['log_m_r_k']).exp().sum(0) - 1).abs().max().item() < 1e-4 * tol
    torch.testing.assert_close(_d(recon), (used.exp() * x_r).sum(0), rtol=0, atol=2e-5 * tol)
    std_t = (std.detach().double().cpu() if torch.is_tensor(std) else torch.as_tensor(std, dtype=torch.float64)).reshape(-1)
    std_arg = float(std_t[0]) if std_t.numel() == 1 else std_t
    err = O.mixture_nll(x, list(used.unbind(0)), list(x_r.unbind(0)), std_arg)
    torch.testing.assert_close(_d(losses['err']), err, rtol=2e-5 * tol, atol=1e-2 * tol)
    # Monte-Carlo KL terms from the returned statistics
    def kl_of(stat, k, pm, ps):
        return O.mc_kl(_d(stat['z_k'][k]), _d(stat['mu_k'][k]), _d(stat['sigma_k'][k]),
                       None if pm is None else _d(pm), None if ps is None else _d(ps))
    if model_name == 'genesis':
        pmu, psig = att['pmu_k'], att['psigma_k']                        # K-1 entries: priors of steps 1..K-1
        for k in range(K):
            ref = kl_of(att, k, pmu[k - 1] if k else None, psig[k - 1] if k else None)
            torch.testing.assert_close(_d(losses['kl_m_k'][k]), ref, rtol=1e-4 * tol, atol=2e-3 * tol)
        if 'pmu_k' in comp:
            for k in range(K):
                ref = kl_of(comp, k, comp['pmu_k'][k], comp['psigma_k'][k])
                torch.testing.assert_close(_d(losses['kl_l_k'][k]), ref, rtol=1e-4 * tol, atol=2e-3 * tol)
    elif model_name == 'genesisv2':
        pmu, psig = comp['pmu_k'], comp['psigma_k']
        for k in range(K):
            has = k > 0 and len(pmu) >= k
            ref = kl_of(comp, k, pmu[k - 1] if has else None, psig[k - 1] if has else None)
            torch.testing.assert_close(_d(losses['kl_l_k'][k]), ref, rtol=1e-4 * tol, atol=2e-3 * tol)
    else:
        for k in range(K):
            torch.testing.assert_close(_d(losses['kl_l_k'][k]), kl_of(comp, k, None, None), rtol=1e-4 * tol, atol=2e-3 * tol)


def check_subset_invariance(full, sub, n, rtol, atol):
    """full / sub: forward outputs for a batch and for its first n images (same per-image noise)."""
    torch.testing.assert_close(_d(sub[1]['err']), _d(full[1]['err'])[:n], rtol=rtol, atol=atol * 100)
    torch.testing.assert_close(_d(sub[0]), _d(full[0])[:n], rtol=rtol, atol=atol)
    torch.testing.assert_close(_stack(sub[2]['log_m_k']), _stack(full[2]['log_m_k'])[:, :n], rtol=rtol, atol=atol * 10)
    for key in ('kl_l_k', 'kl_m_k'):
        if key in full[1] and len(full[1][key]):
            torch.testing.assert_close(_stack(sub[1][key]), _stack(full[1][key])[:, :n], rtol=rtol, atol=atol * 100)
