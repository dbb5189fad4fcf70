"""Shared helpers for the GPU parity tests: run the engine and the oracle on identical parameters,
inputs and noise, and compare outputs and per-parameter gradients."""
import torch

from oracle import functional as O
from oracle import models as M


def engine_total_loss(losses, beta=1.0):
    """train.py:227-259 with GECO's beta held fixed."""
    loss = losses['err'].mean(0)
    kl = 0.0
    for key in ('kl_l_k', 'kl_m_k'):
        if key in losses and len(losses[key]):
            kl = kl + torch.stack(list(losses[key]), dim=1).mean(0).sum()
    if 'kl_m' in losses and torch.is_tensor(losses['kl_m']):
        kl = kl + losses['kl_m'].mean(0)
    return loss + beta * kl


def run_engine(model, x, tape):
    model.set_noise_tape(tape)
    model.zero_grad(set_to_none=True)
    recon, losses, stats, att, comp = model(x.cuda())
    engine_total_loss(losses).backward()
    torch.cuda.synchronize()
    model.set_noise_tape(None)
    return recon, losses, stats, att, comp


def run_oracle(name, state_dict, x, tape, cfg, training=True):
    P = {k: (v.detach().cpu().clone().requires_grad_(True) if v.is_floating_point() else v.detach().cpu().clone())
         for k, v in state_dict.items()}
    out = M.FORWARD[name](P, x.cpu(), tape, cfg, training=training)
    M.total_loss(out).backward()
    return out, P


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def compare_grads(model, P, tol, floor_frac=1e-4, skip=()):
    """per-parameter relative L2 error of the gradients; parameters whose true gradient is (numerically)
    zero -- e.g. conv biases feeding BatchNorm -- are compared on an absolute scale."""
    gmax = max(p.grad.norm().item() for p in P.values() if torch.is_tensor(p) and p.grad is not None)
    worst = (0.0, None)
    for name, p in model.named_parameters():
        ref = P[name].grad
        if any(s in name for s in skip):
            continue
        if ref is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, name
            continue
        if p.grad is None:      # parameter unused by the engine (e.g. LSTM weight_hh when K=2): true grad is 0
            assert ref.abs().max().item() == 0, 'no gradient for ' + name
            continue
        g = p.grad.detach().double().cpu()
        denom = max(ref.double().norm().item(), floor_frac * gmax)
        e = (g - ref.double()).norm().item() / denom
        if e > worst[0]:
            worst = (e, name)
        assert e <= tol, 'grad mismatch %s: rel %.3e (|ref| %.3e)' % (name, e, ref.norm().item())
    return worst


def make_tape(seed):
    return O.NoiseTape(seed=seed)


def global_grad_rel_l2(model, P):
    """|| g_engine - g_oracle || / || g_oracle || over the concatenation of all parameter gradients."""
    num = den = 0.0
    for name, p in model.named_parameters():
        ref = P[name].grad
        if ref is None:
            continue
        den += ref.double().pow(2).sum().item()
        g = p.grad.detach().double().cpu() if p.grad is not None else torch.zeros_like(ref, dtype=torch.float64)
        num += (g - ref.double()).pow(2).sum().item()
    return (num / max(den, 1e-300)) ** 0.5


def run_reference(name, cfg, x, tape, seed=0, state_dict=None):
    """The REAL reference nn.Module (live checkout, or the byte-identical copy under oracle/_ref on the GPU box) on the CPU:
    seeded construction (or a given state dict), noise replayed from the tape, total loss as train.py:227-259 with beta = 1,
    backward.  Returns (module, outputs 5-tuple)."""
    from oracle import ref_loader
    ref = ref_loader.load_reference(name, dict(cfg), seed=seed).train()
    if state_dict is not None:
        ref.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()})
    ref.zero_grad(set_to_none=True)
    with ref_loader.replay_noise(tape):
        out = ref(x.cpu())
    engine_total_loss(out[1]).backward()
    return ref, out


def compare_grads_with_module(model, ref, tol, floor_frac=1e-4):
    P = dict(ref.named_parameters())
    return compare_grads(model, P, tol, floor_frac)
