"""Noise source shared by the plug-ins (kept out of the model-config files so that importing it registers no Forge flags)."""
import torch


class NoiseMixin(object):
    """eps / u source: torch's device RNG in production, a recorded tape for parity runs."""
    _tape = None

    def set_noise_tape(self, tape):
        self._tape = tape

    def _normal(self, shape, like):
        if self._tape is not None:
            return self._tape.normal(shape).to(like.device)
        return torch.randn(shape, device=like.device, dtype=torch.float32)

    def _uniform(self, shape, like):
        if self._tape is not None:
            return self._tape.uniform(shape).to(like.device)
        return torch.rand(shape, device=like.device, dtype=torch.float32)


def seed_rank(seed, rank, device):
    """Re-seed the generator the production noise is drawn from (the default generator of `device`) with seed + rank.
    Data-parallel ranks construct the model under the SAME torch.manual_seed (identical parameters) and then call this, so
    that each rank draws independent eps / IC-SBP seed noise for its shard of the global batch -- the reference's
    nn.DataParallel replicas use their own per-device generators the same way (train.py:153-155)."""
    device = torch.device(device)
    if device.type == 'cuda':
        with torch.cuda.device(device):
            torch.cuda.manual_seed(int(seed) + int(rank))
    else:
        torch.manual_seed(int(seed) + int(rank))
