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
