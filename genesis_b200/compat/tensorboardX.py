"""Stand-in for `tensorboardX` (train.py:145): torch's writer has the same API."""
try:
    from torch.utils.tensorboard import SummaryWriter  # noqa: F401
except Exception:  # tensorboard not installed: no-op writer
    class SummaryWriter(object):
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return lambda *a, **k: None
