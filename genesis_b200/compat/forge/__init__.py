"""Forge stand-in (reference: forge/forge/__init__.py:25-34): `forge.config()`, `forge.flags`, `forge.experiment_tools`,
`forge.load`, `forge.load_from_checkpoint`.  Put `genesis_b200/compat` on sys.path to use it (genesis_b200.enable_compat())."""
from . import flags  # noqa: F401
from . import experiment_tools  # noqa: F401
from .experiment_tools import load, load_from_checkpoint  # noqa: F401


def config():
    """Parse the flags registered so far and return the global flag container (attribute and item access)."""
    experiment_tools.parse_flags()
    return flags.FLAGS
