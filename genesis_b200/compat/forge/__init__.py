"""Forge stand-in (reference: forge/forge/__init__.py:26-34): `forge.config()`, `forge.flags`,
`forge.experiment_tools`.  Put `genesis_b200/compat` on sys.path to use it."""
from . import flags  # noqa: F401
from . import experiment_tools  # noqa: F401
from .experiment_tools import load  # noqa: F401


def config():
    return experiment_tools.parse_flags()
