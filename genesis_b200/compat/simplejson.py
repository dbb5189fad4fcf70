"""Stand-in for `simplejson` (train.py, forge/experiment_tools.py:34): the stdlib json API."""
from json import *  # noqa: F401,F403
from json import load, loads, dump, dumps  # noqa: F401
