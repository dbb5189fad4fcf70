"""Empty stand-in: the reference imports tensorflow only for data loading (utils/misc.py:22,
forge/flags.py:25); nothing on the hot path uses it."""
