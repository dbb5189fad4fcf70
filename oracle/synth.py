"""ORACLE / test infrastructure -- re-export of the product's synthetic-image generators (genesis_b200/datasets/synth.py) under
the name the tests and the golden-vector generator have always used.  The goldens in tests/golden/ were made with these
functions; they are pure numpy and seeded, so the move did not change them."""
from genesis_b200.datasets.synth import GENERATORS, multid, rooms, stacks  # noqa: F401
