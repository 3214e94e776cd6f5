"""Drop-in mirror of the reference's ``utils/`` directory (losses, consistency masks, ramp schedules, the two
metrics the training loop calls), backed by the b200caps kernels.

This is a REGULAR package on purpose.  The reference's ``utils/`` has no ``__init__.py`` (namespace package), and
Python's import system prefers a regular package found ANYWHERE on ``sys.path`` over namespace portions, so with
this directory's parent on ``PYTHONPATH`` the reference scripts' ``from utils.losses import ...`` resolve here even
though the script directory comes first on ``sys.path`` (tests/test_dropin_boundary.py pins that)."""
