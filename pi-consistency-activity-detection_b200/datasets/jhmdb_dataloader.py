"""Alias of datasets.load_jhmdb_pytorch_multi (the reference ships this file under that other name's banner)."""
from datasets.load_jhmdb_pytorch_multi import JHMDB  # noqa: F401
