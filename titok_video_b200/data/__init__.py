"""The step before the path: token-budget batching of decoded clips (what feeds TiTok.forward in train.py)."""
from .batching import canonical_order, dynamic_batches  # noqa: F401
