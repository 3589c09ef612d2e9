"""Either side of the path: token-budget batching of decoded clips (what feeds TiTok.forward in train.py) and a container
for the tokens it produces (what TiTok.decode_indices consumes)."""
from .batching import canonical_order, dynamic_batches  # noqa: F401
from .tokens_io import read_tokens, write_tokens  # noqa: F401
