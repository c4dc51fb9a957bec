"""Drop-in import name: `flash_cosine_sim_attention` re-exports the B200 implementation so the
reference's own scripts (`benchmark.py`, `train.py`, `tests/test.py`) import it unchanged
(reference package surface: __init__.py:1)."""
from flash_cosine_sim_attention_b200 import (  # noqa: F401
    debug,
    flash_cosine_sim_attention,
    l2norm_tensors,
    plain_cosine_sim_attention,
)
from flash_cosine_sim_attention_b200.version import __version__  # noqa: F401
