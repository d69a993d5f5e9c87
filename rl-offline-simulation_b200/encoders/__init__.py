"""State encoders: observation -> integer state id, the PSRS queue key (offsim4rl/encoders)."""
from .heuristic import CartpoleBoxEncoder, GridEncoder  # noqa: F401
from .homer import HOMEREncoder  # noqa: F401
