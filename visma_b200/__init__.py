"""visma_b200 — B200 (sm_100a) implementation of VISMA's orientation-constrained ICP path and its
render_depth rasteriser, behind the C ABI in include/visma_b200.h.  No CPU fallback."""
from . import _lib  # noqa: F401
from ._lib import EST_P2P, EST_P2PLANE, EST_P2PLANE_GRAVITY, VismaB200Error  # noqa: F401

__all__ = ["registration", "renderer", "synth", "shard"]
