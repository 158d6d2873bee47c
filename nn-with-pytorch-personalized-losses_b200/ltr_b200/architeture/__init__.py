"""Rankers of the reference's `architeture` package that have a fused B200 path."""
from .tripleLayer import TripleLayerNet, fused_triple_approx_ndcg  # noqa: F401
