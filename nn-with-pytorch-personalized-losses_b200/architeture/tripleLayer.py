from ltr_b200.architeture.tripleLayer import *  # noqa: F401,F403
from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg, fused_triple_listnet  # noqa: F401
