from ltr_b200.utils.metrics import *  # noqa: F401,F403
from ltr_b200.utils.metrics import dcg, getGeoRiskDefault, mNdcg, ndcg, ndcg_sweep, torchNdcg  # noqa: F401
