from ltr_b200.utils.dataset import *  # noqa: F401,F403
from ltr_b200.utils.dataset import get_baseline_data, get_data, load_baselines, load_slates, svmDataset  # noqa: F401
