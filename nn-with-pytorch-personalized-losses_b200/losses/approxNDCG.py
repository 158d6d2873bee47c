from ltr_b200.losses.approxNDCG import *  # noqa: F401,F403
