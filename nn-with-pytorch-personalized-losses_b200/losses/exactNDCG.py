from ltr_b200.losses.exactNDCG import *  # noqa: F401,F403
