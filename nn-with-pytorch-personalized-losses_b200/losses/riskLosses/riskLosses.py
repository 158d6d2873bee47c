from ltr_b200.losses.riskLosses.riskLosses import *  # noqa: F401,F403
