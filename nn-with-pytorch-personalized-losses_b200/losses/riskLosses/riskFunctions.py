from ltr_b200.losses.riskLosses.riskFunctions import *  # noqa: F401,F403
