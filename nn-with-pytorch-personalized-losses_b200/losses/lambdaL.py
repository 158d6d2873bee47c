from ltr_b200.losses.lambdaL import *  # noqa: F401,F403
