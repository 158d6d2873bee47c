from ltr_b200.losses.listnet import *  # noqa: F401,F403
