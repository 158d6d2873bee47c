from ltr_b200.losses.orderScore import *  # noqa: F401,F403
