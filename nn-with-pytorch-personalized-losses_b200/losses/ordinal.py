from ltr_b200.losses.ordinal import *  # noqa: F401,F403
