from ltr_b200.losses.approxNDCG import approxNDCGLoss  # noqa: F401
