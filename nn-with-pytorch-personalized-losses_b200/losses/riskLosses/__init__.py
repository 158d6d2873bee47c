from . import riskFunctions, riskLosses  # noqa: F401
