"""Top-level drop-in: put this directory on sys.path in place of the reference checkout and the
reference drivers' imports (`from losses import *`, `from losses.lambdaL import lambdaLoss`,
`from losses.listnet import listnetLoss`, `from losses.riskLosses.riskFunctions import geoRisk`, ...)
resolve to the B200 implementations."""
from ltr_b200.losses import approxNDCG, exactNDCG, lambdaL, orderScore, riskLosses  # noqa: F401
