"""Drop-in mirror of the reference's `losses` package (losses/__init__.py:1-5 exposes exactly these five
modules; listnet and ordinal are imported by path, as the reference drivers do)."""
from . import approxNDCG  # noqa: F401
from . import exactNDCG  # noqa: F401
from . import lambdaL  # noqa: F401
from . import orderScore  # noqa: F401
from . import riskLosses  # noqa: F401
