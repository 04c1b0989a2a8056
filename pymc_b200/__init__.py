"""pymc_b200 -- B200-native NUTS engine behind PyMC's sampler seams."""
from . import models  # noqa: F401
from ._lib import B200Error  # noqa: F401
from .models import ModelSpec  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require the built library
    if name in ("CompiledModel", "NutsResult"):
        from . import engine

        return getattr(engine, name)
    if name in ("sample_b200_nuts", "SampleResult", "SamplingError"):
        from . import sampling

        return getattr(sampling, name)
    if name in ("B200LogpDlogp", "B200NUTS", "from_pymc"):
        from . import step

        return getattr(step, name)
    raise AttributeError(name)
