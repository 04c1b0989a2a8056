"""Seams B1/B3: objects the reference's own step machinery accepts.

``B200LogpDlogp`` satisfies what ``GradientSharedStep`` / ``CpuLeapfrogIntegrator`` read from a
``ValueGradFunction`` (pymc/step_methods/arraystep.py:175-205, hmc/integration.py:42-66), so the UNMODIFIED
reference NUTS can run on the CUDA logp/grad one point at a time.  That proves correctness behind the
reference's own sampler; speed comes from the whole-run seam (``sample_b200_nuts``).
"""
from __future__ import annotations

import numpy as np

from .engine import CompiledModel


class B200LogpDlogp:
    def __init__(self, compiled: CompiledModel):
        self._cm = compiled
        self._raveled_inputs = True
        self.dtype = "float64"
        self._extra_vars_shared = {}
        self.trust_input = True

    def _pytensor_function(self, q):
        lp, g = self._cm.logp_dlogp(np.asarray(q, dtype=np.float64))
        return lp, g

    def set_extra_values(self, point):  # no non-gradient (shared) variables in the supported models
        pass

    def __call__(self, q):
        return self._pytensor_function(q)


def from_pymc(model):
    """Lower a ``pm.Model`` to a ``ModelSpec``.  Not implemented this round: PyMC/PyTensor are
    not importable in this image, so the lowering cannot be exercised; use ``pymc_b200.models`` specs."""
    raise NotImplementedError(
        "from_pymc: graph lowering is not implemented yet; build a pymc_b200.models.ModelSpec "
        "(eight_schools(), radon(), ...) and pass it as model="
    )
