"""assimulo.gpu for B200: batched problems and solvers with the reference's names.

    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    prob = Batched_Explicit_Problem(model="vdp", y0=[2.0, -0.6], p0=mu[:, None])
    sim = Dopri5(prob); sim.rtol = 1e-6
    t, y = sim.simulate(2.0)

Importing this package does not need a GPU; creating a solver's device handle does (no CPU fallback).
"""
from .exception import (AssimuloException, Dopri5_Exception, Explicit_ODE_Exception, Radau5Error,  # noqa: F401
                        Radau_Exception, Rodas_Exception, TerminateSimulation, TimeLimitExceeded)
from .problem import BUILTIN_MODELS, Batched_Explicit_Problem  # noqa: F401
from .solvers import (Batched_Explicit_ODE, Dopri5, ExplicitEuler, ImplicitEuler, Radau5ODE,  # noqa: F401
                      RodasODE, RungeKutta4, RungeKutta34)

__all__ = ["Batched_Explicit_Problem", "Batched_Explicit_ODE", "Dopri5", "RungeKutta34", "RungeKutta4",
           "ExplicitEuler", "ImplicitEuler", "Radau5ODE", "RodasODE", "BUILTIN_MODELS"]
