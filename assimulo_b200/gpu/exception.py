"""Exception names of the reference (src/exception.py), so that user code that catches them keeps working with the
batched solvers.  They are plain subclasses of one root; the table below is all there is to them."""


class AssimuloException(Exception):
    """Root of every exception the batched solvers raise for option validation and call-order errors."""


def _named(name, doc):
    cls = type(name, (AssimuloException,), {"__doc__": doc, "__module__": __name__})
    globals()[name] = cls
    return cls


for _name, _doc in (
        ("ODE_Exception", "general solver error"),
        ("Explicit_ODE_Exception", "problem / option errors of the explicit-ODE solver base class"),
        ("Dopri5_Exception", "Dopri5 option validation"),
        ("Radau_Exception", "Radau5ODE option validation"),
        ("Rodas_Exception", "RodasODE option validation"),
        ("TerminateSimulation", "what handle_event raises in the reference to stop a run; per-member status 1 here"),
        ("TimeLimitExceeded", "the reference's time_limit abort (src/explicit_ode.pyx:290-292): raised by simulate() when members "
         "were still waiting for a warp at the deadline; their status is -9, everything else is kept"),
):
    _named(_name, _doc)
del _name, _doc


class Radau5Error(AssimuloException):
    """Carries the solver's negative return flag, like the reference's Radau5Error (src/lib/radau_core.py)."""

    def __init__(self, value, t=0.0, err_msg=""):
        self.value, self.t, self.err_msg = value, t, err_msg
        AssimuloException.__init__(self, "Radau5 failed with flag %s at time %s. Message: %s" % (value, t, err_msg))
