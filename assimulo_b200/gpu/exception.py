"""Exception names follow the reference's src/exception.py so that user code catching them keeps working."""


class AssimuloException(Exception):
    pass


class Explicit_ODE_Exception(AssimuloException):
    pass


class ODE_Exception(AssimuloException):
    pass


class Dopri5_Exception(AssimuloException):
    pass


class Radau_Exception(AssimuloException):
    pass


class TerminateSimulation(AssimuloException):
    pass


class Radau5Error(AssimuloException):
    """Mirrors src/lib/radau_core.py Radau5Error: carries the solver's negative flag."""

    def __init__(self, value, t=0.0, err_msg=""):
        self.value = value
        self.t = t
        self.err_msg = err_msg
        AssimuloException.__init__(self, "Radau5 failed with flag %s at time %s. Message: %s" % (value, t, err_msg))


class Rodas_Exception(AssimuloException):
    """src/exception.py: raised by RodasODE option validation."""
