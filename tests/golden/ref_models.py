"""Python twins of the built-in models, written against the REFERENCE's problem API.

Used only to drive the reference itself (oracle/_ref) when generating golden vectors
(tests/golden/make_golden.py), in tests that re-check them, and in bench.py's --impl reference leg.
Each function returns an ``assimulo.problem.Explicit_Problem`` for ONE ensemble member.
Model sources in the reference: examples/radau5ode_vanderpol.py:42-48 (VdP), examples/lsodar_bouncing_ball.py:28-75
(ball), examples/radau5ode_with_jac_sparse.py:46-62 (Robertson), examples/cvode_gyro.py:39-57 (gyro),
tests/conftest.py:26-114 (Extended_Problem), tests/test_solvers.py:27-33 (decay with event).
"""
import numpy as np


def _problem_cls():
    from assimulo.problem import Explicit_Problem
    return Explicit_Problem


def vdp(mu, y0=(2.0, -0.6)):
    def f(t, y):
        return np.array([y[1], mu * ((1. - y[0] * y[0]) * y[1] - y[0])])

    def jac(t, y):
        return np.array([[0., 1.], [mu * (-2. * y[0] * y[1] - 1.), mu * (1. - y[0] * y[0])]])
    p = _problem_cls()(f, np.array(y0, dtype=float), name="vdp")
    p.jac = jac
    return p


def ball(h0, g=-9.81, e=0.88, v0=0.0):
    EP = _problem_cls()

    class Ball(EP):
        def rhs(self, t, y, sw):
            return np.array([y[1], g])

        def state_events(self, t, y, sw):
            return np.array([y[0] if sw[0] else 5., y[1] if sw[1] else 5.])

        def handle_event(self, solver, event_info):
            info = event_info[0]
            self.events.append((solver.t, [int(v) for v in info]))
            if info[0] != 0:
                solver.sw[0] = False
                solver.sw[1] = True
                solver.y[1] = -e * solver.y[1]
            else:
                solver.sw[0] = True
                solver.sw[1] = False
    b = Ball(y0=np.array([h0, v0], dtype=float), sw0=[True, False], name="ball")
    b.events = []
    return b


def robertson(k1=0.04, k2=1e4, k3=3e7, with_jac=True):
    def f(t, y):
        yd0 = -k1 * y[0] + k2 * y[1] * y[2]
        yd2 = k3 * y[1] * y[1]
        return np.array([yd0, -yd0 - yd2, yd2])

    def jac(t, y):
        return np.array([[-k1, k2 * y[2], k2 * y[1]],
                         [k1, -k2 * y[2] - 2. * k3 * y[1], -k2 * y[1]],
                         [0., 2. * k3 * y[1], 0.]])
    p = _problem_cls()(f, np.array([1.0, 0.0, 0.0]), name="robertson")
    if with_jac:
        p.jac = jac
    return p


def gyro(pi0=(1e4, 5e4, 6e4), I=(1000., 5000., 6000.)):
    I1, I2, I3 = I

    def f(t, u):
        om0, om1, om2 = u[0] / I1, u[1] / I2, u[2] / I3
        ud = np.empty(12)
        ud[0] = om2 * u[1] + (-om1) * u[2]
        ud[1] = (-om2) * u[0] + om0 * u[2]
        ud[2] = om1 * u[0] + (-om0) * u[1]
        for c in range(3):
            q0, q1, q2 = u[3 + c], u[6 + c], u[9 + c]
            ud[3 + c] = om2 * q1 + (-om1) * q2
            ud[6 + c] = (-om2) * q0 + om0 * q2
            ud[9 + c] = om1 * q0 + (-om0) * q1
        return ud
    y0 = np.hstack([np.array(pi0, dtype=float), np.eye(3).reshape(9)])
    return _problem_cls()(f, y0, name="gyro")


def gyro_reference_form():
    """The reference's own formulation (np.dot with curl matrices), examples/cvode_gyro.py:39-57."""
    def curl(v):
        return np.array([[0, v[2], -v[1]], [-v[2], 0, v[0]], [v[1], -v[0], 0]])

    def f(t, u):
        I1, I2, I3 = 1000., 5000., 6000.
        pi = u[0:3]
        Q = (u[3:12]).reshape((3, 3))
        omega = np.array([pi[0] / I1, pi[1] / I2, pi[2] / I3])
        pid = np.dot(curl(omega), pi)
        Qd = np.dot(curl(omega), Q)
        return np.hstack([pid, Qd.reshape((9,))])
    y0 = np.hstack([[1000. * 10, 5000. * 10, 6000 * 10], np.eye(3).reshape((9,))])
    return _problem_cls()(f, y0, name="gyro_ref")


def linear(lam=-1.0, c=0.0, y0=1.0):
    def f(t, y):
        return np.array([lam * y[0] + c])
    return _problem_cls()(f, np.array([y0], dtype=float), name="linear")


def linear_ev(lam=-1.0, c=0.0, thr=0.5, y0=1.0):
    EP = _problem_cls()

    class LinEv(EP):
        def rhs(self, t, y, sw):
            return np.array([lam * y[0] + c])

        def state_events(self, t, y, sw):
            return np.array([y[0] - thr])

        def handle_event(self, solver, event_info):
            self.events.append((solver.t, [int(v) for v in event_info[0]]))
            solver.sw[0] = False
    b = LinEv(y0=np.array([y0], dtype=float), sw0=[True], name="linear_ev")
    b.events = []
    return b


def extended():
    EP = _problem_cls()

    class Extended(EP):
        def rhs(self, t, y, sw):
            return np.array([1.0 if sw[0] else -1.0, 0.0, 0.0])

        def state_events(self, t, y, sw):
            return np.array([y[1] - 1.0, -y[2] + 1.0, -t + 1.0])

        def handle_event(self, solver, event_info):
            info = event_info[0]
            self.events.append((solver.t, [int(v) for v in info]))
            while True:
                for i in range(len(info)):
                    if info[i] != 0:
                        solver.sw[i] = not solver.sw[i]
                b = self.state_events(solver.t, solver.y, solver.sw).copy()
                solver.y[1] = (-1.0 if solver.sw[1] else 3.0)
                solver.y[2] = (0.0 if solver.sw[2] else 2.0)
                a = self.state_events(solver.t, solver.y, solver.sw).copy()
                info = [(b[i] < 0.0 and a[i] > 0.0) or (b[i] > 0.0 and a[i] < 0.0) for i in range(len(b))]
                if True not in info:
                    break
    b = Extended(y0=np.array([0.0, -1.0, 0.0]), sw0=[False, True, True], name="extended")
    b.events = []
    return b


def time_ev(tes=(1.0, 2.0, 2.5, 3.0), jump=1.0, slope=1.0, y0=0.0):
    """tests/solvers/test_rungekutta.py:49-82 test_time_event: y' = slope, time events at `tes`, handle_event
    (event_info = [[], True]): y += jump."""
    EP = _problem_cls()

    class TimeEv(EP):
        def rhs(self, t, y):
            return np.array([slope])

        def time_events(self, t, y, sw):
            for ev in tes:
                if t < ev:
                    return ev
            return None

        def handle_event(self, solver, event_info):
            assert event_info[0] == [] and event_info[1]
            self.events.append((solver.t, "time"))
            solver.y = solver.y + jump
    b = TimeEv(y0=np.array([y0], dtype=float), name="time_ev")
    b.events = []
    return b
