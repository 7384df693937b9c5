"""Deterministic synthetic ensembles of BASELINE.json's configs, shared by the golden
generator, the parity tests and bench.py.  No RNG: member i of N is a closed-form function of (i, N)."""
import numpy as np


def vdp_sweep(N, idx=None):
    """cfg 1/2: mu_i = 10^(-1 + 3 i/(N-1)), y0 = [2, -0.6]."""
    i = np.arange(N) if idx is None else np.asarray(idx)
    mu = 10.0 ** (-1.0 + 3.0 * i / max(N - 1, 1))
    y0 = np.tile(np.array([2.0, -0.6]), (len(i), 1))
    return y0, mu.reshape(-1, 1)


def ball_sweep(N, idx=None):
    """cfg 3: y1(0) = 1 + 4 i/(N-1), y2(0) = 0, sw0 = [T, F]; p = [g, e]."""
    i = np.arange(N) if idx is None else np.asarray(idx)
    h0 = 1.0 + 4.0 * i / max(N - 1, 1)
    y0 = np.stack([h0, np.zeros_like(h0)], axis=1)
    p = np.tile(np.array([-9.81, 0.88]), (len(i), 1))
    sw0 = np.tile(np.array([1, 0], dtype=np.uint8), (len(i), 1))
    return y0, p, sw0


def robertson_sweep(N, idx=None):
    """cfg 4: k1 = 0.04 (0.5 + i/(N-1)), k2 = 1e4, k3 = 3e7 (0.5 + (i*7919 mod N)/(N-1))."""
    i = np.arange(N) if idx is None else np.asarray(idx)
    k1 = 0.04 * (0.5 + i / max(N - 1, 1))
    k2 = np.full(len(i), 1e4)
    k3 = 3e7 * (0.5 + ((i * 7919) % N) / max(N - 1, 1))
    y0 = np.tile(np.array([1.0, 0.0, 0.0]), (len(i), 1))
    return y0, np.stack([k1, k2, k3], axis=1)


def gyro_sweep(N, idx=None):
    """cfg 5: pi(0) = [1e4, 5e4, 6e4] (1 + 0.1 (i/(N-1) - 0.5)), Q(0) = I; p = [I1, I2, I3]."""
    i = np.arange(N) if idx is None else np.asarray(idx)
    s = 1.0 + 0.1 * (i / max(N - 1, 1) - 0.5)
    pi0 = np.array([1e4, 5e4, 6e4])[None, :] * s[:, None]
    y0 = np.hstack([pi0, np.tile(np.eye(3).reshape(9), (len(i), 1))])
    p = np.tile(np.array([1000.0, 5000.0, 6000.0]), (len(i), 1))
    return y0, p


ROBERTSON_ATOL = [1e-8, 1e-14, 1e-6]
