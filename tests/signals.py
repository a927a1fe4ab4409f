"""Seeded synthetic test signals: TAU-SED-2019-shaped mono clips at 48 kHz."""
import numpy as np

SR = 48000


def white(n, seed=0, sigma=0.1):
    rng = np.random.default_rng(1234 + seed)
    return np.clip(rng.standard_normal(n) * sigma, -1.0, 1.0)


def pink(n, rng):
    f = np.fft.rfft(rng.standard_normal(n))
    f[1:] /= np.sqrt(np.arange(1, len(f)))
    f[0] = 0
    p = np.fft.irfft(f, n)
    return p / p.std()


def hdr(n, seed=0):
    """0.3*pink(sigma 0.1) + 0.5*sin(2 pi 440 t) + decaying-noise 'door slam' bursts + 1e-4 floor."""
    rng = np.random.default_rng(4321 + seed)
    t = np.arange(n) / SR
    y = 0.03 * pink(n, rng) + 0.5 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * rng.standard_normal(n)
    n_bursts = max(1, n // (10 * SR))
    for s in rng.integers(0, max(1, n - 4800), size=n_bursts):
        m = min(4800, n - s)
        y[s:s + m] += 0.8 * rng.standard_normal(m) * np.exp(-np.arange(m) / 800.0)
    return np.clip(y, -1.0, 1.0)


def silence(n, seed=0):
    rng = np.random.default_rng(999 + seed)
    return 1e-4 * rng.standard_normal(n)


def tone(n, freq=1000.0, amp=0.5):
    return amp * np.sin(2 * np.pi * freq * np.arange(n) / SR)


def impulse(n, pos=0):
    y = np.zeros(n)
    y[pos] = 1.0
    return y


ALL = {"white": white, "hdr": hdr, "silence": silence}
