"""Deterministic synthetic audio shared by tests and bench (numpy, host side).
"Piano-shaped" generator of SURVEY.md section 8d: decaying inharmonic partials
with a 5 ms attack plus a -60 dBFS noise floor, peak-normalised to 0.9."""
import numpy as np


def piano_clip(seed, n_samples, sr=44100, n_notes=24, pitch_range=(21, 108)):
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sr
    y = np.zeros(n_samples)
    dur = n_samples / sr
    for _ in range(n_notes):
        onset = rng.uniform(0, max(dur * 0.9, 1e-3))
        pitch = int(rng.integers(pitch_range[0], pitch_range[1] + 1))
        vel = int(rng.integers(30, 121))
        f0 = 440.0 * 2 ** ((pitch - 69) / 12)
        tau = 0.3 + 1.2 * (108 - pitch) / 87
        tt = t - onset
        env = np.where(tt >= 0, np.exp(-np.maximum(tt, 0) / tau) * np.minimum(np.maximum(tt, 0) / 0.005, 1.0), 0.0)
        for h in range(1, 9):
            f = h * f0 * np.sqrt(1 + 1e-4 * h * h)
            if f < sr / 2:
                y += (vel / 128.0) ** 2 / h * env * np.sin(2 * np.pi * f * tt)
    y += 1e-3 * rng.standard_normal(n_samples)
    return (0.9 * y / np.abs(y).max()).astype(np.float32)
