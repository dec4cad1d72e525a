"""Deterministic synthetic inputs (SURVEY.md 8(d) sets S1-S4).  TEST INFRASTRUCTURE ONLY.

numpy ``RandomState`` is used so the same seed gives the same bits on every machine; seed 1234 is
the reference's default seed (Thesis/01_Models/01_Baseline_Models/maze5.py:449).
"""
from __future__ import annotations

import numpy as np

SEED = 1234
SR = 16000
UTT_LEN = 64600


def s1_noise(n: int, T: int = UTT_LEN, seed: int = SEED) -> np.ndarray:
    """S1 (throughput): ``0.1 * N(0,1)`` clipped to [-1, 1]."""
    rs = np.random.RandomState(seed)
    return np.clip(0.1 * rs.standard_normal((n, T)), -1.0, 1.0).astype(np.float32)


def s2_speechlike(n: int, T: int = UTT_LEN, seed: int = SEED) -> np.ndarray:
    """S2 (parity, speech-like dynamic range): 8 random sinusoids (50 Hz - 7.5 kHz, amplitudes
    ``10**U(-3,0)``) + ``1e-3 * N(0,1)``, times a piecewise-linear envelope that contains at least
    0.25 s of exact zeros."""
    rs = np.random.RandomState(seed + 1)
    t = np.arange(T, dtype=np.float64) / SR
    out = np.zeros((n, T), dtype=np.float64)
    for i in range(n):
        f = rs.uniform(50.0, 7500.0, size=8)
        a = 10.0 ** rs.uniform(-3.0, 0.0, size=8)
        ph = rs.uniform(0, 2 * np.pi, size=8)
        sig = (a[:, None] * np.sin(2 * np.pi * f[:, None] * t[None, :] + ph[:, None])).sum(0)
        sig += 1e-3 * rs.standard_normal(T)
        knots = np.sort(rs.choice(np.arange(1, T - 1), size=6, replace=False))
        xs = np.r_[0, knots, T - 1]
        ys = rs.uniform(0.05, 1.0, size=xs.size)
        env = np.interp(np.arange(T), xs, ys)
        z0 = rs.randint(0, max(1, T - SR // 4 - 1))
        env[z0:z0 + min(T, SR // 4 + rs.randint(0, SR // 8))] = 0.0
        out[i] = sig * env
    out /= max(1.0, np.abs(out).max())
    return out.astype(np.float32)


def s3_edge(T: int = UTT_LEN, seed: int = SEED) -> np.ndarray:
    """S3 (edge): all-zero utterance (the reference emits these for unreadable files,
    maze5.py:313-319), impulse at t=0, impulse at t=T-1, full-scale square wave, and one noise
    utterance 100 dB quieter than a loud neighbour (batch independence of ``top_db``)."""
    rs = np.random.RandomState(seed + 2)
    x = np.zeros((6, T), dtype=np.float32)
    x[1, 0] = 1.0
    x[2, T - 1] = 1.0
    x[3] = np.where((np.arange(T) // 40) % 2 == 0, 1.0, -1.0)
    x[4] = np.clip(0.5 * rs.standard_normal(T), -1, 1)
    x[5] = (1e-5 * x[4]).astype(np.float32)
    return x


def s4_ragged(n: int, seed: int = SEED, lo: int = 16000, hi: int = 160000):
    """S4 (ragged, config 5): clip lengths ``U{lo..hi}``; returns ``(flat, offsets, lengths)``."""
    rs = np.random.RandomState(seed + 3)
    lengths = rs.randint(lo, hi + 1, size=n).astype(np.int32)
    offsets = np.zeros(n, dtype=np.int64)
    offsets[1:] = np.cumsum(lengths[:-1])
    flat = np.clip(0.1 * rs.standard_normal(int(lengths.sum())), -1, 1).astype(np.float32)
    return flat, offsets, lengths
