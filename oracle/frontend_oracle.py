"""numpy restatement of the reference feature path (TEST INFRASTRUCTURE, not product code).

The reference repo (Ansh4121/audio-deepfake-detection-fmsl) holds no LFCC / mel / STFT / DCT / delta
code of its own: the arithmetic of the path lives in its un-vendored third-party dependency
**torchaudio** (the reference imports it at ``Thesis/01_Models/01_Baseline_Models/maze5.py:32``; no
version is pinned by the reference, the version restated here is torchaudio 2.11.0).  Each function
below cites the torchaudio ``file:line`` it follows (paths relative to the installed ``torchaudio/``
package) or the reference ``file:line`` where the code does live in the reference (``pad``, EER).

Pinning status: the reference has no tests, golden vectors or fixtures for this path ("parity
unpinned" by the reference itself).  This restatement is pinned instead against
(1) torchaudio 2.11.0 executed live (``tests/test_oracle.py``; torchaudio ships in the image on both
the CPU container and the GPU box) and (2) golden vectors generated from torchaudio and committed
under ``tests/golden/`` together with ``tests/golden/make_golden.py``.

All functions take/return numpy arrays.  ``dtype=np.float32`` mirrors the reference arithmetic;
``dtype=np.float64`` is the "ground truth" used to put the fp32 error of *both* implementations in
perspective.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

UTT_LEN = 64600  # Thesis/01_Models/01_Baseline_Models/maze5.py:280 (pad max_len)


# --------------------------------------------------------------------------------------------
# input contract
# --------------------------------------------------------------------------------------------
def pad_repeat(x: np.ndarray, max_len: int = UTT_LEN) -> np.ndarray:
    """Repeat-tile or truncate a clip to ``max_len`` samples.

    Follows ``pad()`` at Thesis/01_Models/01_Baseline_Models/maze5.py:280-285: clips at least
    ``max_len`` long keep their first ``max_len`` samples, shorter clips are tiled
    ``int(max_len / len) + 1`` times and cut, i.e. sample ``i`` equals ``x[i mod len]``.
    """
    n = x.shape[0]
    if n >= max_len:
        return x[:max_len]
    reps = int(max_len / n) + 1
    return np.tile(x, reps)[:max_len]


def preemphasis(wave: np.ndarray, coeff: float = 0.97) -> np.ndarray:
    """``y[0] = x[0]; y[t] = x[t] - coeff * x[t-1]`` (torchaudio functional/functional.py:2426-2448)."""
    out = wave.copy()
    out[..., 1:] -= wave.dtype.type(coeff) * wave[..., :-1]
    return out


# --------------------------------------------------------------------------------------------
# constant tables
# --------------------------------------------------------------------------------------------
def hann_window(win_length: int, dtype=np.float32) -> np.ndarray:
    """Periodic Hann window, ``torch.hann_window(win_length)`` as used by
    ``Spectrogram.__init__`` (torchaudio transforms/_transforms.py:25 ff., window_fn default)."""
    n = np.arange(win_length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)).astype(dtype)


def centred_window(window: np.ndarray, n_fft: int) -> np.ndarray:
    """Zero-pad a ``win_length`` window to ``n_fft`` the way ``torch.stft`` does (centred, left pad
    ``(n_fft - win_length) // 2``)."""
    win_length = window.shape[0]
    left = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=window.dtype)
    out[left:left + win_length] = window
    return out


def _triangular_filterbank(all_freqs: np.ndarray, f_pts: np.ndarray) -> np.ndarray:
    """torchaudio functional/functional.py:492-516 (``_create_triangular_filterbank``)."""
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up)).astype(np.float32)


def linear_fbanks(n_freqs: int, f_min: float, f_max: float, n_filter: int, sample_rate: int) -> np.ndarray:
    """torchaudio functional/functional.py:590-634 (``linear_fbanks``); float32 like torch.linspace."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=np.float32)
    f_pts = np.linspace(f_min, f_max, n_filter + 2, dtype=np.float32)
    return _triangular_filterbank(all_freqs, f_pts)


def _hz_to_mel_htk(freq: float) -> float:
    """torchaudio functional/functional.py:425-457 (htk branch)."""
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """torchaudio functional/functional.py:518-588 (``melscale_fbanks``, norm=None, mel_scale='htk')."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=np.float32)
    m_pts = np.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2, dtype=np.float32)
    f_pts = (np.float32(700.0) * (np.float32(10.0) ** (m_pts / np.float32(2595.0)) - np.float32(1.0))).astype(np.float32)
    return _triangular_filterbank(all_freqs, f_pts)


def create_dct(n_coef: int, n_filter: int, norm: Optional[str] = "ortho") -> np.ndarray:
    """DCT-II matrix of shape ``(n_filter, n_coef)`` (torchaudio functional/functional.py:636-663)."""
    n = np.arange(n_filter, dtype=np.float32)
    k = np.arange(n_coef, dtype=np.float32)[:, None]
    dct = np.cos(np.float32(math.pi / float(n_filter)) * (n + np.float32(0.5)) * k).astype(np.float32)
    if norm is None:
        dct *= np.float32(2.0)
    else:
        dct[0] *= np.float32(1.0 / math.sqrt(2.0))
        dct *= np.float32(math.sqrt(2.0 / float(n_filter)))
    return np.ascontiguousarray(dct.T)


# --------------------------------------------------------------------------------------------
# stages
# --------------------------------------------------------------------------------------------
def n_frames_of(T: int, hop_length: int) -> int:
    """center=True framing of ``torch.stft``: ``1 + T // hop``."""
    return 1 + T // hop_length


def frames(wave: np.ndarray, n_fft: int, hop_length: int) -> np.ndarray:
    """Reflect-pad by ``n_fft // 2`` and cut ``n_fft``-long frames at stride ``hop_length``
    (``torch.stft(center=True, pad_mode='reflect')`` called from torchaudio
    functional/functional.py:123-134).  ``wave`` is ``(R, T)``; returns ``(R, n_frames, n_fft)``."""
    half = n_fft // 2
    padded = np.pad(wave, ((0, 0), (half, half)), mode="reflect")
    nf = n_frames_of(wave.shape[-1], hop_length)
    idx = np.arange(nf)[:, None] * hop_length + np.arange(n_fft)[None, :]
    return padded[:, idx]


def power_spectrogram(wave: np.ndarray, n_fft: int, win_length: int, hop_length: int,
                      window: Optional[np.ndarray] = None) -> np.ndarray:
    """``|rfft(frame * window)|**2`` -> ``(R, n_fft//2+1, n_frames)``
    (torchaudio functional/functional.py:119-145 with power=2.0, normalized=False)."""
    dt = wave.dtype
    if window is None:
        window = hann_window(win_length, dtype=dt)
    w = centred_window(window.astype(dt), n_fft)
    fr = frames(wave, n_fft, hop_length) * w
    spec = np.fft.rfft(fr, axis=-1)
    if dt == np.float32:
        spec = spec.astype(np.complex64)
    p = (spec.real * spec.real + spec.imag * spec.imag).astype(dt)
    return np.ascontiguousarray(np.swapaxes(p, -1, -2))


def apply_fbank(spec: np.ndarray, fbank: np.ndarray) -> np.ndarray:
    """``(spec^T @ fb)^T`` -> ``(R, n_filter, n_frames)`` (torchaudio transforms/_transforms.py:818)."""
    fb = fbank.astype(spec.dtype)
    return np.einsum("rkt,kf->rft", spec, fb, optimize=True).astype(spec.dtype)


def amplitude_to_db(x: np.ndarray, top_db: Optional[float] = 80.0, group: int = 1) -> np.ndarray:
    """``AmplitudeToDB('power', top_db)``: ``10*log10(clamp(x, 1e-10))`` then
    ``max(x_db, amax - top_db)`` (torchaudio functional/functional.py:356-405).  ``x`` is
    ``(R, F, T)``; the clamp maximum is taken over ``group`` consecutive rows: 1 reproduces a
    ``(B,1,T)`` waveform input (per-utterance, what the maze models feed the slot,
    maze5.py:235-241), ``R`` reproduces torchaudio's packing of a 2-D ``(B,T)`` input."""
    dt = x.dtype
    x_db = (dt.type(10.0) * np.log10(np.maximum(x, dt.type(1e-10)))).astype(dt)
    if top_db is not None:
        R = x.shape[0]
        g = x_db.reshape(R // group, -1)
        floor = (g.max(axis=1) - dt.type(top_db)).astype(dt)
        x_db = np.maximum(x_db, np.repeat(floor, group)[:, None, None])
    return x_db


def log_offset(x: np.ndarray, offset: float = 1e-6) -> np.ndarray:
    """``log(x + 1e-6)`` — the ``log_lf=True`` branch (torchaudio transforms/_transforms.py:820-822)."""
    return np.log(x + x.dtype.type(offset)).astype(x.dtype)


def apply_dct(x: np.ndarray, dct: np.ndarray) -> np.ndarray:
    """``(x^T @ dct)^T`` -> ``(R, n_coef, n_frames)`` (torchaudio transforms/_transforms.py:827)."""
    return np.einsum("rft,fc->rct", x, dct.astype(x.dtype), optimize=True).astype(x.dtype)


def compute_deltas(c: np.ndarray, win_length: int = 5) -> np.ndarray:
    """``d_t = sum_{n=-N..N} n * c_{t+n} / denom`` with replicate-padded edges
    (torchaudio functional/functional.py:961-1008, mode='replicate')."""
    if win_length < 3:
        raise ValueError("win_length must be >= 3")
    n = (win_length - 1) // 2
    denom = n * (n + 1) * (2 * n + 1) / 3
    T = c.shape[-1]
    padded = np.pad(c, [(0, 0)] * (c.ndim - 1) + [(n, n)], mode="edge")
    out = np.zeros_like(c)
    for k in range(-n, n + 1):
        out += c.dtype.type(k) * padded[..., n + k:n + k + T]
    return (out / c.dtype.type(denom)).astype(c.dtype)


def cmvn(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """Per-utterance, per-coefficient mean/variance normalisation over time:
    ``(x - mean_t) / (std_t + eps)`` with the population std.  Absent from the reference and from
    torchaudio's LFCC: parity unpinned, defined here."""
    m = x.mean(axis=-1, keepdims=True, dtype=np.float64)
    v = ((x.astype(np.float64) - m) ** 2).mean(axis=-1, keepdims=True)
    return ((x - m) / (np.sqrt(v) + eps)).astype(x.dtype)


# --------------------------------------------------------------------------------------------
# end-to-end
# --------------------------------------------------------------------------------------------
def lfcc(wave: np.ndarray, sample_rate: int = 16000, n_filter: int = 20, n_lfcc: int = 20,
         n_fft: int = 512, win_length: Optional[int] = 320, hop_length: Optional[int] = 160,
         f_min: float = 0.0, f_max: Optional[float] = None, norm: Optional[str] = "ortho",
         log_lf: bool = False, top_db: Optional[float] = 80.0, top_db_group: int = 1,
         deltas: int = 0, delta_win: int = 5, preemph: Optional[float] = None,
         do_cmvn: bool = False, dtype=np.float32) -> np.ndarray:
    """``LFCC.forward`` (torchaudio transforms/_transforms.py:807-828) followed by ``deltas`` rounds
    of ``ComputeDeltas`` concatenated on the coefficient axis.  ``wave`` is ``(R, T)``; returns
    ``(R, n_lfcc * (1 + deltas), n_frames)`` contiguous."""
    win_length = win_length or n_fft
    hop_length = hop_length or win_length // 2
    f_max = float(sample_rate // 2) if f_max is None else f_max
    x = np.ascontiguousarray(wave, dtype=dtype)
    if preemph is not None:
        x = preemphasis(x, preemph)
    spec = power_spectrogram(x, n_fft, win_length, hop_length)
    fb = linear_fbanks(n_fft // 2 + 1, f_min, f_max, n_filter, sample_rate)
    e = apply_fbank(spec, fb)
    e = log_offset(e) if log_lf else amplitude_to_db(e, top_db, top_db_group)
    c = apply_dct(e, create_dct(n_lfcc, n_filter, norm))
    parts = [c]
    for _ in range(deltas):
        parts.append(compute_deltas(parts[-1], delta_win))
    out = np.concatenate(parts, axis=1)
    if do_cmvn:
        out = cmvn(out)
    return np.ascontiguousarray(out)


def mel_spectrogram(wave: np.ndarray, sample_rate: int = 16000, n_fft: int = 1024,
                    win_length: Optional[int] = None, hop_length: Optional[int] = 256,
                    f_min: float = 0.0, f_max: Optional[float] = None, n_mels: int = 80,
                    log: Optional[str] = None, top_db: Optional[float] = 80.0,
                    top_db_group: int = 1, dtype=np.float32) -> np.ndarray:
    """``MelSpectrogram.forward`` (torchaudio transforms/_transforms.py:515 ff.: Spectrogram power=2
    then MelScale htk/no-norm), optionally followed by ``AmplitudeToDB('power', top_db)``
    (``log='db'``) or ``log(x + 1e-6)`` (``log='log'``).  Returns ``(R, n_mels, n_frames)``."""
    win_length = win_length or n_fft
    hop_length = hop_length or win_length // 2
    f_max = float(sample_rate // 2) if f_max is None else f_max
    x = np.ascontiguousarray(wave, dtype=dtype)
    spec = power_spectrogram(x, n_fft, win_length, hop_length)
    fb = melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)
    e = apply_fbank(spec, fb)
    if log == "db":
        e = amplitude_to_db(e, top_db, top_db_group)
    elif log == "log":
        e = log_offset(e)
    elif log is not None:
        raise ValueError("log must be None, 'db' or 'log'")
    return np.ascontiguousarray(e)


# --------------------------------------------------------------------------------------------
# scoring tail
# --------------------------------------------------------------------------------------------
def roc_curve(y_true: np.ndarray, y_score: np.ndarray):
    """Restatement of ``sklearn.metrics.roc_curve(y_true, y_score)`` with its default
    ``drop_intermediate=True`` (scikit-learn 1.x ``_ranking.py``: stable sort by decreasing score,
    distinct-threshold indices, collinear-point pruning, a leading (0,0) point with threshold inf)
    as called by the reference at Thesis/02_Evaluation_Scripts/Maze5_eval.py:588."""
    y_true = np.asarray(y_true) == 1
    y_score = np.asarray(y_score, dtype=np.float64)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score = y_score[order]
    y_true = y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    thr = y_score[idx]
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    thr = np.r_[np.inf, thr]
    return fps / fps[-1], tps / tps[-1], thr


def eer_min_dcf(y_true: np.ndarray, y_score: np.ndarray):
    """``eer = fpr[nanargmin |fnr - fpr|]``, ``min_dcf = min(fnr + fpr)`` exactly as the reference
    computes them (Thesis/02_Evaluation_Scripts/Maze5_eval.py:588-594;
    score_file_processor.py:176-196)."""
    fpr, tpr, thr = roc_curve(y_true, y_score)
    fnr = 1 - tpr
    i = int(np.nanargmin(np.absolute(fnr - fpr)))
    return float(fpr[i]), float(np.min(fnr + fpr)), float(thr[i])
