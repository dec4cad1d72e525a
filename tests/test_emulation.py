"""The kernels' phase functions (csrc/fe_fft.cuh, csrc/fe_tail.cuh), executed on the CPU with lanes
and threads as loops (tests/emu/fe_emu.cpp), against the golden vectors and the oracle.  This is the
no-GPU check of the CUDA path's index logic: Stockham addressing, real-FFT split, band filterbank,
top_db, DCT, replicate-clamped delta tiles, reflect padding, repeat-pad, pre-emphasis."""
import numpy as np
import pytest

from helpers import (FULL_TONAL, LFCC_CFG, MEL_CFG, SHORT_TONAL, TOL, TOL_TONAL_MEL, assert_feat_close,
                     assert_rows_close, emulate, golden, golden_full_rows, golden_short_rows)
from oracle import frontend_oracle as O
from oracle import synth


def test_emulated_lfcc_full_vs_golden(fe):
    m = fe.LFCCDelta(**LFCC_CFG)
    out, pw = emulate(m, golden_full_rows())
    assert_rows_close(out, golden()["lfcc_dd_full"], FULL_TONAL, "emulated kernels vs golden")


@pytest.mark.parametrize("ft,tt", [(32, 104), (16, 8), (4, 24), (32, 128)])
def test_emulated_tiling_is_invariant(fe, ft, tt):
    """Frame tile / tail tile sizes only change the decomposition, never the result."""
    m = fe.LFCCDelta(**LFCC_CFG)
    x = golden_short_rows()
    base, _ = emulate(m, x, ft=32, tt=104)
    out, _ = emulate(m, x, ft=ft, tt=tt)
    assert np.array_equal(out, base)
    assert_rows_close(out, golden()["lfcc_dd_short"], SHORT_TONAL, "short")


def test_emulated_variants_vs_golden(fe):
    g = golden()
    x = golden_short_rows()
    out, _ = emulate(fe.LFCC(**LFCC_CFG, log_lf=True), x)
    assert_rows_close(out, g["lfcc_loglf_short"], SHORT_TONAL, "log_lf")
    out, _ = emulate(fe.LFCC(16000, n_filter=128, n_lfcc=40, deltas=1, speckwargs=LFCC_CFG["speckwargs"]), x)
    assert_rows_close(out, g["lfcc_default128_short"], SHORT_TONAL, "n_filter=128")
    out, _ = emulate(fe.LFCCDelta(**LFCC_CFG, preemphasis=0.97), x)
    assert_rows_close(out, g["lfcc_preemph_short"], SHORT_TONAL, "preemphasis")


def test_emulated_mel_vs_golden(fe):
    g = golden()
    out, _ = emulate(fe.MelSpectrogram(**MEL_CFG, log="db"), golden_full_rows()[:2], ft=16, tt=128)
    assert_rows_close(out, g["mel_db_full"], (1,), "mel db", TOL_TONAL_MEL)
    x = golden_short_rows()
    out, _ = emulate(fe.MelSpectrogram(**MEL_CFG), x, ft=8, tt=16)
    ref = g["mel_power_short"]
    assert np.abs(out - ref).max() <= 2e-5 * ref.max()


def test_emulated_ragged_repeat_pad(fe):
    """Ragged clips padded on the fly equal pad() (maze5.py:280-285) followed by the dense path."""
    T = 6000
    rs = np.random.RandomState(5)
    lengths = np.array([700, 6000, 9000, 2999, 6001, 1], dtype=np.int32)
    lengths[-1] = 1500
    offsets = np.zeros(len(lengths), np.int64)
    offsets[1:] = np.cumsum(lengths[:-1])
    flat = rs.randn(int(lengths.sum())).astype(np.float32) * 0.1
    dense = np.stack([O.pad_repeat(flat[o:o + l], T) for o, l in zip(offsets, lengths)])
    m = fe.LFCCDelta(**LFCC_CFG)
    a, _ = emulate(m, flat, offsets=offsets, lengths=lengths, T=T)
    b, _ = emulate(m, dense)
    assert np.array_equal(a, b)
    assert_feat_close(a, O.lfcc(dense, deltas=2), TOL, "ragged vs oracle")


def test_emulated_top_db_group(fe):
    s3 = synth.s3_edge(8000)
    x = s3[4:6]
    m = fe.LFCC(**LFCC_CFG, top_db_scope="torchaudio")
    a, _ = emulate(m, x, group=2)
    assert_feat_close(a, O.lfcc(x, top_db_group=2), 2e-4, "coupled")
    b, _ = emulate(m, x, group=1)
    assert_feat_close(b, O.lfcc(x, top_db_group=1), 2e-4, "per utterance")


@pytest.mark.parametrize("n_fft,win,hop", [(64, 64, 16), (128, 100, 37), (256, 200, 80), (2048, 1200, 512), (4096, 4096, 1024)])
def test_emulated_power_spectrum_sizes(fe, n_fft, win, hop):
    """Every supported FFT size (radix-4 chains and the radix-2 first stage) against numpy."""
    T = 3 * n_fft + 77
    x = synth.s1_noise(2, T, seed=n_fft)
    m = fe.MelSpectrogram(16000, n_fft=n_fft, win_length=win, hop_length=hop, n_mels=8)
    _, pw = emulate(m, x, ft=4, tt=8)
    ref = O.power_spectrogram(x.astype(np.float64), n_fft, win, hop, window=O.hann_window(win, np.float64))
    assert pw.shape == ref.shape
    assert np.abs(pw - ref).max() <= 2e-6 * ref.max()
