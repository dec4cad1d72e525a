"""The DFT-GEMM variant without a GPU: host-packed operand tiles / filterbank tables plus the kernel's
per-thread arithmetic (csrc/fe_gemm.cuh), with tcgen05.mma replaced by loops over the fp16 operand
images at their UMMA layout offsets (tests/emu/fe_emu.cpp).  Checks the folded-DFT math, the parity
split, the split-fp16 scaling, the chunk-local filterbank tables, bin n_fft/4 and the reflect edges."""
import numpy as np
import pytest

from helpers import LFCC_CFG, MEL_CFG, emulate_gemm_energies
from oracle import frontend_oracle as O
from oracle import synth


def _ref_energies(x, n_fft, win, hop, n_filter, sr=16000):
    x64 = x.astype(np.float64)
    spec = O.power_spectrogram(x64, n_fft, win, hop, window=O.hann_window(win, np.float64))
    return O.apply_fbank(spec, O.linear_fbanks(n_fft // 2 + 1, 0.0, sr / 2, n_filter, sr).astype(np.float64))


def test_gemm_tables_available_for_lfcc_config(fe):
    m = fe.LFCCDelta(**LFCC_CFG, variant="dft_gemm")
    assert m.engine.resolved_variant() == "dft_gemm"
    auto = fe.LFCCDelta(**LFCC_CFG)
    assert auto.engine.resolved_variant() == "dft_gemm"  # measured faster: AUTO picks it
    assert fe.LFCCDelta(**LFCC_CFG, variant="fft").engine.resolved_variant() == "fft"


@pytest.mark.parametrize("kw", [
    dict(speckwargs=dict(n_fft=512, win_length=400, hop_length=160)),          # win != 2*hop
    dict(speckwargs=dict(n_fft=1024, win_length=1024, hop_length=512)),        # n_fft/4 > 128 columns
    dict(speckwargs=dict(n_fft=512, win_length=320, hop_length=160, window_fn=__import__("torch").hamming_window)),
    dict(speckwargs=dict(n_fft=512, win_length=320, hop_length=160), n_filter=128),
])
def test_gemm_unsupported_configs_fall_to_fft_or_raise(fe, kw):
    base = dict(sample_rate=16000, n_filter=20, n_lfcc=20)
    base.update(kw)
    assert fe.LFCC(**base).engine.resolved_variant() == "fft"         # AUTO: the FFT variant
    with pytest.raises(NotImplementedError):
        fe.LFCC(**base, variant="dft_gemm")                           # explicit request: no silent fallback
    with pytest.raises(NotImplementedError):
        fe.MelSpectrogram(**MEL_CFG, variant="dft_gemm")


def test_preemphasis_keeps_the_tensor_core_variant(fe):
    # pre-emphasised (and ragged) input reaches the streaming kernel as dense rows (fe_dense_rows_kernel)
    m = fe.LFCC(16000, n_filter=20, n_lfcc=20, speckwargs=dict(n_fft=512, win_length=320, hop_length=160), preemphasis=0.97)
    assert m.engine.resolved_variant() == "dft_gemm"


def test_emulated_gemm_energies_lfcc(fe):
    m = fe.LFCCDelta(**LFCC_CFG, variant="dft_gemm")
    x = np.concatenate([synth.s1_noise(1), synth.s2_speechlike(1), synth.s3_edge()[[1, 2, 3, 5]]], 0)
    e = emulate_gemm_energies(m, x)
    ref = _ref_energies(x, 512, 320, 160, 20)
    assert e.shape == ref.shape == (6, 20, 404)
    for r in range(x.shape[0]):
        assert np.abs(e[r] - ref[r]).max() <= 3e-6 * ref[r].max(), r
    # all-zero utterance: exactly zero energies, no NaN from the frame scale
    z = emulate_gemm_energies(m, np.zeros((1, 64600), np.float32))
    assert not z.any()


def test_emulated_gemm_scale_invariance(fe):
    """The per-frame power-of-two scale makes the result (nearly) exactly homogeneous: scaling the
    input by 2^k scales the energies by 4^k — loud (int16-range) and very quiet inputs included."""
    m = fe.LFCCDelta(**LFCC_CFG, variant="dft_gemm")
    x = synth.s1_noise(1, 8000)
    base = emulate_gemm_energies(m, x)
    for k in (-30, 15):
        e = emulate_gemm_energies(m, np.ldexp(x, k).astype(np.float32))
        assert np.array_equal(e, np.ldexp(base, 2 * k).astype(np.float32))


@pytest.mark.parametrize("T", [64000, 8000, 4000 + 77 * 4])
def test_emulated_gemm_edges_and_short_inputs(fe, T):
    """T % hop == 0 (two trailing edge frames), and utterances shorter than one 128-frame tile."""
    m = fe.LFCCDelta(**LFCC_CFG, variant="dft_gemm")
    x = synth.s1_noise(2, T, seed=T)
    e = emulate_gemm_energies(m, x)
    ref = _ref_energies(x, 512, 320, 160, 20)
    assert e.shape == ref.shape
    assert np.abs(e - ref).max() <= 3e-6 * ref.max()


def test_emulated_gemm_other_geometry(fe):
    m = fe.LFCC(16000, n_filter=10, n_lfcc=10, speckwargs=dict(n_fft=256, win_length=128, hop_length=64), variant="dft_gemm")
    x = synth.s1_noise(2, 6000, seed=3)
    e = emulate_gemm_energies(m, x)
    ref = _ref_energies(x, 256, 128, 64, 10)
    assert np.abs(e - ref).max() <= 3e-6 * ref.max()


def test_drain_tables_qualification_rules(fe):
    """The drain stores a finished filter segment as the filter's final energy, so the packer must refuse banks
    where that does not hold (fe_gemm_tables.cpp: pack_drain_tables): AUTO then takes the FFT variant and an
    explicit dft_gemm request raises."""
    import torch
    base = dict(sample_rate=16000, speckwargs=dict(n_fft=512, win_length=320, hop_length=160))
    # 20 .. 32 linear filters: every class segment is >= 8 columns wide -> qualifies
    for nf in (20, 24, 32):
        assert fe.LFCC(n_filter=nf, n_lfcc=nf, **base).engine.resolved_variant() == "dft_gemm", nf
    # f_max far below Nyquist: the upper filters get no bin at all / the top bins carry no filter
    m = fe.LFCC(n_filter=20, n_lfcc=20, f_max=1000.0, **base)
    assert m.engine.resolved_variant() in ("fft", "dft_gemm")
    x = synth.s1_noise(2, 8000, seed=5)
    if m.engine.resolved_variant() == "dft_gemm":       # whatever the packer decided must be numerically right
        e = emulate_gemm_energies(m, x)
        spec = O.power_spectrogram(x.astype(np.float64), 512, 320, 160, window=O.hann_window(320, np.float64))
        ref = O.apply_fbank(spec, O.linear_fbanks(257, 0.0, 1000.0, 20, 16000).astype(np.float64))
        assert np.abs(e - ref).max() <= 3e-6 * ref.max()
    # a mel bank on the LFCC geometry: narrow low filters (several boundaries per 8-column window) -> refused
    mel = fe.MelSpectrogram(16000, n_fft=512, win_length=320, hop_length=160, n_mels=32)
    assert mel.engine.resolved_variant() == "fft"
    with pytest.raises(NotImplementedError):
        fe.MelSpectrogram(16000, n_fft=512, win_length=320, hop_length=160, n_mels=32, variant="dft_gemm")


@pytest.mark.parametrize("n_mels", [10, 16, 20])
def test_emulated_gemm_energies_mel_bank_on_the_lfcc_geometry(fe, n_mels):
    """A mel bank (segments of very different widths, filter-less bins at the top) through the tensor-core variant's
    tables and drain: n_fft 512 / win 320 / hop 160 with few enough mel filters qualifies (AUTO picks dft_gemm)."""
    m = fe.MelSpectrogram(16000, n_fft=512, win_length=320, hop_length=160, n_mels=n_mels, variant="dft_gemm")
    assert m.engine.resolved_variant() == "dft_gemm"
    x = np.concatenate([synth.s1_noise(1, 8000, seed=n_mels), synth.s2_speechlike(1, 8000, seed=n_mels)], 0)
    e = emulate_gemm_energies(m, x)
    spec = O.power_spectrogram(x.astype(np.float64), 512, 320, 160, window=O.hann_window(320, np.float64))
    ref = O.apply_fbank(spec, O.melscale_fbanks(257, 0.0, 8000.0, n_mels, 16000).astype(np.float64))
    assert e.shape == ref.shape
    for r in range(2):
        assert np.abs(e[r] - ref[r]).max() <= 3e-6 * ref[r].max(), (n_mels, r)
