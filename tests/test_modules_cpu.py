"""Host-side behaviour of the drop-in modules: constructor compatibility with torchaudio, table
bits, eager errors, no CPU fallback.  CPU only (no kernel runs)."""
import inspect

import numpy as np
import pytest
import torch
import torchaudio.transforms as T

from helpers import LFCC_CFG, MEL_CFG


def _positional(sig):
    return [(n, p.default) for n, p in sig.parameters.items()
            if p.kind == inspect.Parameter.POSITIONAL_OR_KEYWORD and n != "self"]


@pytest.mark.parametrize("name", ["LFCC", "MelSpectrogram", "ComputeDeltas", "Spectrogram"])
def test_constructor_signature_matches_torchaudio(fe, name):
    ours = _positional(inspect.signature(getattr(fe, name).__init__))
    theirs = _positional(inspect.signature(getattr(T, name).__init__))
    assert [n for n, _ in ours][: len(theirs)] == [n for n, _ in theirs]
    for (n, d1), (_, d2) in zip(ours, theirs):
        if n == "window_fn":
            assert d1 is d2
        else:
            assert d1 == d2, f"default of {name}.{n}: {d1!r} != {d2!r}"


def test_tables_are_bitwise_torchaudio(fe):
    m = fe.LFCC(**LFCC_CFG)
    ref = T.LFCC(**LFCC_CFG)
    assert torch.equal(m.filter_mat, ref.filter_mat)
    assert torch.equal(m.dct_mat, ref.dct_mat)
    mm = fe.MelSpectrogram(**MEL_CFG)
    rm = T.MelSpectrogram(**MEL_CFG)
    assert torch.equal(mm.fb, rm.mel_scale.fb)
    ms = fe.MelSpectrogram(16000, n_fft=512, n_mels=40, norm="slaney", mel_scale="slaney")
    rs = T.MelSpectrogram(16000, n_fft=512, n_mels=40, norm="slaney", mel_scale="slaney")
    assert torch.equal(ms.fb, rs.mel_scale.fb)
    m128 = fe.LFCC(16000, speckwargs=dict(n_fft=512))
    r128 = T.LFCC(16000, speckwargs=dict(n_fft=512))
    assert torch.equal(m128.filter_mat, r128.filter_mat) and torch.equal(m128.dct_mat, r128.dct_mat)


def test_no_parameters_and_no_persistent_buffers(fe):
    """Checkpoints of the maze models load unchanged around the module."""
    m = fe.LFCCDelta(**LFCC_CFG)
    assert list(m.parameters()) == []
    assert len(m.state_dict()) == 0


def test_cpu_tensor_raises_no_fallback(fe):
    m = fe.LFCCDelta(**LFCC_CFG)
    with pytest.raises(ValueError, match="no CPU path"):
        m(torch.zeros(2, 64600))
    with pytest.raises(TypeError):
        m(torch.zeros(2, 64600, dtype=torch.float64))
    with pytest.raises(ValueError):
        fe.ComputeDeltas()(torch.zeros(2, 3, 50))


@pytest.mark.parametrize("kw,exc", [
    (dict(dct_type=3), ValueError),
    (dict(norm="bogus", speckwargs=dict(n_fft=512)), ValueError),
    (dict(speckwargs=dict(n_fft=400)), NotImplementedError),
    (dict(speckwargs=dict(n_fft=512, center=False)), NotImplementedError),
    (dict(speckwargs=dict(n_fft=512, pad_mode="constant")), NotImplementedError),
    (dict(speckwargs=dict(n_fft=512, power=1.0)), NotImplementedError),
    (dict(speckwargs=dict(n_fft=512), n_lfcc=600), ValueError),
    (dict(speckwargs=dict(n_fft=512), variant="triton"), ValueError),
    (dict(speckwargs=dict(n_fft=512), deltas=3), ValueError),
])
def test_unsupported_arguments_raise_eagerly(fe, kw, exc):
    with pytest.raises(exc):
        fe.LFCC(16000, **kw)


def test_compute_deltas_arguments(fe):
    with pytest.raises(ValueError):
        fe.ComputeDeltas(win_length=2)
    with pytest.raises(NotImplementedError):
        fe.ComputeDeltas(mode="reflect")


def test_output_geometry(fe):
    m = fe.LFCCDelta(**LFCC_CFG)
    assert m.engine.n_out == 60 and m.engine.n_frames(64600) == 404
    assert fe.B200LFCC is fe.LFCC and fe.B200LFCCDelta is fe.LFCCDelta


def test_torchaudio_group_rule(fe):
    from importlib import import_module
    tr = import_module("audio-deepfake-detection-fmsl_b200.transforms")
    assert tr._torchaudio_group((64600,)) == 1
    assert tr._torchaudio_group((8, 64600)) == 8
    assert tr._torchaudio_group((8, 1, 64600)) == 1
    assert tr._torchaudio_group((4, 2, 64600)) == 2


def test_missing_library_fails_loudly(fe, monkeypatch):
    monkeypatch.setattr(fe._lib, "_lib", None)
    monkeypatch.setattr(fe._lib, "LIB_PATH", "/nonexistent/libb200fe.so")
    with pytest.raises(OSError, match="no CPU fallback"):
        fe._lib.load()


def test_pack_clips_aligns_clip_starts(fe):
    """pack_clips: the ragged input with every clip on a 16-byte boundary (what lets the streaming kernel read the
    clips that pad() only truncates in place), clips recoverable bit for bit, empty clips refused like pad()."""
    rs = np.random.RandomState(0)
    clips = [rs.standard_normal(n).astype(np.float32) for n in (5, 64600, 1, 70001, 16000)]
    flat, offsets, lengths = fe.pack_clips(clips)
    assert flat.dtype == torch.float32 and offsets.dtype == torch.int64 and lengths.dtype == torch.int32
    assert lengths.tolist() == [5, 64600, 1, 70001, 16000]
    assert all(o % 4 == 0 for o in offsets.tolist())
    assert offsets.tolist() == [0, 8, 64608, 64612, 134616] and flat.numel() == 150616
    for c, o, l in zip(clips, offsets.tolist(), lengths.tolist()):
        assert np.array_equal(flat[o:o + l].numpy(), c)
    f1, o1, _ = fe.pack_clips([torch.from_numpy(c) for c in clips], align=1)     # back to back
    assert o1.tolist() == [0, 5, 64605, 64606, 134607] and f1.numel() == 150607
    with pytest.raises(ValueError):
        fe.pack_clips([clips[0], np.zeros(0, np.float32)])
    with pytest.raises(ValueError):
        fe.pack_clips([])
