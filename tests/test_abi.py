"""The C-ABI library loads on a machine without a GPU, exports every symbol include/b200fe.h
declares, validates parameters, and packs the constant tables correctly (host-only entry points)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from helpers import LFCC_CFG, MEL_CFG, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200fe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(fe):
    lib = fe._lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200fe.h but not exported"
    assert set(syms) == set(fe._lib._SIGNATURES), "ctypes binding and header disagree"
    assert lib.b200fe_version() == 1
    assert lib.b200fe_status_string(-2).decode().startswith("unsupported")


def test_params_struct_layout_matches_header(fe):
    # 14 four-byte fields, no padding
    assert C.sizeof(fe._lib.Params) == 56


def _params(fe, **kw):
    p = fe._lib.Params()
    p.abi_version = 1
    p.n_fft, p.win_length, p.hop_length = 512, 320, 160
    p.n_filter, p.n_coef = 20, 20
    p.log_mode, p.top_db, p.top_db_group = 1, 80.0, 1
    p.deltas, p.delta_win = 2, 5
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_geometry_queries(fe):
    lib = fe._lib.load()
    p = _params(fe)
    assert lib.b200fe_n_frames(C.byref(p), 64600) == 404
    assert lib.b200fe_n_out_channels(C.byref(p)) == 60
    ws = lib.b200fe_workspace_bytes(C.byref(p), 4096, 64600)
    assert ws > 0 and lib.b200fe_workspace_bytes_ex(C.byref(p), 4096, 64600, 0) == ws
    # ragged clips / pre-emphasised input on the tensor-core variant: one chunk of dense rows more
    g = _params(fe, variant=fe._lib.VARIANT_DFT_GEMM)
    base = lib.b200fe_workspace_bytes_ex(C.byref(g), 4096, 64600, 0)
    ragged = lib.b200fe_workspace_bytes_ex(C.byref(g), 4096, 64600, 1)
    assert 0 < ragged - base <= 4096 * 64600 * 4 + 256 and (ragged - base) % (64600 * 4) < 256
    assert lib.b200fe_workspace_bytes_ex(C.byref(_params(fe, variant=fe._lib.VARIANT_DFT_GEMM, preemph=0.97)), 4096, 64600, 0) == ragged
    assert lib.b200fe_workspace_bytes_ex(C.byref(_params(fe, variant=fe._lib.VARIANT_FFT)), 4096, 64600, 1) == \
        lib.b200fe_workspace_bytes_ex(C.byref(_params(fe, variant=fe._lib.VARIANT_FFT)), 4096, 64600, 0)
    q = _params(fe, n_fft=1024, win_length=1024, hop_length=256, n_filter=80, n_coef=0, deltas=0)
    assert lib.b200fe_n_frames(C.byref(q), 64600) == 253
    assert lib.b200fe_n_out_channels(C.byref(q)) == 80


@pytest.mark.parametrize("field,value,status", [
    ("n_fft", 400, -2), ("n_fft", 0, -1), ("win_length", 600, -1), ("hop_length", 0, -1),
    ("deltas", 3, -1), ("delta_win", 4, -2), ("log_mode", 7, -1), ("top_db_group", 0, -1),
    ("abi_version", 99, -1), ("variant", 9, -1),
])
def test_bad_params_are_rejected(fe, field, value, status):
    lib = fe._lib.load()
    p = _params(fe, **{field: value})
    assert lib.b200fe_n_frames(C.byref(p), 64600) == status
    assert len(fe._lib.last_error()) > 0


def test_short_input_rejected_like_torch_stft(fe):
    lib = fe._lib.load()
    assert lib.b200fe_n_frames(C.byref(_params(fe)), 256) == -1  # reflect pad needs T > n_fft/2


def test_tables_blob_contents(fe):
    m = fe.LFCC(**LFCC_CFG)
    blob = m.engine._blob_host
    hdr = np.frombuffer(blob[:96].tobytes(), dtype=np.int32)
    assert np.uint32(hdr[0]) == 0xB200FE03 and hdr[1] == 1
    n_fft, win, hop, n_freq, n_filter, n_coef, total = hdr[2:9]
    assert (n_fft, win, hop, n_freq, n_filter, n_coef) == (512, 320, 160, 257, 20, 20)
    assert total == blob.nbytes
    off_window, off_tw, off_rtw, off_bs, off_bl, off_bo, off_bw, total_w, max_len, off_dct = hdr[9:19]
    w = np.frombuffer(blob[off_window:off_window + 4 * 512].tobytes(), np.float32)
    assert np.array_equal(w[96:416], torch.hann_window(320).numpy()) and not w[:96].any() and not w[416:].any()
    tw = np.frombuffer(blob[off_tw:off_tw + 8 * 256].tobytes(), np.float32).reshape(256, 2)
    k = np.arange(256)
    assert np.abs(tw[:, 0] - np.cos(2 * np.pi * k / 256)).max() < 1e-7
    assert np.abs(tw[:, 1] + np.sin(2 * np.pi * k / 256)).max() < 1e-7
    # the band form reproduces the dense filterbank exactly
    bs = np.frombuffer(blob[off_bs:off_bs + 80].tobytes(), np.int32)
    bl = np.frombuffer(blob[off_bl:off_bl + 80].tobytes(), np.int32)
    bo = np.frombuffer(blob[off_bo:off_bo + 80].tobytes(), np.int32)
    bw = np.frombuffer(blob[off_bw:off_bw + 4 * total_w].tobytes(), np.float32)
    dense = np.zeros((257, 20), np.float32)
    for f in range(20):
        dense[bs[f]:bs[f] + bl[f], f] = bw[bo[f]:bo[f] + bl[f]]
    assert np.array_equal(dense, m.filter_mat.numpy())
    assert max_len == bl.max() and total_w == bl.sum()
    dct = np.frombuffer(blob[off_dct:off_dct + 1600].tobytes(), np.float32).reshape(20, 20)
    assert np.array_equal(dct, m.dct_mat.numpy())


def test_tables_pack_rejects_small_buffer(fe):
    lib = fe._lib.load()
    p = _params(fe)
    w = np.ones(320, np.float32)
    fb = np.zeros((257, 20), np.float32)
    dct = np.zeros((20, 20), np.float32)
    small = np.zeros(128, np.uint8)
    rc = lib.b200fe_tables_pack(C.byref(p), w.ctypes.data, fb.ctypes.data, dct.ctypes.data, small.ctypes.data, small.nbytes)
    assert rc == -3
    rc = lib.b200fe_tables_pack(C.byref(p), None, fb.ctypes.data, dct.ctypes.data, small.ctypes.data, small.nbytes)
    assert rc == -1


def test_mel_tables(fe):
    m = fe.MelSpectrogram(**MEL_CFG, log="db")
    assert m.engine.n_out == 80 and m.engine.n_frames(64600) == 253
    assert m.engine.resolved_variant() in ("fft", "dft_gemm")


def test_eer_entry_point_validates_arguments_without_a_gpu(fe):
    """b200fe_eer_min_dcf / b200fe_eer_workspace_bytes: sizing and argument checks are host code."""
    lib = fe._lib.load()
    assert lib.b200fe_eer_workspace_bytes(71237) >= 2 * 71238 * 8 + 2 * 71238 * 4
    assert lib.b200fe_eer_workspace_bytes(-1) == -1
    assert lib.b200fe_eer_min_dcf(None, None, 10, None, None, 0, None) == -1          # NULL pointers
    assert "b200fe_eer_min_dcf" in fe._lib.last_error()
    buf = (C.c_char * 64)()
    assert lib.b200fe_eer_min_dcf(buf, buf, 0, buf, buf, 64, None) == -1               # n < 1
    assert lib.b200fe_eer_min_dcf(buf, buf, 1000, buf, buf, 64, None) == -3            # workspace too small
    with pytest.raises(TypeError):
        fe.eer_min_dcf_device(torch.ones(4, dtype=torch.int64), torch.zeros(4))        # CPU tensors: no fallback
