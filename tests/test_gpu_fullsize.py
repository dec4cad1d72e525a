"""Parity at the BASELINE configurations' own sizes ("weak" items 1-2): the CUDA path against
torchaudio executed live on the host, and the measured distance of both fp32 implementations from a float64
evaluation on the tonal set S2 (written to gpurun_out/ so that the numbers, not only pass/fail, are kept).

  config 2   4096 S1 utterances, LFCC+delta+delta-delta, both kernel families        <= 1e-4
  config 3   8192 S1 utterances, 80-band mel (several workspace chunks), log = db/log/None  <= 1e-4
  config 4   the whole 71,237-utterance sweep: torchaudio features on the host -> the same classifier ->
             EER / min-DCF equal to the CUDA sweep's, scores within 1e-4
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import LFCC_CFG, MEL_CFG, ROOT, TOL, feat_err
from oracle import frontend_oracle as O
from oracle import synth
from oracle.torchaudio_ref import LFCCDeltaRef, LogMelRef

pytestmark = pytest.mark.gpu

# S2 (tonal) rows: the CUDA path may be this much further from the float64 evaluation than torchaudio's own
# fp32 result is (ratio measured on B200: see DESIGN.md section 2 / profiles/r2_parity_s2_table.json), plus an
# absolute 2e-5 for rows where torchaudio happens to sit very close to the truth
S2_SLACK = 1.9


def dev():
    return torch.device("cuda", 0)


def _host_threads():
    torch.set_num_threads(os.cpu_count() or 1)


def _ref_in_chunks(ref, x, chunk=256):
    return torch.cat([ref(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)], 0)


def _max_err_gpu(out, ref_host, chunk=512):
    """max over the batch of |a-b| / max(|b|,1), evaluated on the device chunk by chunk."""
    worst = 0.0
    for i in range(0, out.shape[0], chunk):
        b = ref_host[i:i + chunk].to(out.device)
        e = ((out[i:i + chunk] - b).abs() / b.abs().clamp_min(1.0)).amax()
        worst = max(worst, float(e))
    return worst


def _record(name, payload):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, name), "w") as fh:
            json.dump(payload, fh, indent=1)
    print(name, json.dumps(payload))


@pytest.mark.parametrize("variant", ["dft_gemm", "fft"])
def test_config2_full_batch_against_torchaudio(fe, variant):
    """BASELINE config 2 at its stated size: 4096 S1 utterances, CUDA vs torchaudio live, <= 1e-4."""
    _host_threads()
    x = torch.from_numpy(synth.s1_noise(4096))
    ref = _ref_in_chunks(LFCCDeltaRef(), x)
    m = fe.LFCCDelta(**LFCC_CFG, variant=variant)
    out = m(x.to(dev()))
    assert out.shape == (4096, 60, 404) and bool(torch.isfinite(out).all())
    err = _max_err_gpu(out, ref)
    _record(f"parity_config2_{variant}.json", {"utterances": 4096, "variant": variant, "max_err": err, "tol": TOL})
    assert err <= TOL, (variant, err)


def test_config3_full_batch_against_torchaudio(fe):
    """BASELINE config 3 at its stated size: 8192 S1 utterances through the multi-chunk mel path, all three
    log flavours, CUDA vs torchaudio live."""
    _host_threads()
    x = torch.from_numpy(synth.s1_noise(8192, seed=synth.SEED + 11))
    xd = x.to(dev())
    res = {}
    for log in ("db", "log", None):
        ref = _ref_in_chunks(LogMelRef(log=log), x)
        m = fe.MelSpectrogram(**MEL_CFG, log=log)
        out = m(xd)
        assert out.shape == (8192, 80, 253) and bool(torch.isfinite(out).all())
        if log is None:   # plain mel power: relative to the row maximum (stage tolerance)
            worst = 0.0
            for i in range(0, 8192, 512):
                b = ref[i:i + 512].to(dev())
                worst = max(worst, float(((out[i:i + 512] - b).abs().amax(dim=(1, 2)) / b.amax(dim=(1, 2))).amax()))
            res["power"] = worst
            assert worst <= 2e-5, worst
        else:
            res[log] = _max_err_gpu(out, ref)
            assert res[log] <= TOL, (log, res[log])
        del ref, out
    _record("parity_config3_mel.json", {"utterances": 8192, "max_err": res, "tol": TOL,
                                        "variant": fe.MelSpectrogram(**MEL_CFG, log="db").engine.resolved_variant()})


def test_config4_whole_sweep_eer_equals_reference_features(fe):
    """BASELINE config 4, all 71,237 utterances: features from torchaudio on the host -> the SAME classifier (same
    device, same batch shape) -> EER / min-DCF / threshold, against the CUDA front-end's sweep
    (Maze5_eval.py:588-594: 'downstream EER must be unchanged')."""
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")
    _host_threads()
    d = dev()
    scorer = fe.MazeScorer(fe.LFCC_FILTS, fmsl=False)
    fe.fill_deterministic(scorer, sweep.SEED)
    scorer.to(d)
    front = fe.LFCCDelta(**LFCC_CFG)
    ref_front = LFCCDeltaRef()
    n = sweep.N_EVAL
    ours = np.empty(n, np.float32)
    theirs = np.empty(n, np.float32)
    worst_feat = 0.0
    for block in range((n + sweep.BLOCK - 1) // sweep.BLOCK):
        xb = sweep.synthetic_block(block, d)
        lo = block * sweep.BLOCK
        feats = front(xb)
        ref_feats = _ref_in_chunks(ref_front, xb.cpu()).to(d)
        worst_feat = max(worst_feat, float(((feats - ref_feats).abs() / ref_feats.abs().clamp_min(1.0)).amax()))
        with torch.no_grad():
            ours[lo:lo + xb.shape[0]] = scorer(feats)[:, 1].cpu().numpy()
            theirs[lo:lo + xb.shape[0]] = scorer(ref_feats)[:, 1].cpu().numpy()
    y = sweep.labels()
    a, b = fe.eer_min_dcf(y, ours), fe.eer_min_dcf(y, theirs)
    _record("parity_config4_sweep.json", {
        "utterances": n, "max_feature_err": worst_feat, "max_score_diff": float(np.abs(ours - theirs).max()),
        "cuda": {"eer": a[0], "min_dcf": a[1], "threshold": a[2]},
        "torchaudio": {"eer": b[0], "min_dcf": b[1], "threshold": b[2]}})
    assert worst_feat <= TOL
    assert np.abs(ours - theirs).max() <= 1e-4
    assert a[0] == b[0] and a[1] == b[1], (a, b)       # EER and min-DCF: the same numbers, digit for digit
    assert abs(a[2] - b[2]) <= 1e-4                    # the threshold is a score: within the score tolerance


def test_s2_distance_from_float64_is_measured_and_bounded(fe):
    """Tonal set S2 (+ edge set S3): per-row error of the CUDA path (both LFCC kernel families, and the mel path) and
    of torchaudio's own fp32 result against a float64 evaluation.  The table is recorded; the CUDA path must be
    within S2_SLACK x torchaudio's own error (+2e-5)."""
    _host_threads()
    x = np.concatenate([synth.s2_speechlike(6), synth.s3_edge()], 0)
    xt = torch.from_numpy(x)
    table = {"rows": ["S2"] * 6 + ["S3 zero", "S3 impulse t=0", "S3 impulse t=T-1", "S3 square", "S3 loud", "S3 quiet"]}
    # LFCC + deltas
    g64 = O.lfcc(x.astype(np.float64), deltas=2, dtype=np.float64)
    ref = LFCCDeltaRef()(xt).numpy()
    e_ref = feat_err(ref, g64)
    table["lfcc"] = {"torchaudio_vs_f64": e_ref.tolist()}
    for variant in ("dft_gemm", "fft"):
        out = fe.LFCCDelta(**LFCC_CFG, variant=variant)(xt.to(dev())).cpu().numpy()
        e = feat_err(out, g64)
        table["lfcc"][variant + "_vs_f64"] = e.tolist()
        table["lfcc"][variant + "_vs_torchaudio"] = feat_err(out, ref).tolist()
        table["lfcc"][variant + "_ratio_max"] = float((e / np.maximum(e_ref, 1e-12))[e_ref > 2e-5].max()) if (e_ref > 2e-5).any() else 0.0
    # 80-band mel, dB and log
    table["mel"] = {}
    for log in ("db", "log"):
        g64 = O.mel_spectrogram(x.astype(np.float64), log=log, dtype=np.float64)
        ref = LogMelRef(log=log)(xt).numpy()
        out = fe.MelSpectrogram(**MEL_CFG, log=log)(xt.to(dev())).cpu().numpy()
        e_ref, e = feat_err(ref, g64), feat_err(out, g64)
        table["mel"][log] = {"torchaudio_vs_f64": e_ref.tolist(), "cuda_vs_f64": e.tolist(),
                             "cuda_vs_torchaudio": feat_err(out, ref).tolist(),
                             "ratio_max": float((e / np.maximum(e_ref, 1e-12))[e_ref > 2e-5].max()) if (e_ref > 2e-5).any() else 0.0}
    _record("parity_s2_table.json", table)
    for variant in ("dft_gemm", "fft"):
        e = np.array(table["lfcc"][variant + "_vs_f64"])
        assert (e <= S2_SLACK * np.array(table["lfcc"]["torchaudio_vs_f64"]) + 2e-5).all(), (variant, table["lfcc"])
    for log in ("db", "log"):
        t = table["mel"][log]
        assert (np.array(t["cuda_vs_f64"]) <= S2_SLACK * np.array(t["torchaudio_vs_f64"]) + 2e-5).all(), (log, t)


@pytest.mark.parametrize("n_lfcc,T", [(13, 16000), (19, 16000), (13, 4000)])
def test_fast_tail_odd_coefficient_count_and_odd_tile_width(fe, n_lfcc, T):
    """n_coef odd and an odd number of frames per tile (T=16000, hop 160 -> 101 frames): the fast tail kernel's
    shared-memory carve-up must keep its float4 table aligned."""
    x = synth.s1_noise(3, T, seed=n_lfcc)
    for deltas in (0, 1, 2):
        m = fe.LFCC(16000, n_filter=20, n_lfcc=n_lfcc, speckwargs=LFCC_CFG["speckwargs"], deltas=deltas)
        out = m(torch.from_numpy(x).to(dev())).cpu().numpy()
        ref = LFCCDeltaRef(n_lfcc=n_lfcc, deltas=deltas)(torch.from_numpy(x)).numpy()
        assert out.shape == ref.shape
        assert (feat_err(out, ref) <= TOL).all(), (n_lfcc, T, deltas)


def test_two_devices_in_one_process(fe):
    """One process driving two GPUs: per-device caches (shared-memory opt-in, device check, SM count)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    m = fe.LFCCDelta(**LFCC_CFG)
    x = torch.from_numpy(synth.s1_noise(3))
    a = m(x.to("cuda:0"))
    b = m(x.to("cuda:1"))
    torch.cuda.synchronize("cuda:0")
    torch.cuda.synchronize("cuda:1")
    assert torch.equal(a.cpu(), b.cpu())


def test_ragged_validation(fe):
    m = fe.LFCCDelta(**LFCC_CFG)
    flat = torch.zeros(40000, device=dev())
    off = torch.tensor([0, 20000], dtype=torch.int64, device=dev())
    with pytest.raises(ValueError):
        m.forward_ragged(flat, off, torch.tensor([20000, 0], dtype=torch.int32, device=dev()))
    with pytest.raises(ValueError):
        m.forward_ragged(flat, off, torch.tensor([20000, 20001], dtype=torch.int32, device=dev()))
    with pytest.raises(ValueError):
        m.engine.features(flat, offsets=off, lengths=None, T=64600)
    out = m.forward_ragged(flat, off, torch.tensor([20000, 20000], dtype=torch.int32, device=dev()))
    assert out.shape == (2, 60, 404)


def test_pcm16_host_path_is_bit_identical_to_float_path(fe):
    """int16 PCM over PCIe + x/32768 on the device == the float32 host path on the converted samples, bit for bit."""
    m = fe.LFCCDelta(**LFCC_CFG)
    rs = np.random.RandomState(3)
    pcm = rs.randint(-32768, 32768, size=(300, 64600)).astype(np.int16)
    pcm[0] = 0
    pcm[1, :10] = [-32768, 32767, 1, -1, 0, 2, 3, 4, 5, 6]
    xf = (pcm.astype(np.float32) / np.float32(32768.0))
    out_i = m.forward_host(torch.from_numpy(pcm).pin_memory(), chunk_rows=64, n_streams=3)
    out_f = m.forward_host(torch.from_numpy(xf).pin_memory(), chunk_rows=64, n_streams=3)
    assert out_i.shape == (300, 60, 404)
    assert torch.equal(out_i, out_f)
    assert torch.equal(out_i, m(torch.from_numpy(xf).to(dev())).cpu())


def _sk_metrics(y, s):
    from sklearn.metrics import roc_curve
    fpr, tpr, thr = roc_curve(y, s)
    fnr = 1 - tpr
    i = int(np.nanargmin(np.absolute(fnr - fpr)))
    return float(fpr[i]), float(np.min(fnr + fpr)), float(thr[i])


@pytest.mark.parametrize("n,decimals", [(71237, None), (71237, 2), (7123, 1), (1000, 0), (3, None), (2, None), (40000, 3)])
def test_device_eer_equals_sklearn_digit_for_digit(fe, n, decimals):
    """b200fe_eer_min_dcf against sklearn.metrics.roc_curve + the reference's formulas
    (Maze5_eval.py:588-594): the same float64 numbers, with heavy score ties (rounded scores), negative zeros,
    tiny inputs and the sweep's own size."""
    rs = np.random.RandomState(n + (decimals or 7))
    y = (rs.rand(n) < 0.103).astype(np.int64)
    y[0], y[-1] = 1, 0
    s = (rs.randn(n) + 1.1 * y).astype(np.float32)
    if decimals is not None:
        s = np.round(s, decimals).astype(np.float32)
        s[s == 0] = np.where(rs.rand(int((s == 0).sum())) < 0.5, np.float32(-0.0), np.float32(0.0))
    want = _sk_metrics(y, s)
    got = fe.eer_min_dcf_device(torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev()))
    assert got == want, (n, decimals, got, want)
    assert got == fe.eer_min_dcf(y, s)
    # a perfect separation and a single-class input
    yy = np.r_[np.ones(5), np.zeros(9)].astype(np.int64)
    ss = np.r_[np.full(5, 2.0), np.linspace(-1, 1, 9)].astype(np.float32)
    assert fe.eer_min_dcf_device(torch.from_numpy(yy).to(dev()), torch.from_numpy(ss).to(dev())) == _sk_metrics(yy, ss)
    with pytest.raises(ValueError):
        fe.eer_min_dcf_device(torch.ones(10, dtype=torch.int64, device=dev()), torch.randn(10, device=dev()))
    with pytest.raises(ValueError, match="NaN"):     # scikit-learn raises on NaN scores; the kernel reports status 2
        bad = torch.randn(10, device=dev())
        bad[3] = float("nan")
        fe.eer_min_dcf_device(torch.tensor([1, 0] * 5, device=dev()), bad)
    with pytest.raises(TypeError):
        fe.eer_min_dcf_device(torch.ones(10, dtype=torch.int64), torch.randn(10))


def test_device_eer_in_the_sweep_equals_the_host_restatement(fe):
    """The config-4 sweep takes its EER / min-DCF / threshold from the device kernel; the host restatement of the
    reference's sklearn call on the same gathered scores must give the same numbers."""
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")
    scorer = fe.MazeScorer(fe.LFCC_FILTS, fmsl=False)
    fe.fill_deterministic(scorer, sweep.SEED)
    scorer.to(dev())
    r = sweep.run_sweep(fe.LFCCDelta(**LFCC_CFG), scorer, dev(), n_total=3000, n_bonafide=310, batch=512)
    assert (r["eer"], r["min_dcf"], r["eer_threshold"]) == (r["eer_host"], r["min_dcf_host"], r["eer_threshold_host"])
    assert _sk_metrics(sweep.labels(3000, 310), r["scores"]) == (r["eer"], r["min_dcf"], r["eer_threshold"])


@pytest.mark.parametrize("n_mels", [10, 20])
def test_mel_bank_through_the_tensor_core_variant(fe, n_mels):
    """MelSpectrogram on the LFCC geometry (n_fft 512, win 320, hop 160): mel banks with few filters qualify for the
    tcgen05 variant (non-uniform segments, filter-less top bins); dB features against torchaudio, and bit-equal
    energies... between the two kernel families within the stage tolerance."""
    import torchaudio
    x = torch.from_numpy(np.concatenate([synth.s1_noise(5, seed=n_mels), synth.s3_edge()[[0, 1, 3]]], 0))
    ref_t = torch.nn.Sequential(torchaudio.transforms.MelSpectrogram(16000, n_fft=512, win_length=320, hop_length=160, n_mels=n_mels),
                                torchaudio.transforms.AmplitudeToDB("power", top_db=80.0))
    ref = ref_t(x.unsqueeze(1)).squeeze(1).numpy()
    outs = {}
    for variant in ("dft_gemm", "fft"):
        m = fe.MelSpectrogram(16000, n_fft=512, win_length=320, hop_length=160, n_mels=n_mels, log="db", variant=variant)
        assert m.engine.resolved_variant() == variant
        outs[variant] = m(x.to(dev())).cpu().numpy()
        assert outs[variant].shape == ref.shape
        assert (feat_err(outs[variant], ref) <= TOL).all(), (variant, n_mels, feat_err(outs[variant], ref))


def test_scores_from_host_pcm_equal_the_device_path(fe):
    """sweep.score_host_pcm: 16-bit PCM in host memory -> scores, features never leaving the device; equal to the
    classifier on the front-end's features of the converted samples, chunk for chunk."""
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")
    scorer = fe.MazeScorer(fe.LFCC_FILTS, fmsl=False)
    fe.fill_deterministic(scorer, sweep.SEED)
    scorer.to(dev()).eval()
    front = fe.LFCCDelta(**LFCC_CFG)
    rs = np.random.RandomState(11)
    pcm = torch.from_numpy(rs.randint(-20000, 20000, size=(700, 64600)).astype(np.int16)).pin_memory()
    got = sweep.score_host_pcm(front, scorer, pcm, dev(), chunk_rows=256, n_streams=2)
    assert got.shape == (700,) and got.dtype == torch.float32 and got.device.type == "cpu"
    x = (pcm.to(torch.float32) / 32768.0).to(dev())
    with torch.no_grad():
        want = torch.cat([scorer(front(x[i:i + 256]))[:, 1] for i in range(0, 700, 256)]).cpu()
    assert torch.equal(got, want)
    with pytest.raises(TypeError):
        sweep.score_host_pcm(front, scorer, pcm.to(torch.float32), dev())
