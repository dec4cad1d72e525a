"""The classifier body behind the feature slot (MazeScorer) against golden outputs of the REFERENCE's own
model classes (tests/golden/make_maze_golden.py), and configs 4-5 end to end on the GPU."""
import os

import numpy as np
import pytest
import torch

from helpers import LFCC_CFG, ROOT
from oracle import frontend_oracle as O
from oracle import synth
from oracle.torchaudio_ref import LFCCDeltaRef

MAZE_GOLDEN = os.path.join(ROOT, "tests", "golden", "maze_golden.npz")
SCORE_TOL = 1e-4   # scores (log-softmax) within 1e-4 absolute


def _wave():
    return np.concatenate([synth.s1_noise(3), synth.s2_speechlike(3)], 0)


def _scorer(fe, fmsl):
    m = fe.MazeScorer(fe.LFCC_FILTS, fmsl=fmsl)
    fe.fill_deterministic(m, 1234)
    return m


@pytest.mark.parametrize("key,fmsl", [("maze5", False), ("maze5_fmsl", True)])
def test_scorer_reproduces_reference_classes(fe, key, fmsl):
    g = np.load(MAZE_GOLDEN)
    feats = LFCCDeltaRef()(torch.from_numpy(_wave()))
    out = _scorer(fe, fmsl)(feats).numpy()
    assert out.shape == (6, 2)
    assert np.abs(out - g[key + "_from_features"]).max() <= 2e-5   # fp32 summation order differs with the thread count
    # waveform in, the slot holding the reference CPU feature path
    m = _scorer(fe, fmsl)
    m.frontend = fe.FeatureSlot(LFCCDeltaRef())
    assert np.abs(m(torch.from_numpy(_wave())).numpy() - g[key + "_from_wave"]).max() <= 2e-5   # fp32 summation order differs with the thread count


def test_scorer_is_inference_only_and_checks_checkpoints(fe):
    m = _scorer(fe, False)
    with pytest.raises(RuntimeError):
        m.train()
    sd = {k: v for k, v in m.state_dict().items()}
    sd["sinc_conv.low_hz_"] = torch.zeros(60)          # entries of the replaced slot are dropped
    fe.MazeScorer(fe.LFCC_FILTS).load_reference_state_dict(sd)
    del sd["fc2.bias"]
    with pytest.raises(KeyError):
        fe.MazeScorer(fe.LFCC_FILTS).load_reference_state_dict(sd)
    with pytest.raises(ValueError):
        fe.FeatureSlot(torch.nn.Identity())(torch.zeros(2, 64600))


def test_sweep_labels_and_blocks_are_world_size_independent(fe):
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")
    y = sweep.labels()
    assert y.size == 71237 and y.sum() == 7355 and y[:7355].all() and not y[7355:].any()
    spans = [fe.shard_range(sweep.N_EVAL, r, 8) for r in range(8)]
    assert sum(hi - lo for lo, hi in spans) == 71237


def test_sweep_blocks_are_deterministic_and_class_dependent(fe):
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")
    cpu = torch.device("cpu")
    a = sweep.synthetic_block(3, cpu, n_total=4000, n_bonafide=3100)
    b = sweep.synthetic_block(3, cpu, n_total=4000, n_bonafide=3100)
    assert a.shape == (4000 - 3 * 1024, 64600) and torch.equal(a, b)       # the last block is short
    assert not torch.equal(a[:10], sweep.synthetic_block(2, cpu, n_total=4000, n_bonafide=3100)[:10])
    assert float(a.abs().max()) <= 1.0 * float(torch.exp(torch.tensor(0.35 * 6)))
    # utterances 3072..3099 are bonafide (median gain 0.7), the rest spoofed (median gain 1.0)
    rms = a.pow(2).mean(dim=1).sqrt()
    assert float(rms[:28].median()) < float(rms[28:].median())


# ---- GPU: the CUDA front-end in the slot ------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("key,fmsl", [("maze5", False), ("maze5_fmsl", True)])
def test_cuda_frontend_in_the_slot_matches_reference_scores(fe, key, fmsl):
    g = np.load(MAZE_GOLDEN)
    dev = torch.device("cuda", 0)
    m = _scorer(fe, fmsl).to(dev)
    m.frontend = fe.FeatureSlot(fe.LFCCDelta(**LFCC_CFG))
    x = torch.from_numpy(_wave()).to(dev)
    out = m(x)                                   # (B,T) -> unsqueeze -> slot (B,1,T) -> (B,60,404) -> classifier
    assert out.shape == (6, 2)
    assert np.abs(out.cpu().numpy() - g[key + "_from_wave"]).max() <= SCORE_TOL
    feats = m.frontend(x.unsqueeze(1))
    assert feats.shape == (6, 60, 404) and feats.is_contiguous()


@pytest.mark.gpu
def test_config5_ragged_clips_to_fmsl_scores(fe):
    """Variable-length clips (1-10 s) -> fused repeat-pad + LFCC -> maze5-FMSL classifier; against the same
    classifier fed the reference CPU features of the pad()-ed clips: scores within 1e-4, EER identical."""
    dev = torch.device("cuda", 0)
    n = 48
    flat, offsets, lengths = synth.s4_ragged(n)
    flat = flat * np.repeat(np.where(np.arange(n) < 12, 0.4, 1.0), lengths).astype(np.float32)   # two "classes"
    dense = np.stack([O.pad_repeat(flat[o:o + l], 64600) for o, l in zip(offsets, lengths)])
    cpu = _scorer(fe, True)
    ref = cpu(LFCCDeltaRef()(torch.from_numpy(dense))).numpy()
    gpu = _scorer(fe, True).to(dev)
    front = fe.LFCCDelta(**LFCC_CFG)
    feats = front.forward_ragged(*(torch.from_numpy(a).to(dev) for a in (flat, offsets, lengths)), 64600)
    out = gpu(feats).cpu().numpy()
    assert np.abs(out - ref).max() <= SCORE_TOL
    y = (np.arange(n) < 12).astype(int)
    assert fe.eer_min_dcf(y, out[:, 1])[:2] == fe.eer_min_dcf(y, ref[:, 1])[:2]


@pytest.mark.gpu
def test_config4_sweep_sample_matches_reference_features(fe):
    """A 2,048-utterance slice of the config-4 sweep: EER from CUDA features == EER from the reference CPU
    features through the same classifier, scores within 1e-4; the result does not depend on the batch size."""
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")
    dev = torch.device("cuda", 0)
    n, nb = 2048, 204
    front = fe.LFCCDelta(**LFCC_CFG)
    scorer = _scorer(fe, False).to(dev)
    r = sweep.run_sweep(front, scorer, dev, n_total=n, n_bonafide=nb, batch=512)
    r2 = sweep.run_sweep(front, scorer, dev, n_total=n, n_bonafide=nb, batch=256)
    # cuDNN picks its convolution algorithm per batch size: fp32 summation order changes, nothing else
    assert np.abs(r["scores"] - r2["scores"]).max() <= SCORE_TOL and abs(r["eer"] - r2["eer"]) <= 2.0 / n
    assert r["features_sha256"] == r2["features_sha256"]      # the front-end itself is batch-invariant bit for bit
    # reference CPU features for a 256-utterance sample of the same sweep (both classes present)
    idx = np.r_[0:128, 1024:1152]
    x = torch.cat([sweep.synthetic_block(0, dev, n, nb)[:128], sweep.synthetic_block(1, dev, n, nb)[:128]]).cpu()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    ref = _scorer(fe, False)(LFCCDeltaRef()(x)).numpy()[:, 1]
    assert np.abs(r["scores"][idx] - ref).max() <= SCORE_TOL
    y = sweep.labels(n, nb)[idx]
    assert fe.eer_min_dcf(y, r["scores"][idx])[:2] == fe.eer_min_dcf(y, ref)[:2]
    assert 0.0 <= r["eer"] <= 1.0 and len(r["scores_sha256"]) == 64
