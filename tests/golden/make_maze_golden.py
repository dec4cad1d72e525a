"""Generates tests/golden/maze_golden.npz by running the REFERENCE's own model classes
(/root/reference/Thesis/01_Models/.../maze5.py and maze5_fmsl_standardized.py, imported unmodified) on CPU
with their feature slot (``model.sinc_conv``, maze5.py:241) replaced, and records their log-softmax
outputs.  ``tests/test_maze.py`` rebuilds the same weights from the seed (``fill_deterministic``: values
depend only on parameter name, shape and seed) and requires ``MazeScorer`` to reproduce these numbers.

    python tests/golden/make_maze_golden.py          # only in the build container: needs /root/reference

librosa and tensorboardX are imported at the top of the reference files but not used on this path and
not installed here; empty stand-in modules satisfy the imports.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/Thesis"
for sub in ("01_Models/01_Baseline_Models", "01_Models/02_FMSL_Enhanced_Models", "06_Utilities", ""):
    sys.path.insert(0, os.path.join(REF, sub))
sys.modules.setdefault("librosa", types.ModuleType("librosa"))
tbx = types.ModuleType("tensorboardX")
tbx.SummaryWriter = object
sys.modules.setdefault("tensorboardX", tbx)

import b200_frontend as fe  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.torchaudio_ref import LFCCDeltaRef  # noqa: E402

SEED = 1234
CFG = dict(filts=[60, [128, 128], [128, 256]], first_conv=251, sample_rate=16000, nb_fc_node=1024,
           fc_dropout=0.5, nb_classes=2)   # maze5.py:459-474 with the first width set to 3 * n_lfcc


def golden_features():
    """(6, 60, 404) LFCC + delta + delta-delta of seeded S1/S2 utterances from the reference CPU path."""
    x = np.concatenate([synth.s1_noise(3), synth.s2_speechlike(3)], 0)
    return x, LFCCDeltaRef()(torch.from_numpy(x)).numpy()


def main():
    torch.set_num_threads(1)
    import maze5
    import maze5_fmsl_standardized as maze5f
    wave, feats = golden_features()
    out = {}
    for key, cls in (("maze5", maze5.Model5_RawNetSinc_SpecAugment_FocalLoss),
                     ("maze5_fmsl", maze5f.Model5_RawNetSinc_SpecAugment_FocalLoss_FMSL_Standardized)):
        model = cls(dict(CFG), "cpu")
        fe.fill_deterministic(model, SEED)
        model.eval()
        # (a) features handed to the slot's consumer directly: a 3-D input skips the unsqueeze (maze5.py:235)
        model.sinc_conv = torch.nn.Identity()
        with torch.no_grad():
            out[key + "_from_features"] = model(torch.from_numpy(feats)).numpy()
        # (b) waveform -> slot -> classifier, the slot holding the reference CPU feature path
        model.sinc_conv = fe.FeatureSlot(LFCCDeltaRef())
        with torch.no_grad():
            out[key + "_from_wave"] = model(torch.from_numpy(wave)).numpy()
        assert np.allclose(out[key + "_from_wave"], out[key + "_from_features"], atol=1e-5)
        # checkpoint compatibility: the reference's state dict loads into MazeScorer
        ours = fe.MazeScorer(CFG["filts"], CFG["nb_fc_node"], CFG["nb_classes"], fmsl=key.endswith("fmsl"))
        ours.load_reference_state_dict(model.state_dict())
        with torch.no_grad():
            mine = ours(torch.from_numpy(feats)).numpy()
        print(key, "MazeScorer vs reference class: max abs diff", np.abs(mine - out[key + "_from_features"]).max())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "maze_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype, v.ravel()[:4])


if __name__ == "__main__":
    main()
