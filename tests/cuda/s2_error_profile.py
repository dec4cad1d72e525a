"""Where the distance from a float64 evaluation sits on the tonal set S2 ("what's weak" 1): relative
error of the filterbank ENERGIES of the two CUDA kernel families and of torchaudio's own fp32 path against a float64
evaluation, binned by the energy's level below its utterance's maximum (the top_db floor is at -80 dB; everything
below it is clamped away by AmplitudeToDB and cannot show in the features).  Writes gpurun_out/r2_s2_error_profile.json.
Run on a GPU box:  python tests/cuda/s2_error_profile.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torchaudio

import b200_frontend as fe
import helpers
from oracle import frontend_oracle as O
from oracle import synth

ROOT = helpers.ROOT
x = synth.s2_speechlike(6)
n_fft, win, hop, nfil = 512, 320, 160, 20
fb64 = O.linear_fbanks(n_fft // 2 + 1, 0.0, 8000.0, nfil, 16000).astype(np.float64)
truth = O.apply_fbank(O.power_spectrogram(x.astype(np.float64), n_fft, win, hop, window=O.hann_window(win, np.float64)), fb64)

# torchaudio's own fp32 energies: Spectrogram (torch.stft) -> matmul with the fp32 bank (transforms/_transforms.py:818)
spec = torchaudio.transforms.Spectrogram(n_fft=n_fft, win_length=win, hop_length=hop)(torch.from_numpy(x))
fb32 = torchaudio.functional.linear_fbanks(n_fft // 2 + 1, 0.0, 8000.0, nfil, 16000)
ta = torch.matmul(spec.transpose(-1, -2), fb32).transpose(-1, -2).numpy().astype(np.float64)

got = {"torchaudio": ta}
for variant in ("dft_gemm", "fft"):
    m = fe.LFCCDelta(**helpers.LFCC_CFG, variant=variant)
    got[variant] = m.engine.fbank_energies(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)

level_db = 10.0 * np.log10(np.maximum(truth, 1e-300) / truth.max(axis=(1, 2), keepdims=True))
edges = [0, -20, -40, -60, -70, -80, -100, -400]
out = {"bins_db_below_utterance_max": [f"{edges[i]}..{edges[i + 1]}" for i in range(len(edges) - 1)], "rows": 6,
       "what": "relative error |E - E64| / E64 of the filterbank energies; rms and max per level bin; the -80 dB top_db floor "
               "clamps everything below it"}
for name, e in got.items():
    rel = np.abs(e - truth) / np.maximum(truth, 1e-300)
    rms, mx, cnt = [], [], []
    for i in range(len(edges) - 1):
        sel = (level_db <= edges[i]) & (level_db > edges[i + 1]) & (truth > 0)
        cnt.append(int(sel.sum()))
        rms.append(float(np.sqrt(np.mean(rel[sel] ** 2))) if sel.any() else None)
        mx.append(float(rel[sel].max()) if sel.any() else None)
    # error relative to the frame's own maximum energy (what a fixed-point-like error model predicts to be flat)
    frame_max = truth.max(axis=1, keepdims=True)
    out[name] = {"count": cnt, "rel_rms": rms, "rel_max": mx,
                 "abs_over_frame_max_rms": float(np.sqrt(np.mean((np.abs(e - truth) / np.maximum(frame_max, 1e-300)) ** 2)))}
d = os.path.join(ROOT, "gpurun_out")
os.makedirs(d, exist_ok=True)
json.dump(out, open(os.path.join(d, "r2_s2_error_profile.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
