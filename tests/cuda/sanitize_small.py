"""Small driver for compute-sanitizer: every kernel family once on small inputs (dense, ragged, pre-emphasis, mel,
spectrogram sizes with and without the warp-FFT fast path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b200_frontend as fe
import helpers
from oracle import synth

dev = "cuda"
x = torch.from_numpy(synth.s1_noise(5)).to(dev)
for variant in ("fft", "auto"):
    m = fe.LFCCDelta(**helpers.LFCC_CFG, variant=variant)
    m(x)
    flat, offsets, lengths = synth.s4_ragged(6)
    m.forward_ragged(*(torch.from_numpy(a).to(dev) for a in (flat, offsets, lengths)), 64600)
    fe.LFCCDelta(**helpers.LFCC_CFG, variant=variant, preemphasis=0.97)(x)
for log in ("db", "log", None):
    fe.MelSpectrogram(**helpers.MEL_CFG, log=log)(x)
for n_fft, win, hop in ((256, 128, 64), (512, 320, 160), (1024, 400, 200), (2048, 2048, 512), (128, 128, 32), (512, 400, 161)):
    fe.Spectrogram(n_fft=n_fft, win_length=win, hop_length=hop)(x[:2, :9000].contiguous())
fe.LFCC(16000, n_filter=128, n_lfcc=40, speckwargs=dict(n_fft=512, win_length=320, hop_length=160))(x[:2])
fe.ComputeDeltas()(torch.randn(3, 20, 404, device=dev))
torch.cuda.synchronize()
print("sanitize_small ok")
