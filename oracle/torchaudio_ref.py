"""The reference CPU path itself: torchaudio 2.11.0 transforms, called unmodified.

TEST INFRASTRUCTURE / BASELINE ONLY (see ``oracle/__init__.py``).  The reference repo's spectral
front-end arithmetic is the torchaudio package it depends on (``import torchaudio.transforms as T``
at Thesis/01_Models/01_Baseline_Models/maze5.py:32); this module wires the transforms together the
way SURVEY.md 8(d) / BASELINE.md section 3 define the CPU path:

    LFCC((B,1,T)) -> ComputeDeltas -> ComputeDeltas -> cat       -> (B, 60, n_frames)
    MelSpectrogram((B,1,T)) [-> AmplitudeToDB('power', 80)]      -> (B, n_mels, n_frames)

A ``(B,1,T)`` input is used so that ``top_db`` clamps per utterance.
"""
from __future__ import annotations

from typing import Optional

import torch
import torchaudio.functional as AF
import torchaudio.transforms as T


class LFCCDeltaRef(torch.nn.Module):
    def __init__(self, sample_rate=16000, n_filter=20, n_lfcc=20, n_fft=512, win_length=320,
                 hop_length=160, log_lf=False, deltas=2, preemph: Optional[float] = None, **kw):
        super().__init__()
        self.lfcc = T.LFCC(sample_rate=sample_rate, n_filter=n_filter, n_lfcc=n_lfcc, log_lf=log_lf,
                           speckwargs=dict(n_fft=n_fft, win_length=win_length, hop_length=hop_length), **kw)
        self.delta = T.ComputeDeltas(win_length=5, mode="replicate")
        self.deltas = deltas
        self.preemph = preemph

    @torch.no_grad()
    def forward(self, wave: torch.Tensor) -> torch.Tensor:
        """``wave`` is ``(B,T)`` or ``(B,1,T)`` float32 on CPU; returns ``(B, n_lfcc*(1+deltas), F)``."""
        if wave.dim() == 2:
            wave = wave.unsqueeze(1)
        if self.preemph is not None:
            wave = AF.preemphasis(wave, self.preemph)
        c = self.lfcc(wave).squeeze(1)
        parts = [c]
        for _ in range(self.deltas):
            parts.append(self.delta(parts[-1]))
        return torch.cat(parts, dim=-2).contiguous()


class LogMelRef(torch.nn.Module):
    def __init__(self, sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, log: Optional[str] = "db",
                 win_length=None, **kw):
        super().__init__()
        self.mel = T.MelSpectrogram(sample_rate=sample_rate, n_fft=n_fft, win_length=win_length,
                                    hop_length=hop_length, n_mels=n_mels, **kw)
        self.to_db = T.AmplitudeToDB("power", top_db=80.0)
        self.log = log

    @torch.no_grad()
    def forward(self, wave: torch.Tensor) -> torch.Tensor:
        if wave.dim() == 2:
            wave = wave.unsqueeze(1)
        m = self.mel(wave)
        if self.log == "db":
            m = self.to_db(m)
        elif self.log == "log":
            m = torch.log(m + 1e-6)
        return m.squeeze(1).contiguous()
