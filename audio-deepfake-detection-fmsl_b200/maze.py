"""Downstream consumer of the features: the maze5 classifier body behind the feature slot, inference only.

The reference builds its detector as ``sinc_conv -> first_bn -> SELU -> 5 x (residual block + squeeze-excite)
-> global average -> fc1 -> fc2 -> log-softmax`` (Thesis/01_Models/01_Baseline_Models/maze5.py:178-260) and, in
the FMSL variant, routes ``fc1``'s output through the FMSL projection and L2 normalisation before ``fc2``
(Thesis/01_Models/02_FMSL_Enhanced_Models/maze5_fmsl_standardized.py:302-332,
Thesis/06_Utilities/fmsl_advanced.py:257-304).  The front-end of this repository drops into the
``sinc_conv`` slot (maze5.py:241); ``MazeScorer`` is everything after that slot, evaluated the way
``model.eval()`` evaluates it (running BatchNorm statistics, no dropout, no SpecAugment), so that the sweep of
BASELINE config 4 and the ragged path of config 5 can be scored on a GPU box where the reference tree is absent.

It is stock PyTorch: the classifier is *unchanged downstream* (SURVEY.md 8a), not part of the accelerated
path.  Parameter names equal the reference's, so ``load_reference_state_dict`` accepts a checkpoint saved by
``maze5.py`` / ``maze5_fmsl_standardized.py`` (the ``sinc_conv.*`` entries are dropped: the slot is replaced).
``tests/golden/maze_golden.npz`` pins the arithmetic against the reference classes themselves.
"""
from __future__ import annotations

import zlib
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

# filts for a 60-channel LFCC+delta+delta-delta front-end (verified drop-in configuration)
LFCC_FILTS = (60, (128, 128), (128, 256))


class _Gate(nn.Module):
    """Squeeze-excite channel gate (maze5.py:148-161): time average -> c/16 -> c -> sigmoid."""

    def __init__(self, channels: int, reduction: int = 16) -> None:
        super().__init__()
        hidden = channels // reduction
        self.fc = nn.Sequential(nn.Linear(channels, hidden, bias=False), nn.ReLU(),
                                nn.Linear(hidden, channels, bias=False), nn.Sigmoid())

    def forward(self, x: Tensor) -> Tensor:
        return x * self.fc(x.mean(dim=2)).unsqueeze(2)


class _Stage(nn.Module):
    """Pre-activation residual stage (maze5.py:105-146) in eval mode: [BN-ReLU] conv3 BN-ReLU conv3 + skip
    (1x1 conv when the width or the stride changes), then AvgPool(2s-1, stride s, pad s-1) for stride s > 1."""

    def __init__(self, c_in: int, c_out: int, first: bool, stride: int) -> None:
        super().__init__()
        if not first:
            self.bn1 = nn.BatchNorm1d(c_in)
        self.conv1 = nn.Conv1d(c_in, c_out, 3, padding=1)
        self.bn2 = nn.BatchNorm1d(c_out)
        self.conv2 = nn.Conv1d(c_out, c_out, 3, padding=1)
        if c_in != c_out or stride != 1:
            self.conv_downsample = nn.Conv1d(c_in, c_out, 1)
        self.first, self.stride = first, stride

    def forward(self, x: Tensor) -> Tensor:
        y = x if self.first else F.relu(self.bn1(x))
        y = self.conv2(F.relu(self.bn2(self.conv1(y))))
        y = y + (self.conv_downsample(x) if hasattr(self, "conv_downsample") else x)
        if self.stride > 1:
            y = F.avg_pool1d(y, 2 * self.stride - 1, self.stride, self.stride - 1)
        return y


class _FMSLProjection(nn.Module):
    """The part of ``AdvancedFMSLSystem.forward`` the maze5-FMSL model consumes at inference
    (fmsl_advanced.py:277-283): Linear -> BN -> ReLU (-> Dropout, identity in eval) -> L2 normalise.
    ``prototypes`` / ``weight`` / ``temperature`` exist only so that checkpoints load."""

    def __init__(self, dim: int, n_classes: int, n_prototypes: int) -> None:
        super().__init__()
        self.projection = nn.Sequential(nn.Linear(dim, dim), nn.BatchNorm1d(dim), nn.ReLU(), nn.Dropout(0.1))
        self.prototypes = nn.Parameter(torch.zeros(n_prototypes, dim))
        self.weight = nn.Parameter(torch.zeros(n_classes, dim))
        self.temperature = nn.Parameter(torch.tensor(1.0))

    def forward(self, x: Tensor) -> Tensor:
        return F.normalize(self.projection(x), p=2.0, dim=1, eps=1e-12)


class FeatureSlot(nn.Module):
    """Adapter for the reference's slot contract ``(B,1,T) -> (B,C,T')`` (maze5.py:241-242: the result goes
    straight into ``BatchNorm1d(filts[0])``).  The transforms keep torchaudio's layout ``(..., C, n_frames)``,
    i.e. ``(B,1,C,n_frames)`` for the ``(B,1,T)`` tensor the models pass; the slot drops that channel axis:

        model.sinc_conv = FeatureSlot(LFCCDelta(16000, n_filter=20, n_lfcc=20, speckwargs=...))
    """

    def __init__(self, transform: nn.Module) -> None:
        super().__init__()
        self.transform = transform

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"the feature slot is called with (B,1,T), got {tuple(x.shape)}")
        return self.transform(x).squeeze(1)


class MazeScorer(nn.Module):
    """``features (B, filts[0], n_frames) -> log-softmax (B, nb_classes)``; with a ``frontend`` module also
    ``waveform (B,T) / (B,1,T) -> log-softmax`` exactly like ``Model5_...forward`` (maze5.py:233-260)."""

    def __init__(self, filts: Sequence = LFCC_FILTS, nb_fc_node: int = 1024, nb_classes: int = 2,
                 fmsl: bool = False, fmsl_n_prototypes: int = 3, frontend: Optional[nn.Module] = None) -> None:
        super().__init__()
        c0, (c1, c2), (c3, c4) = int(filts[0]), filts[1], filts[2]
        self.first_bn = nn.BatchNorm1d(c0)
        self.block0 = _Stage(c0, c0, first=True, stride=1)
        self.se0 = _Gate(c0)
        widths = [(c0, c1), (c1, c2), (c2, c3), (c3, c4)]   # maze5.py:211-222
        self.res_blocks = nn.ModuleList(_Stage(a, b, first=False, stride=2) for a, b in widths)
        self.se_blocks = nn.ModuleList(_Gate(b) for _, b in widths)
        self.fc1 = nn.Linear(c4, nb_fc_node)
        self.fc2 = nn.Linear(nb_fc_node, nb_classes)
        self.fmsl_system = _FMSLProjection(nb_fc_node, nb_classes, fmsl_n_prototypes) if fmsl else None
        self.frontend = frontend
        self.eval()

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("MazeScorer is inference-only (the reference's training loop is out of scope)")
        return super().train(False)

    def classify(self, feats: Tensor) -> Tensor:
        # fp32 convolutions (cuDNN would otherwise take TF32 on a B200 and move the scores by ~1e-2)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return self._classify(feats)

    def _classify(self, feats: Tensor) -> Tensor:
        y = F.selu(self.first_bn(feats))
        y = self.se0(self.block0(y))
        for stage, gate in zip(self.res_blocks, self.se_blocks):
            y = gate(stage(y))
        y = self.fc1(y.mean(dim=2))
        if self.fmsl_system is not None:
            y = self.fmsl_system(y)
        return F.log_softmax(self.fc2(y), dim=1)

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        if self.frontend is None:
            return self.classify(x)
        if x.dim() == 1:
            x = x.reshape(1, 1, -1)
        elif x.dim() == 2:
            x = x.unsqueeze(1)          # (B,T) -> (B,1,T): per-utterance top_db, as the models feed the slot
        feats = self.frontend(x)
        return self.classify(feats.squeeze(1) if feats.dim() == 4 else feats)

    @torch.no_grad()
    def scores(self, x: Tensor) -> Tensor:
        """Bonafide score per utterance: column 1 of the log-softmax (maze5.py:425)."""
        return self.forward(x)[:, 1]

    def load_reference_state_dict(self, state: Dict[str, Tensor]) -> None:
        """Load a checkpoint of the reference model; entries of the replaced slot and of training-only
        members are ignored, everything else must match."""
        skip = ("sinc_conv.", "criterion.", "focal_loss.", "spec_augment.", "frontend.")
        own = {k: v for k, v in state.items() if not k.startswith(skip)}
        missing, unexpected = self.load_state_dict(own, strict=False)
        missing = [k for k in missing if not k.startswith("frontend.")]
        if missing or unexpected:
            raise KeyError(f"checkpoint does not match: missing {missing}, unexpected {list(unexpected)}")


def fill_deterministic(module: nn.Module, seed: int = 1234) -> None:
    """Give every parameter / buffer a value that depends only on its NAME, its shape and ``seed`` — the
    same values land in the reference model and in ``MazeScorer`` whatever the construction order, so
    golden logits can be regenerated from nothing but the seed (tests/golden/make_maze_golden.py)."""
    with torch.no_grad():
        for name, t in sorted(module.state_dict().items()):
            if name.startswith(("sinc_conv.", "frontend.", "criterion.")) or not t.dtype.is_floating_point:
                continue
            rs = np.random.RandomState((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
            if name.endswith("running_var"):
                v = rs.uniform(0.5, 1.5, t.shape)
            elif name.endswith("running_mean") or name.endswith(".bias"):
                v = 0.1 * rs.standard_normal(t.shape)
            elif t.dim() <= 1:
                v = rs.uniform(0.8, 1.2, t.shape)       # BatchNorm scales, temperature
            else:
                v = rs.standard_normal(t.shape) / np.sqrt(np.prod(t.shape[1:]))
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(t.shape))
