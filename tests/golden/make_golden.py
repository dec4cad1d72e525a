"""Generates tests/golden/*.npz from the reference CPU path (torchaudio 2.11.0, the dependency the
reference's feature path lives in — see oracle/frontend_oracle.py header).  Run from the repo root:

    python tests/golden/make_golden.py

Inputs come from oracle/synth.py (numpy RandomState, seed 1234) and are NOT stored — the tests
regenerate them; only the reference outputs are committed.  `/root/reference` itself holds no code
for this path, so nothing is imported from it here.
"""
import os
import sys

import numpy as np
import torch
import torchaudio

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synth  # noqa: E402
from oracle.torchaudio_ref import LFCCDeltaRef, LogMelRef  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def full_rows():
    s3 = synth.s3_edge()
    return np.concatenate([synth.s1_noise(1), synth.s2_speechlike(1), s3[0:1], s3[2:3]], 0)


def short_rows(T=4000):
    s3 = synth.s3_edge(T)
    return np.concatenate([synth.s1_noise(2, T), synth.s2_speechlike(2, T), s3], 0)


def main():
    torch.set_num_threads(1)
    meta = dict(torch=torch.__version__, torchaudio=torchaudio.__version__)
    xf, xs = full_rows(), short_rows()
    out = {}
    out["lfcc_dd_full"] = LFCCDeltaRef()(torch.from_numpy(xf)).numpy()
    out["lfcc_dd_short"] = LFCCDeltaRef()(torch.from_numpy(xs)).numpy()
    out["lfcc_loglf_short"] = LFCCDeltaRef(log_lf=True, deltas=0)(torch.from_numpy(xs)).numpy()
    out["lfcc_default128_short"] = LFCCDeltaRef(n_filter=128, n_lfcc=40, deltas=1)(torch.from_numpy(xs)).numpy()
    out["lfcc_preemph_short"] = LFCCDeltaRef(preemph=0.97)(torch.from_numpy(xs)).numpy()
    out["mel_db_full"] = LogMelRef()(torch.from_numpy(xf[:2])).numpy()
    out["mel_power_short"] = LogMelRef(log=None)(torch.from_numpy(xs)).numpy()
    out["mel_log_short"] = LogMelRef(log="log")(torch.from_numpy(xs)).numpy()
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), **out)
    with open(os.path.join(HERE, "frontend_golden.meta.txt"), "w") as fh:
        for k, v in meta.items():
            fh.write(f"{k} {v}\n")
        for k, v in out.items():
            fh.write(f"{k} shape={v.shape} dtype={v.dtype}\n")
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
