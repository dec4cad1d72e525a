"""Quick on-GPU check of the DFT-GEMM variant's energies against the float64 oracle and the FFT variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_frontend as fe
import helpers
from oracle import frontend_oracle as O, synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = np.concatenate([synth.s1_noise(max(1, R - 2)), synth.s2_speechlike(1), synth.s3_edge()[1:2]], 0)[:R]
ref = O.apply_fbank(O.power_spectrogram(x.astype(np.float64), 512, 320, 160), O.linear_fbanks(257, 0, 8000, 20, 16000).astype(np.float64))
xd = torch.from_numpy(x).cuda()
for v in ("fft", "dft_gemm"):
    m = fe.LFCCDelta(**helpers.LFCC_CFG, variant=v)
    e = m.engine.fbank_energies(xd)
    torch.cuda.synchronize()
    e = e.cpu().numpy()
    for r in range(R):
        sc = max(ref[r].max(), 1e-30)
        print(v, "row", r, "rel-to-max err %.3e" % (np.abs(e[r] - ref[r]).max() / sc), "nan" if not np.isfinite(e[r]).all() else "")
    out = m(xd).cpu().numpy()
    print(v, "features vs oracle f32:", helpers.feat_err(out, O.lfcc(x, deltas=2)))
print("ok")
