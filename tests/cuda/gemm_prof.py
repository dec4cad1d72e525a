"""Small driver for ncu: a few launches of the energies kernel of one variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200_frontend as fe
import helpers
variant = sys.argv[1] if len(sys.argv) > 1 else "dft_gemm"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 592
N = int(sys.argv[3]) if len(sys.argv) > 3 else 5     # timed launches (a long run shows the sustained clocks)
m = fe.LFCCDelta(**helpers.LFCC_CFG, variant=variant)
x = (0.1 * torch.randn(R, 64600, device="cuda")).clamp_(-1, 1)
for _ in range(3):
    e = m.engine.fbank_energies(x)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(N):
    e = m.engine.fbank_energies(x)
t1.record()
torch.cuda.synchronize()
print(variant, R, "rows: %.3f ms per launch" % (t0.elapsed_time(t1) / N), "-> %.0f utt/s" % (R / (t0.elapsed_time(t1) / (N * 1e3))))
