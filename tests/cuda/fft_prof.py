"""Small driver for ncu / timing: the energies kernel of the FFT variant on the mel or the LFCC configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200_frontend as fe
import helpers
which = sys.argv[1] if len(sys.argv) > 1 else "mel"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
m = fe.MelSpectrogram(**helpers.MEL_CFG, log="db") if which == "mel" else fe.LFCCDelta(**helpers.LFCC_CFG, variant="fft")
x = (0.1 * torch.randn(R, 64600, device="cuda")).clamp_(-1, 1)
for _ in range(3):
    e = m.engine.fbank_energies(x)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    e = m.engine.fbank_energies(x)
t1.record()
torch.cuda.synchronize()
print(which, R, "rows: %.3f ms per launch" % (t0.elapsed_time(t1) / 5), "-> %.0f utt/s" % (R / (t0.elapsed_time(t1) / 5e3)))
