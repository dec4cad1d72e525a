"""Times the feature path's two kernels separately on the LFCC configuration (energies kernel, then the whole
forward) so the tail kernel's share can be read off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200_frontend as fe
import helpers
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
m = fe.LFCCDelta(**helpers.LFCC_CFG, variant="dft_gemm")
x = (0.1 * torch.randn(R, 64600, device="cuda")).clamp_(-1, 1)
def timed(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n
te = timed(lambda: m.engine.fbank_energies(x))
tf = timed(lambda: m(x))
print("rows %d: energies %.3f ms (%.2f M utt/s), forward %.3f ms (%.2f M utt/s), tail share %.3f ms" % (R, te, R / te / 1e3, tf, R / tf / 1e3, tf - te))
