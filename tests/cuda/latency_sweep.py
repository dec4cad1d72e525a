"""Latency of one LFCC+delta+delta-delta forward call against the batch size (the reference's evaluation loop uses
batch 32, Maze5_eval.py:657; training batch 12, maze5.py:536), eager and replayed from a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200_frontend as fe
import helpers

m = fe.LFCCDelta(**helpers.LFCC_CFG)
for B in (1, 12, 32, 64, 256, 1024, 4096):
    x = (0.1 * torch.randn(B, 1, 64600, device="cuda")).clamp_(-1, 1)
    for _ in range(5):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if B <= 256 else 50
    e0.record()
    for _ in range(n):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / n
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        m(x)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            y = m(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / n
    print(f"B={B:5d}  eager {eager*1e3:8.1f} us/call ({B/eager*1e3:10.0f} utt/s)   graph {graph*1e3:8.1f} us/call ({B/graph*1e3:10.0f} utt/s)")
