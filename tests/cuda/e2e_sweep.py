"""End-to-end (host int16 PCM in, host features out) throughput against chunk size and stream count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200_frontend as fe
import helpers
B = 4096
m = fe.LFCCDelta(**helpers.LFCC_CFG)
pcm = torch.randint(-3000, 3000, (B, 64600), dtype=torch.int16).pin_memory()
out = torch.empty((B, 60, 404), dtype=torch.float32).pin_memory()
for chunk in (128, 256, 512, 1024):
    for ns in (2, 3, 4):
        for _ in range(2):
            m.forward_host(pcm, out, chunk_rows=chunk, n_streams=ns)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 4
        for _ in range(n):
            m.forward_host(pcm, out, chunk_rows=chunk, n_streams=ns)
        dt = (time.perf_counter() - t0) / n
        print(f"chunk_rows {chunk:5d} streams {ns}: {dt*1e3:7.2f} ms/step  {B/dt/1e3:7.1f} k utt/s  H2D {B*64600*2/dt/1e9:5.1f} GB/s  D2H {B*60*404*4/dt/1e9:5.1f} GB/s")
