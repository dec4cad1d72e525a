"""Sharding, score gather (gloo, world_size 2) and EER of the evaluation sweep.  CPU only."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from helpers import ROOT
from oracle import frontend_oracle as O


def test_shard_range_covers_everything(fe):
    for n, w in [(71237, 8), (71237, 4), (71237, 2), (71237, 1), (5, 8), (64, 3)]:
        spans = [fe.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        per = -(-n // w)
        assert all(hi - lo <= per for lo, hi in spans)
    with pytest.raises(ValueError):
        fe.shard_range(10, 3, 2)


def test_eer_matches_sklearn_and_oracle(fe):
    from sklearn.metrics import roc_curve
    rs = np.random.RandomState(3)
    y = np.r_[np.ones(735), np.zeros(6388)].astype(int)  # ASVspoof LA eval proportions / 10
    s = np.round(rs.randn(y.size) + 1.2 * y, 3)
    fpr, tpr, thr = roc_curve(y, s)
    fnr = 1 - tpr
    i = np.nanargmin(np.absolute(fnr - fpr))
    eer, dcf, t = fe.eer_min_dcf(y, s)
    assert eer == fpr[i] and dcf == min(fnr + fpr) and t == thr[i]
    assert (eer, dcf, t) == O.eer_min_dcf(y, s)
    with pytest.raises(ValueError):
        fe.eer_min_dcf(np.ones(10), rs.randn(10))


def test_score_file_format(fe, tmp_path):
    p = tmp_path / "scores.txt"
    fe.write_score_file(str(p), ["LA_E_1", "LA_E_2"], [-0.5, -1.25])
    assert p.read_text() == "LA_E_1 -0.5\nLA_E_2 -1.25\n"


def test_single_process_gather(fe):
    s = torch.arange(7, dtype=torch.float32)
    assert torch.equal(fe.gather_scores(s, 7), s)


_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["B200FE_ROOT"])
    import b200_frontend as fe
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 1001
    rs = np.random.RandomState(7)
    all_scores = torch.from_numpy(rs.randn(n).astype(np.float32))
    lo, hi = fe.shard_range(n, rank, world)
    full = fe.gather_scores(all_scores[lo:hi].clone(), n)
    assert torch.equal(full, all_scores), "gathered scores differ from the unsharded vector"
    labels = (rs.rand(n) < 0.1).astype(int)
    eer = fe.eer_min_dcf(labels, full.numpy())
    ref = fe.eer_min_dcf(labels, all_scores.numpy())
    assert eer == ref
    dist.barrier()
    dist.destroy_process_group()
    print("ok", rank)
""")


@pytest.mark.timeout(300)
def test_gloo_world2_gather_equals_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, B200FE_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
