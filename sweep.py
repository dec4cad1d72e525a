#!/usr/bin/env python
"""BASELINE config 4: ASVspoof2019-LA-eval-sized sweep (71,237 synthetic 4 s utterances) sharded over the
GPUs of one box, LFCC+delta+delta-delta front-end -> maze5 classifier (seeded weights) -> score gather -> EER.

    python sweep.py                                                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 sweep.py                                      # one rank per GPU, NCCL gather

Rank 0 prints ONE JSON line: EER / min-DCF, a SHA-256 of the gathered per-utterance feature checksums (identical
for every world size and batch size), a SHA-256 of the score vector (stock cuDNN classifier: identical for
identical batch shapes), front-end-only and end-to-end utterances/s (max over ranks of the device time of the local shard).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=None, help="sweep size (default 71,237)")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fmsl", action="store_true", help="maze5-FMSL head instead of plain maze5")
    ap.add_argument("--variant", default="auto")
    ap.add_argument("--scores-out", default=None, help='write "<utt_id> <score>" lines (maze5.py:428)')
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import b200_frontend as fe
    from importlib import import_module
    sweep = import_module("audio-deepfake-detection-fmsl_b200.sweep")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    n_total = args.utterances or sweep.N_EVAL
    n_bona = min(sweep.N_BONAFIDE, max(1, n_total // 10)) if args.utterances else sweep.N_BONAFIDE
    frontend = fe.LFCCDelta(16000, n_filter=20, n_lfcc=20,
                            speckwargs=dict(n_fft=512, win_length=320, hop_length=160), variant=args.variant)
    scorer = fe.MazeScorer(fe.LFCC_FILTS, fmsl=args.fmsl)
    fe.fill_deterministic(scorer, sweep.SEED)
    scorer.to(device)

    # warm-up (kernel images, cuDNN plans, NCCL communicator), then the sweep itself
    warm = sweep.synthetic_block(0, device)[: args.batch]
    for _ in range(2):
        scorer(frontend(warm))
    if world > 1:
        dist.barrier()
    r = sweep.run_sweep(frontend, scorer, device, n_total=n_total, n_bonafide=n_bona, batch=args.batch,
                        rank=rank, world_size=world)
    t = torch.tensor([r["frontend_ms"], r["frontend_ms"] + r["classifier_ms"], r["wall_s"] * 1e3],
                     device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fe_ms, dev_ms, wall_ms = t.tolist()
    if rank == 0:
        if args.scores_out:
            fe.write_score_file(args.scores_out, [f"SYN_E_{i:07d}" for i in range(n_total)], r["scores"])
        print(json.dumps({
            "workload": f"config 4: {n_total} synthetic 64600-sample utterances ({n_bona} bonafide), "
                        f"LFCC+delta+delta-delta -> maze5{'-FMSL' if args.fmsl else ''} classifier -> gather -> EER",
            "n_gpus": world, "batch": args.batch, "variant": frontend.engine.resolved_variant(),
            "eer": r["eer"], "min_dcf": r["min_dcf"], "eer_threshold": r["eer_threshold"],
            "scores_sha256": r["scores_sha256"], "features_sha256": r["features_sha256"],
            "frontend_utt_per_s": n_total / (fe_ms * 1e-3), "frontend_ms": fe_ms,
            "frontend_plus_classifier_utt_per_s": n_total / (dev_ms * 1e-3), "device_ms": dev_ms,
            "sweep_wall_utt_per_s": n_total / (wall_ms * 1e-3), "wall_ms": wall_ms,
            "timing": "CUDA events per batch, summed per rank, max over ranks; wall = synthetic generation + "
                      "front-end + classifier + gather + EER",
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
