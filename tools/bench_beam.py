#!/usr/bin/env python
"""Beam-search throughput (BASELINE config #3 shape: beam k = 3 over 65 536 synthetic embeddings, batch-sharded over the GPUs of one
box, final gather of token ids).  Not the driver's headline bench - a secondary measurement.
    python tools/bench_beam.py [--total 65536] [--topk 3] [--steps K] [--guided W]      or under torchrun for N GPUs
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=65536, help="embeddings over all GPUs")
ap.add_argument("--topk", type=int, default=3); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--guided", type=int, default=0, help="number of synthetic guide nouns (0 = unguided)")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from novic_b200 import synth, default_decoder
from novic_b200.dist import gather_generation, shard_bounds
dims = synth.DecoderDims()
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(dev)
b0, b1 = shard_bounds(args.total, world, rank)
embed = synth.synth_embeddings(b1 - b0, seed=1234 + rank).to(dev)
guide = synth.synth_guide_targets(args.guided, dims, seed=33, first_pool=200).to(dev) if args.guided else None

def step():
    tok, pad, score = model.generate_beam(embed, args.topk, 1.0, 0.0, None, False, 0.0, guide, False)
    if world > 1:
        tok, pad, score = gather_generation(tok, pad, score, args.total, gen_len=dims.token_length - 1)
    return tok

with torch.inference_mode():
    for _ in range(args.warmup): step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps): tok = step()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    assert tok.shape[0] == args.total
    print(json.dumps({"metric": f"labels/sec (beam k={args.topk}{', guided ' + str(args.guided) + ' nouns' if args.guided else ''})", "value": args.total / (ms.item() * 1e-3),
                      "unit": "labels/s", "n_gpus": world, "ms_per_step": ms.item(), "total_embeddings": args.total, "per_gpu": b1 - b0, "topk": args.topk}))
if world > 1: dist.destroy_process_group()
