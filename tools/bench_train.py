#!/usr/bin/env python
"""Training-step throughput (BASELINE config #4 shape: 1024 text-embedding samples per GPU, C = 16 targets, embedding noise,
AdamW, gradient clipping, NCCL all-reduce of the gradients).  Not the driver's headline bench - a secondary measurement.
    python tools/bench_train.py [--steps K]            or under torchrun for N GPUs
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=1024); args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from novic_b200 import synth, default_decoder, EmbeddingNoise
torch.manual_seed(1234 + rank)   # noise and dropout seeds; the loss trajectory at lr 1.5e-3 from random weights is still sensitive to fp32 atomic order
from novic_b200.dist import train_step
dims = synth.DecoderDims()
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(dev).train()   # input / layer dropout 0.1 (train.yaml defaults)
decay = [p for p in model.parameters() if p.dim() >= 2]; no_decay = [p for p in model.parameters() if p.dim() < 2]
opt = torch.optim.AdamW([{'params': no_decay, 'weight_decay': 0.0}, {'params': decay, 'weight_decay': 0.1}], lr=1.5e-3, betas=(0.9, 0.95), fused=True)
noise = EmbeddingNoise.create("GaussElemUniformAngle", 1024, 3.25, 45.0, 75.0, 0.0, 0.15)
B = args.batch
embed0 = synth.synth_embeddings(B, seed=100 + rank).to(dev)
tgt, pad = synth.synth_targets(B, dims, seed=200 + rank); tgt, pad = tgt.to(dev), pad.to(dev)
losses = []
def step():
    loss, ncorrect, ntok, norm = train_step(model, opt, embed0.clone(), tgt, pad, None, noise=noise, gradient_clip=1.0)
    return loss
for _ in range(args.warmup): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
for _ in range(args.steps): losses.append(step())
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = (time.perf_counter() - t0) / args.steps
# device time of forward + backward alone (one C-ABI call), the part this repo's kernels own; the rest of the step is torch
# (noise, clip_grad_norm_ with its host sync, fused AdamW) and NCCL
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
model.zero_grad(set_to_none=True)
ev[0].record()
for _ in range(args.steps):
    _, _, ls, lb, _ = model(embed0, tgt, pad, None, True, True, False, None)
    ls.backward()
    model.zero_grad(set_to_none=True)
ev[1].record(); torch.cuda.synchronize()
fwd_bwd_ms = ev[0].elapsed_time(ev[1]) / args.steps
if rank == 0:
    from novic_b200.dist import GradBucket
    print(json.dumps({"fwd_bwd_ms": fwd_bwd_ms, "in_place_allreduces": GradBucket.in_place_reductions, "metric": "training samples/sec (teacher-forced step, noise + fwd + bwd + allreduce + clip + AdamW)", "value": B * world / dt, "unit": "samples/s",
                      "n_gpus": world, "ms_per_step": dt * 1e3, "batch_per_gpu": B, "loss_first": losses[0].item(), "loss_last": losses[-1].item(), "dropout": "input 0.1, layer 0.1"}))
if world > 1: dist.destroy_process_group()
