#!/usr/bin/env python
"""Benchmark of the NOVIC decoder hot path: labels/sec, greedy decode, batch 4096 per GPU (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one greedy decode (prefix prefill + 15 autoregressive steps, KV-cached) of one batch of 4096 synthetic
unit-norm 1024-d embeddings per GPU through novic_b200.PrefixedIterDecoder (random-init default decoder,
config/train.yaml:224-308).  Rank 0 prints ONE JSON line (see the keys below).  `--impl reference` times the CPU
port of the reference's own algorithm (oracle/, re-forward schedule without KV cache) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 4096
METRIC = "labels/sec (greedy, batch 4096)"
UNIT = "labels/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="embeddings per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=256, help="embeddings in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "BASELINE configs[1]: random-init default EmbeddingDecoder (E=512, FFN 128, L=6, 8 heads, P=4, V=6912, Cmax=16), "
                    f"{args.batch} synthetic unit-norm 1024-d embeddings per GPU, greedy decode (15 steps, tau=1, alpha=0), bf16 operands / fp32 accumulate",
        "batch_per_gpu": args.batch, "global_batch": args.batch * world, "gen_steps": 15, "parallelism": f"dp{world} (batch shards, final id gather)",
        "l2": "L2 flushed (256 MiB write) between timed iterations; per-step working set (KV cache 0.96 GB) also exceeds L2",
    }


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference algorithm (no KV cache, re-forward every step), all host threads
# ----------------------------------------------------------------------------------------------------------------
def cpu_greedy_rate(sample: int, repeats: int = 1, warmup: int = 1):
    from novic_b200 import synth
    from oracle import novic_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    dims = synth.DecoderDims()
    sd = synth.synth_state_dict(dims, seed=1)
    cfg = orc.cfg_from_state_dict(sd)
    embed = synth.synth_embeddings(sample, seed=1234)
    times = []
    with torch.inference_mode():
        for _ in range(warmup):
            orc.generate_greedy(cfg, sd, embed[: min(16, sample)], 1.0, 0.0)
        for _ in range(repeats):
            t0 = time.perf_counter()
            orc.generate_greedy(cfg, sd, embed, 1.0, 0.0)
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    sample = args.cpu_sample
    times, cores = cpu_greedy_rate(sample, repeats=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    ms = statistics.mean(times) * 1e3
    value = sample / (ms / 1e3)
    desc = f"greedy decode of {sample} embeddings per step (bounded sample of the 4096-embedding workload), fp32, torch CPU ops, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def wait_started(self, timeout: float = 10.0) -> None:
        """nvidia-smi takes a while to attach; wait for its first sample so start-up does not perturb the timed region."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def mark(self) -> None:
        self.rows.clear()   # keep only samples taken during the timed region

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# algorithmic work per kernel class for one greedy decode of B embeddings
# ----------------------------------------------------------------------------------------------------------------
def algorithmic_work(B: int, dims) -> dict:
    F, E, K, L, P, V, G = dims.embed_dim, dims.hidden_dim, dims.ffn_dim, dims.num_layers, dims.prefix_len, dims.vocab_size, dims.token_length - 1
    rows = B * (P + G - 1)                       # residual rows that pass through the layers (prefill P + G-1 decode steps)
    attn_bytes = 0
    attn_bytes += B * L * (P * 2 * E * 2 + 2 * P * E * 2)                    # prefill: K,V rows once + q in + out
    for s in range(P, P + G - 1):                                              # decode query at position s sees s+1 keys
        attn_bytes += B * L * ((s + 1) * 2 * E * 2 + 2 * E * 2)
    return {
        "embed_prep": ("hbm", B * F * (4 + 2)),
        "prefix_gemm": ("tensor", 2 * B * F * P * E),
        "qkv_gemm": ("tensor", 2 * rows * E * 3 * E * L),
        "attention": ("hbm", attn_bytes),
        "outproj_gemm": ("tensor", 2 * rows * E * E * L),
        "ffn1_gemm": ("tensor", 2 * rows * E * K * L),
        "ffn2_gemm": ("tensor", 2 * rows * K * E * L),
        "logits_gemm": ("tensor", 2 * B * G * E * V),
        "select": ("hbm", B * G * (-(-V // 64) * 32 + E * (4 + 4 + 2))),
    }


def gemm_hbm_bytes(B: int, dims) -> dict:
    """Algorithmic HBM bytes per decode of the GEMM classes: activation rows read and written once (bf16 operands 2 B, fp32 residual
    stream 4 B per element); weights (25 MB in all) stay L2-resident across the launches of a decode and are not counted.  A GEMM
    class is reported against whichever roof - tensor or HBM - takes longer for its algorithmic work (roof_entry)."""
    F, E, K, L, P, V, G = dims.embed_dim, dims.hidden_dim, dims.ffn_dim, dims.num_layers, dims.prefix_len, dims.vocab_size, dims.token_length - 1
    rows = B * (P + G - 1)
    return {
        "prefix_gemm": B * F * 2 + B * P * E * (4 + 2),                 # bf16 embedding in; fp32 residual rows + their LayerNorm-ed bf16 copy out
        "qkv_gemm": rows * L * (2 * E + 3 * 2 * E),                     # xn in; q, k, v out
        "outproj_gemm": rows * L * (2 * E + 4 * E + 4 * E + 2 * E),     # attention rows + residual in; residual + LN2 rows out
        "ffn1_gemm": rows * L * (2 * E + 2 * K),
        "ffn2_gemm": rows * L * (2 * K + 4 * E + 4 * E + 2 * E),
        "logits_gemm": B * G * (2 * E + -(-V // 64) * 32),              # final rows in; one 32-byte partial record per 64 vocabulary columns out
        # fused kernels keep their intermediates on chip: attention rows + residual in, residual + next LayerNorm rows out
        "fused_block": rows * L * (2 * E + 4 * E + 4 * E + 2 * E),
    }


def roof_entry(bound: str, amount: float, hbm_bytes, ms: float, iso_ms: float, peaks: dict) -> dict:
    """Roofline fields of one kernel class from its algorithmic work and its measured time per decode (ms in graph, iso_ms isolated).
    bound 'hbm': amount = bytes.  bound 'tensor': amount = FLOPs and, when hbm_bytes is given, the class is reported against the roof
    that takes longer for that work (the binding one); both fractions are kept as tensor_frac / hbm_frac."""
    def hbm(nbytes, t_ms):
        return nbytes / (t_ms * 1e-3) / 1e9
    def tensor(flops, t_ms):
        return flops / (t_ms * 1e-3) / 1e12
    if bound == "hbm":
        ach = hbm(amount, ms)
        return {"bound": "hbm", "achieved": ach, "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "isolated_frac": hbm(amount, iso_ms) / peaks["hbm_gbs"]}
    t_frac = tensor(amount, ms) / peaks["bf16_tflops_sustained"]
    out = {"bound": "tensor", "achieved": tensor(amount, ms), "unit": "TFLOP/s", "frac": t_frac,
           "isolated_frac": tensor(amount, iso_ms) / peaks["bf16_tflops_sustained"]}
    if hbm_bytes:
        h_frac = hbm(hbm_bytes, ms) / peaks["hbm_gbs"]
        out.update(tensor_frac=t_frac, hbm_frac=h_frac)
        if h_frac > t_frac:   # the HBM roof binds: moving the rows takes longer than the math
            out.update(bound="hbm", achieved=hbm(hbm_bytes, ms), unit="GB/s", frac=h_frac, isolated_frac=hbm(hbm_bytes, iso_ms) / peaks["hbm_gbs"])
    return out


def _timed_decodes(model, embed, flush, n):
    """Mean device time (CUDA events on the launching stream, L2 flushed before each) of n greedy decodes."""
    total = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / n


def kernel_breakdown(model, embed, flush, steps: int, peaks: dict, dims) -> dict:
    """Two device timings of every kernel class, both with CUDA events on the launching stream:

    in_graph  - the captured decode graph with the launches of every OTHER class dropped (novic_debug_keep_classes), minus the same
                graph with no kernel at all: the class's launches run back to back with their real arguments, programmatic dependent
                launch and the L2 state of the decode, exactly as inside the product's graph.  This is `achieved` / `frac`.
    isolated  - direct launches (no graph) with an event pair around every launch: includes one launch gap per kernel (~2-3 us on a
                ~10 us kernel) - the pessimistic bound, comparable to ncu's serialised per-launch times in profiles/."""
    import ctypes as C
    from novic_b200 import _abi
    lib = _abi.lib()
    st = model._state(embed.device)
    names = list(_abi.KERNEL_CLASSES)
    with torch.inference_mode():
        # isolated: event pair per launch
        _abi.check(lib.novic_set_use_graphs(st["handle"], 0))
        model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        torch.cuda.synchronize()
        _abi.check(lib.novic_kernel_timing(1))
        for _ in range(steps):
            model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        n = len(names)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _abi.check(lib.novic_kernel_times(ms, cnt, n))
        _abi.check(lib.novic_kernel_timing(0))
        _abi.check(lib.novic_set_use_graphs(st["handle"], 1))
        # in graph: one class at a time
        in_graph = {}
        try:
            _abi.check(lib.novic_debug_keep_classes(st["handle"], 1 << 31))      # no kernel class at all: memsets, copies, host sync
            _timed_decodes(model, embed, flush, 2)
            base = _timed_decodes(model, embed, flush, steps)
            for i, name in enumerate(names):
                if cnt[i] == 0:
                    continue
                _abi.check(lib.novic_debug_keep_classes(st["handle"], 1 << i))
                _timed_decodes(model, embed, flush, 2)
                in_graph[name] = _timed_decodes(model, embed, flush, steps) - base
        finally:
            _abi.check(lib.novic_debug_keep_classes(st["handle"], 0))
        model.generate(embed, False, True, 1.0, 0.0, None, None, False)          # re-capture the full graph
    work = algorithmic_work(embed.shape[0], dims)
    if cnt[names.index('ffn1_gemm')] == 0:      # fused feed-forward kernel: its launches do both GEMMs of the block
        work["ffn2_gemm"] = ("tensor", work["ffn1_gemm"][1] + work["ffn2_gemm"][1])
    if cnt[names.index('outproj_gemm')] == 0:   # fused block kernel (out-proj + LN2 + feed-forward + LN): timed as class ffn2_gemm
        work["ffn2_gemm"] = ("tensor", work["ffn2_gemm"][1] + work["outproj_gemm"][1])
    gbytes = gemm_hbm_bytes(embed.shape[0], dims)
    if cnt[names.index('ffn1_gemm')] == 0 or cnt[names.index('outproj_gemm')] == 0:
        gbytes["ffn2_gemm"] = gbytes["fused_block"]
    total = sum(max(v, 0.0) for v in in_graph.values()) or 1.0
    out = {}
    for i, name in enumerate(names):
        if cnt[i] == 0:
            continue
        # classes of a few microseconds per decode drown in the run-to-run noise of the subtraction: never report less than
        # 40 % of the isolated time (the largest in-graph gain seen on any class is 45 %)
        iso_ms = ms[i] / steps
        g_ms = iso_ms if iso_ms < 0.1 else max(in_graph[name], 0.4 * iso_ms)   # (below 0.1 ms per decode the subtraction is all noise: keep the isolated time)
        entry = {"ms_per_step": g_ms, "launches_per_step": cnt[i] // steps, "share": g_ms / total, "isolated_ms_per_step": ms[i] / steps}
        if name in work:
            bound, amount = work[name]
            entry.update(roof_entry(bound, amount, gbytes.get(name), g_ms, iso_ms, peaks))
        out[name] = entry
    if cnt[names.index('outproj_gemm')] == 0 and 'ffn2_gemm' in out:
        # decode path with the fused block kernel (out-proj + LN2 + feed-forward + LN in one cluster kernel): name it for what it is
        out['block_outproj_ffn'] = out.pop('ffn2_gemm')
    return out


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "bf16_tflops": p["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (sm_100a) device: novic_b200 has no CPU fallback. Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from novic_b200 import _abi, default_decoder, synth
    from novic_b200.dist import gather_generation_async
    dims = synth.DecoderDims()
    model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(dev)
    B = args.batch
    embed_host = synth.synth_embeddings(B, seed=1234 + rank).pin_memory()
    embed = embed_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _abi.lib()

    def step_device():
        if world == 1:
            tok, pad, _, _, _, score = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
            return tok, pad, score
        # sharded: decode and gather are enqueued without a host synchronisation; the one sync of the step reads the global early-exit length
        tok, pad, score, T = model.generate_async(embed, 1.0, 0.0)
        tok, pad, score, T = gather_generation_async(tok.unsqueeze(1), pad.unsqueeze(1), score.unsqueeze(1), B * world, dims.token_length - 1, T)
        t = int(T.item())
        return tok[:, :, :t], pad[:, :, :t], score

    from novic_b200.serve import GenerationPipeline
    gather = (lambda t, p, sc, T: gather_generation_async(t, p, sc, B * world, dims.token_length - 1, T)) if world > 1 else None
    pipeline = GenerationPipeline(model, "greedy", post=gather, emit=(rank == 0))

    def run_e2e(steps):
        """The public serving loop (novic_b200.serve.GenerationPipeline): every step copies its batch from pinned host memory to the
        device, decodes it (and gathers across ranks), and rank 0 reads ids / padding / scores back to the host; the copies of
        neighbouring steps overlap the decode.  Timed as a whole: K steps between two synchronised CUDA events."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = None
        for res in pipeline.run(embed_host for _ in range(steps)):
            out = res if res is not None else out
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    def timed(fn, steps):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n0 = lib.novic_launch_count()
        for i in range(steps):
            flush.zero_()                      # L2 flush, outside the timed span
            starts[i].record()
            out = fn()
            ends[i].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches = lib.novic_launch_count() - n0
        per_step = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        total_ms = sum(per_step)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timed.last_per_step = per_step
        return t.item(), launches, out

    with torch.inference_mode():
        sampler = ClockSampler(local_rank) if rank == 0 else None
        for _ in range(max(args.warmup, 3)):
            step_device()
        run_e2e(max(args.warmup, 3))
        if sampler:
            sampler.wait_started()
            sampler.mark()
        total_ms, launches, out = timed(step_device, args.steps)
        dev_steps = list(timed.last_per_step)
        e2e_ms, out_host = run_e2e(args.steps)
        clocks = sampler.stop() if sampler else None
    tok = out[0]
    assert tok.shape[0] == B * world and tok.shape[-1] == dims.token_length - 1

    if rank == 0:
        peaks = load_peaks()
        ms_per_step = total_ms / args.steps
        value = B * world * args.steps / (total_ms / 1e3)
        e2e_value = B * world * args.steps / (e2e_ms / 1e3)
        kernels = kernel_breakdown(model, embed, flush, min(args.steps, 5), peaks, dims)
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        d = kernels[dom]
        if os.path.isfile(tpath) and d.get("bound") == "hbm":
            entry = json.load(open(tpath)).get(dom)
            if entry:  # ncu-measured DRAM bytes of one launch / that launch's algorithmic bytes, applied to the average launch
                work = algorithmic_work(B, dims).get(dom, (None, 0))[1]
                traffic = entry["ratio_to_algorithmic"] * work / max(1, d["launches_per_step"])
        roofline = {"kernel": dom, "bound": d.get("bound"), "achieved": d.get("achieved"), "peak": peaks["hbm_gbs"] if d.get("bound") == "hbm" else peaks["bf16_tflops_sustained"],
                    "unit": d.get("unit"), "frac": d.get("frac"), "traffic": traffic, "peak_source": peaks["source"],
                    "how": "CUDA events on the launching stream around the captured decode graph with only this kernel class's launches kept (real arguments, PDL, L2 flushed "
                           "before each decode), minus the same graph with no kernels; achieved = algorithmic work of the class's launches / that duration, i.e. per-launch work / average "
                           "launch duration.  isolated_* = event pair around every direct launch (adds a launch gap per kernel; comparable to ncu's serialised times in profiles/)",
                    "kernels": kernels}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": B * dims.embed_dim * 4, "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in out_host)),
                    "api": "novic_b200.serve.GenerationPipeline.run (double-buffered H2D / D2H around PrefixedIterDecoder.generate; per-rank H2D, rank 0 reads the gathered result)"},
            "gpu_launches": int(launches), "roofline": roofline,
            "step_ms": {"device_median": statistics.median(dev_steps), "device_max": max(dev_steps)},
        }
        if not args.no_cpu_baseline:
            times, cores = cpu_greedy_rate(args.cpu_sample, repeats=1, warmup=1)
            line["cpu_baseline"] = {"value": args.cpu_sample / times[0], "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle port of the reference greedy path (no KV cache), {args.cpu_sample} of the {B} embeddings, fp32, {cores} torch threads, 1 run after warm-up"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
