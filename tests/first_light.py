"""GPU bring-up diagnostics: each stage runs in its own process (a device trap must not hide later stages).

    python tests/first_light.py            # run every stage, each under `timeout`
    python tests/first_light.py gemm       # one stage in-process
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["gemm", "fwd_l1_c1", "fwd_l1_c4", "fwd_full", "greedy", "beam", "noise", "timing"]


def _lib():
    from novic_b200 import _abi
    return _abi, _abi.lib()


def stage_gemm():
    _abi, lib = _lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ok = True
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (128, 128, 512), (256, 384, 512), (100, 200, 128), (16, 6912, 512), (4096, 1536, 512)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
        out = torch.full((M, N), float("nan"), device=dev)
        _abi.check(lib.novic_debug_gemm(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, 128, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t()
        err = (out - ref).abs()
        bad = ~(err < 1e-2 * max(1.0, K ** 0.5))
        print(f"gemm M={M} N={N} K={K}: max|d|={err.nan_to_num(1e9).max().item():.3e} ref|max|={ref.abs().max().item():.2f} bad={int(bad.sum())}/{M * N}", flush=True)
        if bad.any():
            ok = False
            idx = bad.nonzero()[:8].tolist()
            print("   first bad (row, col, got, want):", [(r, c, round(out[r, c].item(), 3), round(ref[r, c].item(), 3)) for r, c in idx])
            print("   bad rows histogram (mod 8):", torch.bincount(bad.nonzero()[:, 0] % 8, minlength=8).tolist(),
                  " bad cols histogram (mod 16):", torch.bincount(bad.nonzero()[:, 1] % 16, minlength=16).tolist())
            # does the output look like a K-permuted / partial-K product?
            for kk in range(64, K + 1, 64):
                part = a[:, :kk].float() @ w[:, :kk].float().t()
                print(f"   vs first {kk} of K: max|d|={(out - part).abs().nan_to_num(1e9).max().item():.3e}")
    code = C.c_uint32(0)
    lib.novic_watchdog(C.byref(code))
    print("watchdog:", hex(code.value))
    return ok


def _oracle_setup(num_layers=6, seed=2, **synth_kw):
    from novic_b200 import synth, default_decoder
    from oracle import novic_oracle as orc
    dims = synth.DecoderDims(num_layers=num_layers)
    sd = synth.synth_state_dict(dims, seed=seed, token_scale=0.25, jitter_norms=True, **synth_kw)
    cfg = orc.cfg_from_state_dict(sd)
    model = default_decoder(dims, sd, num_layers=num_layers).to("cuda:0")
    return dims, sd, cfg, orc, model


def _deblock(xb: torch.Tensor, rows: int) -> torch.Tensor:
    # inverse of xblk_off: [rows32/32, 128, 32, 4] -> [rows, 512]
    r32 = (rows + 31) // 32 * 32
    return xb[: r32 * 512].view(r32 // 32, 128, 32, 4).permute(0, 2, 1, 3).reshape(r32, 512)[:rows]


def _fwd(num_layers, B, Cc, pad_mode, dump=False):
    from novic_b200 import synth
    _abi, lib = _lib()
    dims, sd, cfg, orc, model = _oracle_setup(num_layers)
    embed = synth.synth_embeddings(B, seed=1234)
    tgt, pad = synth.synth_targets(B, dims, seed=5)
    tgt, pad = tgt[:, :Cc].contiguous(), pad[:, :Cc].contiguous()
    if not pad_mode:
        pad = None
    with torch.inference_mode():
        o_logits, o_ls, o_lb, o_cor = orc.forward_loss(cfg, sd, embed, tgt, pad, None)
        g = model(embed.cuda(), tgt.cuda(), None if pad is None else pad.cuda(), None, True, True, False, None)
    torch.cuda.synchronize()
    logits = g[0].cpu()
    valid = torch.ones_like(tgt, dtype=torch.bool) if pad is None else ~pad
    err = (logits - o_logits)[valid].abs()
    print(f"fwd L={num_layers} B={B} C={Cc} pad={pad_mode}: logits max|d|={err.max().item():.4f} mean|d|={err.mean().item():.5f} "
          f"(oracle |max|={o_logits.abs().max().item():.2f}, std={o_logits.std().item():.3f}) nan={int(torch.isnan(logits).sum())}", flush=True)
    print(f"   loss {g[2].item():.4f} vs {o_ls.item():.4f}; basis {int(g[3])} vs {int(o_lb)}; correct agree {(g[4].cpu() == o_cor)[valid].float().mean().item():.4f}")
    ok = bool(err.max().item() < 0.06)
    if dump or not ok:
        # intermediates of a 1-layer model: recompute with torch and compare each workspace buffer
        st = model._handles[0]
        ws = st["ws"]
        P, E, S = dims.prefix_len, 512, dims.prefix_len + Cc - 1
        def buf(name, nbytes):
            off = C.c_size_t(0)
            _abi.check(lib.novic_debug_ws_offset(st["handle"], B, 1, dims.max_seq_len, name.encode(), C.byref(off)))
            return ws[off.value: off.value + nbytes]
        rows = B * S
        e = torch.nn.functional.normalize(embed, dim=-1)
        ebf = buf("ebf", B * 1024 * 2).view(torch.bfloat16).view(B, 1024).float().cpu()
        print(f"   ebf max|d|={(ebf - e).abs().max().item():.4f}")
        if num_layers == 1:
            x0 = (e.bfloat16().float() @ sd["embed_mlp.mlp.0.weight"].bfloat16().float().t()).view(B, P, E)
            if Cc > 1:
                x0 = torch.cat((x0, sd["logits_linear.weight"][tgt[:, :-1]]), dim=1)
            x0 = x0 + sd["pos_embedding.embedding.weight"][:S]
            ln = lambda x, w: torch.nn.functional.layer_norm(x, (E,), w, None, 1e-5)
            h = ln(x0, sd["transformer.layers.0.norm1.weight"]).bfloat16().float()
            qkv = h @ sd["transformer.layers.0.self_attn.in_proj_weight"].bfloat16().float().t()
            q_ref = qkv[..., :E].reshape(rows, E)
            q = buf("q", rows * E * 2).view(torch.bfloat16).view(rows, E).float().cpu()
            print(f"   q max|d|={(q - q_ref).abs().max().item():.4f} (|max| {q_ref.abs().max().item():.2f})")
            kv = buf("kv", 2 * B * dims.max_seq_len * E * 2).view(torch.bfloat16).view(2, B, dims.max_seq_len, E).float().cpu()
            print(f"   k max|d|={(kv[0, :, :S] - qkv[..., E:2 * E]).abs().max().item():.4f}  v max|d|={(kv[1, :, :S] - qkv[..., 2 * E:]).abs().max().item():.4f}")
            bias = orc.attention_bias(cfg, S, torch.float32).view(1, 1, S, S)
            if pad is not None:
                kb, _ = orc.key_padding_bias(cfg, pad, S, torch.float32)
                bias = bias + kb.view(B, 1, 1, S)
            qh = qkv[..., :E].bfloat16().float().view(B, S, 8, 64).transpose(1, 2)
            kh = qkv[..., E:2 * E].bfloat16().float().view(B, S, 8, 64).transpose(1, 2)
            vh = qkv[..., 2 * E:].bfloat16().float().view(B, S, 8, 64).transpose(1, 2)
            att = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0 + bias, dim=-1)
            ao_ref = (att @ vh).transpose(1, 2).reshape(rows, E)
            ao = buf("ao", rows * E * 2).view(torch.bfloat16).view(rows, E).float().cpu()
            print(f"   attn-out max|d|={(ao - ao_ref).abs().max().item():.4f} (|max| {ao_ref.abs().max().item():.2f})")
            x1 = x0.view(rows, E) + ao_ref.bfloat16().float() @ sd["transformer.layers.0.self_attn.out_proj.weight"].bfloat16().float().t()
            h2 = ln(x1, sd["transformer.layers.0.norm2.weight"]).bfloat16().float()
            hb_ref = torch.nn.functional.gelu(h2 @ sd["transformer.layers.0.linear1.weight"].bfloat16().float().t())
            hb = buf("hb", rows * 128 * 2).view(torch.bfloat16).view(rows, 128).float().cpu()
            print(f"   ffn-hidden max|d|={(hb - hb_ref).abs().max().item():.4f} (|max| {hb_ref.abs().max().item():.2f})")
            x2 = x1 + hb_ref.bfloat16().float() @ sd["transformer.layers.0.linear2.weight"].bfloat16().float().t()
            xg = _deblock(buf("x", ((rows + 31) // 32 * 32) * E * 4).view(torch.float32), rows).cpu()
            print(f"   x (residual out) max|d|={(xg - x2).abs().max().item():.4f} (|max| {x2.abs().max().item():.2f})")
    return ok


def stage_fwd_l1_c1():
    return _fwd(1, 8, 1, False, dump=True)


def stage_fwd_l1_c4():
    return _fwd(1, 8, 4, True, dump=True)


def stage_fwd_full():
    a = _fwd(6, 24, 16, True)
    b = _fwd(6, 200, 16, False)
    return a and b


def stage_greedy():
    from novic_b200 import synth
    ok = True
    for eos in (False, True):
        dims, sd, cfg, orc, model = _oracle_setup(6)
        if eos:
            sd = synth.make_eos_friendly(sd, dims, beta=0.8)
            model.load_state_dict(sd)
        embed = synth.synth_embeddings(32, seed=1234)
        for graphs in (False, True):
            from novic_b200 import _abi
            st = model._state(torch.device("cuda:0"))
            _abi.check(_abi.lib().novic_set_use_graphs(st["handle"], int(graphs)))
            with torch.inference_mode():
                o = orc.generate_greedy(cfg, sd, embed, 1.0, 0.0)
                g = model.generate(embed.cuda(), True, True, 1.0, 0.0, None, None, False)
            torch.cuda.synchronize()
            tok, pad, lg, ls, lb, sc = [None if t is None else t.cpu() for t in g]
            same = tok.shape == o["target"].shape and bool((tok == o["target"]).all())
            agree = (tok[:, : o["target"].shape[1]] == o["target"][:, : tok.shape[1]]).float().mean().item() if tok.numel() else 0
            print(f"greedy eos={eos} graphs={graphs}: T={tok.shape[1]} vs {o['target'].shape[1]} tokens identical={same} agree={agree:.4f} "
                  f"pad identical={bool((pad == o['padding']).all()) if pad.shape == o['padding'].shape else 'shape'} "
                  f"score max|d|={(sc - o['score']).abs().max().item():.4f} loss {ls.item():.3f} vs {o['loss_sum'].item():.3f} basis {int(lb)} vs {int(o['loss_basis'])}", flush=True)
            if lg is not None and lg.shape == o["logits"].shape:
                print(f"   step-logits max|d|={(lg - o['logits'])[~o['padding']].abs().max().item():.4f}")
            ok &= agree > 0.9
    return ok


def stage_beam():
    from novic_b200 import synth
    ok = True
    for eos in (False, True):
        dims, sd, cfg, orc, model = _oracle_setup(6)
        if eos:
            sd = synth.make_eos_friendly(sd, dims, beta=0.8)
            model.load_state_dict(sd)
        embed = synth.synth_embeddings(16, seed=1234)
        for (H, tau, alpha) in ((3, 1.0, 0.0), (5, 1.3, 0.6), (10, 1.0, 0.0)):
            with torch.inference_mode():
                o = orc.generate_beam(cfg, sd, embed, H, tau, alpha)
                g = model.generate_beam(embed.cuda(), H, tau, alpha, None, False, 0.0, None, False)
            torch.cuda.synchronize()
            tok, pad, sc = [t.cpu() for t in g]
            shape_ok = tok.shape == o["target"].shape
            agree = (tok == o["target"]).all(dim=2).float().mean().item() if shape_ok else 0.0
            print(f"beam eos={eos} H={H} tau={tau} alpha={alpha}: T={tok.shape[2]} vs {o['target'].shape[2]} beams identical frac={agree:.3f} "
                  f"score max|d|={(sc - o['score']).abs().max().item():.4f} pad same={bool((pad == o['padding']).all()) if shape_ok else 'shape'}", flush=True)
            ok &= agree > 0.8
    return ok


def stage_noise():
    from novic_b200 import synth, noise
    from oracle import novic_oracle as orc
    dev = torch.device("cuda:0")
    e0 = synth.synth_embeddings(64, seed=9)
    torch.manual_seed(3)
    na, nb = torch.randn(64, 1024), torch.randn(64, 1024)
    ua, ub, ga = torch.rand(64), torch.rand(64), torch.randn(64)
    ok = True
    cases = [
        ("GaussElem", noise.GaussElemNoise(1024, 3.25), (na, None, None, None), orc.noise_gauss_elem(e0, na, 3.25)),
        ("GaussVec", noise.GaussVecNoise(1024, 0.8), (na, None, ga, None), orc.noise_gauss_vec(e0, na, ga, 0.8)),
        ("GaussAngle", noise.GaussAngleNoise(1024, 30.0, 40.0), (na, None, ga, None), orc.noise_angle(e0, na, orc.gauss_angle(ga, 30.0, 40.0))),
        ("UniformAngle", noise.UniformAngleNoise(1024, 45.0, 75.0), (na, None, ua, None), orc.noise_angle(e0, na, orc.uniform_angle(ua, 45.0, 75.0))),
        ("Mix", noise.GaussElemUniformAngleNoise(1024, 3.25, 45.0, 75.0, 0.5), (na, nb, ua, ub),
         orc.noise_gauss_elem_uniform_angle(e0, na, ua, nb, ub, 3.25, 45.0, 75.0, 0.5)),
    ]
    for name, mod, pre, want in cases:
        got = mod.apply_predrawn(e0.clone().to(dev), *[None if t is None else t.to(dev) for t in pre]).cpu()
        err = (got - want).abs().max().item()
        print(f"noise {name}: predrawn max|d|={err:.2e}")
        ok &= err < 1e-5
        e = synth.synth_embeddings(4096, seed=10).to(dev)
        out = mod(e.clone())
        cosang = (out * e).sum(dim=1).clamp(-1, 1)
        print(f"   random: norm dev {(out.norm(dim=1) - 1).abs().max().item():.2e} mean cos {cosang.mean().item():.4f} angle deg [min {torch.rad2deg(torch.acos(cosang)).min().item():.2f}, "
              f"mean {torch.rad2deg(torch.acos(cosang)).mean().item():.2f}, max {torch.rad2deg(torch.acos(cosang)).max().item():.2f}]", flush=True)
    return ok


def stage_timing():
    from novic_b200 import synth, default_decoder, _abi
    dims = synth.DecoderDims()
    model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to("cuda:0")
    embed = synth.synth_embeddings(4096, seed=1234).cuda()
    with torch.inference_mode():
        for graphs in (True, False):
            st = model._state(torch.device("cuda:0"))
            _abi.check(_abi.lib().novic_set_use_graphs(st["handle"], int(graphs)))
            for _ in range(2):
                model.generate(embed, False, True, 1.0, 0.0, None, None, False)
            torch.cuda.synchronize()
            n0 = _abi.lib().novic_launch_count()
            t0 = time.perf_counter()
            iters = 5
            for _ in range(iters):
                model.generate(embed, False, True, 1.0, 0.0, None, None, False)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / iters
            print(f"greedy B=4096 graphs={graphs}: {dt * 1e3:.2f} ms -> {4096 / dt:,.0f} labels/s; launches/iter={(_abi.lib().novic_launch_count() - n0) / iters:.0f}", flush=True)
        for _ in range(2):
            model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"beam H=3 B=4096: {dt * 1e3:.2f} ms -> {4096 / dt:,.0f} labels/s")
    return True


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        ok = globals()["stage_" + name]()
        print(f"STAGE {name}: {'OK' if ok else 'FAILED'}", flush=True)
        sys.exit(0 if ok else 1)
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    results = {}
    for name in STAGES:
        t0 = time.time()
        try:
            r = subprocess.run(["timeout", "300", sys.executable, os.path.abspath(__file__), name], cwd=ROOT)
            results[name] = r.returncode
        except Exception as exc:  # noqa: BLE001
            results[name] = repr(exc)
        print(f"--- {name}: rc={results[name]} ({time.time() - t0:.1f}s)", flush=True)
    print("SUMMARY", results)


if __name__ == "__main__":
    main()
