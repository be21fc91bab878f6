"""Training step: gradients of the CUDA forward+backward against torch autograd through the
CPU oracle (dropout disabled on both sides; bf16 operands vs fp32, so per-tensor relative error is bounded, not bit-exact)."""
import numpy as np
import pytest
import torch

from novic_b200 import default_decoder, synth
from oracle import novic_oracle as orc
from tests.golden_util import weight_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_TOL = 0.05      # ||g_cuda - g_ref|| / ||g_ref|| per parameter tensor
COS_TOL = 0.998


def _oracle_grads(sd, embed, tgt, pad, weight, M, drop=None):
    cfg = orc.cfg_from_state_dict(sd)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "causality_mask"}
    A = tgt.shape[0]
    _, loss_sum, loss_basis, correct = orc.forward_loss(cfg, leaf, embed, tgt, pad, weight, drop=drop)
    loss_sum.backward()
    return loss_sum.item(), float(loss_basis), correct, {k: v.grad for k, v in leaf.items()}


@pytest.mark.parametrize("case", ["plain", "padded", "multi_weighted", "ragged"])
def test_gradients_match_autograd_oracle(case):
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    B = 5 if case == "ragged" else 24     # ragged: 95 rows - not a multiple of 8, 32 or 64 (edge paths of the row-blocked kernels and transposes)
    embed = synth.synth_embeddings(B, seed=21)
    if case == "multi_weighted":
        tgt, pad = synth.synth_targets(B, dims, seed=6, multi=3)
        w = torch.from_numpy(np.random.default_rng(8).random((B, 3)).astype(np.float32))
        w[1, 2] = 0.0
        M = 3
    else:
        tgt, pad = synth.synth_targets(B, dims, seed=5)
        w, M = None, 1
        if case == "plain":
            pad = None
    flat_t = tgt.view(B * M, -1)
    flat_p = None if pad is None else pad.view(B * M, -1)
    flat_w = None if w is None else w.view(-1)
    ref_loss, ref_basis, ref_correct, ref = _oracle_grads(sd, embed, flat_t, flat_p, flat_w, M)

    model = default_decoder(dims, sd, input_dropout=0.0, layer_dropout=0.0)
    if M > 1:
        from novic_b200.factory import synthetic_data_config
        model.data_config = synthetic_data_config(multi_target=True, use_weights=True)
    model = model.to(DEV).train()
    out = model(embed.to(DEV), tgt.to(DEV), None if pad is None else pad.to(DEV), None if w is None else w.to(DEV), True, True, False, None)
    logits, out_pad, loss_sum, loss_basis, correct = out
    assert logits is None
    assert abs(loss_sum.item() - ref_loss) <= 0.12 * ref_basis + 1e-3
    assert abs(float(loss_basis) - ref_basis) <= 1e-3 * max(1.0, ref_basis)
    (loss_sum / loss_basis).backward()
    scale = 1.0 / ref_basis
    got = dict(model.named_parameters())
    worst = {}
    for k, g_ref in ref.items():
        g = got[k].grad
        assert g is not None, k
        g = g.detach().cpu().double()
        r = (g_ref * scale).double()
        rel = (g - r).norm().item() / max(r.norm().item(), 1e-12)
        cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
        worst[k] = (rel, cos)
    bad = {k: v for k, v in worst.items() if v[0] > REL_TOL or v[1] < COS_TOL}
    assert not bad, bad


def test_reference_style_optimizer_step_runs():
    """loss.backward() -> clip_grad_norm_ -> AdamW.step() exactly as train.py:1273-1286 does, on our module's parameters."""
    dims = synth.DecoderDims()
    model = default_decoder(dims, weight_case("lively"), input_dropout=0.0, layer_dropout=0.0).to(DEV).train()
    decay = [p for p in model.parameters() if p.dim() >= 2]
    no_decay = [p for p in model.parameters() if p.dim() < 2]
    opt = torch.optim.AdamW([{'params': no_decay, 'weight_decay': 0.0}, {'params': decay, 'weight_decay': 0.1}], lr=1.5e-3, betas=(0.9, 0.95), fused=True)
    embed = synth.synth_embeddings(64, seed=3).to(DEV)
    tgt, pad = synth.synth_targets(64, dims, seed=4)
    tgt, pad = tgt.to(DEV), pad.to(DEV)
    losses = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        _, _, ls, lb, cor = model(embed, tgt, pad, None, True, True, False, None)
        (ls / lb).backward()
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0, error_if_nonfinite=True)
        opt.step()
        losses.append((ls / lb).item())
    assert losses[-1] < losses[0] - 0.3, losses      # the model learns the fixed batch
    model.eval()
    with torch.inference_mode():
        out = model(embed, tgt, pad, None, True, True, False, None)
    assert abs((out[2] / out[3]).item() - losses[-1]) < 1.0   # inference forward sees the updated weights (re-packed automatically)


@pytest.mark.parametrize("p_in,p_layer", [(0.1, 0.1), (0.0, 0.25), (0.3, 0.0)])
def test_dropout_matches_oracle_with_replayed_masks(p_in, p_layer):
    """Training with dropout (the reference's default: input_dropout = layer_dropout = 0.1, train.yaml; sites listed in SURVEY.md
    section 8 "numerical contract").  torch's dropout stream cannot be reproduced, so the masks are the product's own hash-generated
    ones, replayed on the CPU by the oracle (oracle.DropMasks): same masks -> loss and every gradient must agree as without dropout."""
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    B = 24
    embed = synth.synth_embeddings(B, seed=21)
    tgt, pad = synth.synth_targets(B, dims, seed=5)
    model = default_decoder(dims, sd, input_dropout=p_in, layer_dropout=p_layer).to(DEV).train()
    torch.manual_seed(123)
    _, _, loss_sum, loss_basis, _ = model(embed.to(DEV), tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
    (loss_sum / loss_basis).backward()
    got_p_in, got_p_layer, seed = model._last_dropout
    assert (got_p_in, got_p_layer) == (p_in, p_layer) and seed != 0
    drop = orc.DropMasks(p_in, p_layer, seed)
    ref_loss, ref_basis, _, ref = _oracle_grads(sd, embed, tgt, pad, None, 1, drop=drop)
    assert abs(loss_sum.item() - ref_loss) <= 0.12 * ref_basis + 1e-3
    got = dict(model.named_parameters())
    bad = {}
    for k, g_ref in ref.items():
        g = got[k].grad.detach().cpu().double()
        r = (g_ref / ref_basis).double()
        rel = (g - r).norm().item() / max(r.norm().item(), 1e-12)
        cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
        if rel > REL_TOL or cos < COS_TOL:
            bad[k] = (rel, cos)
    assert not bad, bad
    # and the masks matter: against the same batch WITHOUT dropout the gradients are clearly off (the test discriminates)
    _, plain_basis, _, plain = _oracle_grads(sd, embed, tgt, pad, None, 1)
    off = 0
    for k, g_ref in plain.items():
        g = got[k].grad.detach().cpu().double()
        r = (g_ref / plain_basis).double()
        off += (g - r).norm().item() / max(r.norm().item(), 1e-12) > 4 * REL_TOL
    assert off >= len(plain) // 2


def test_dropout_controls():
    """Fresh masks per call, reproducible under torch.manual_seed, off in eval mode, and utils.rescale_dropout-style mutation of the
    nn.Dropout holders (utils.py:177-192) is honoured."""
    dims = synth.DecoderDims()
    sd = weight_case("lively")
    embed = synth.synth_embeddings(16, seed=3).to(DEV)
    tgt, pad = synth.synth_targets(16, dims, seed=5)
    tgt, pad = tgt.to(DEV), pad.to(DEV)
    model = default_decoder(dims, sd).to(DEV).train()                      # reference defaults: 0.1 / 0.1

    def loss():
        out = model(embed, tgt, pad, None, True, True, False, None)
        return out[2].item()
    torch.manual_seed(7); a = loss()
    b = loss()
    torch.manual_seed(7); c = loss()
    assert a != b and a == c
    for m in model.modules():                                              # what utils.rescale_dropout(model, 0) does
        if isinstance(m, torch.nn.modules.dropout._DropoutNd):
            m.p *= 0.0
    d, e = loss(), loss()
    assert d == e and model._last_dropout[:2] == (0.0, 0.0)
    model.eval()
    with torch.inference_mode():
        f = model(embed, tgt, pad, None, True, True, False, None)[2].item()
    assert abs(f - d) <= 1e-3 * abs(d)


@pytest.mark.parametrize("R,Cc,ld_src,ld_dst,off", [(456, 512, 512, 456, 0), (95, 128, 128, 96, 0), (57, 72, 80, 64, 0), (130, 64, 192, 136, 64), (33, 8, 8, 40, 0)])
def test_transpose_kernel_on_ragged_and_strided_shapes(R, Cc, ld_src, ld_dst, off):
    """dst[c, r] = src[r, c]: aligned 16-byte path and the element-wise edges, bit-exact (it only moves bf16 values)."""
    from novic_b200 import _abi
    lib = _abi.lib()
    g = torch.Generator().manual_seed(R * 1000 + Cc)
    full = torch.randn(R, ld_src + off, generator=g).to(torch.bfloat16).to(DEV)
    src = full[:, off:]                                   # column offset: rows start 128 B into the allocation's rows
    dst = torch.full((Cc, ld_dst), 7.0, dtype=torch.bfloat16, device=DEV)
    _abi.check(lib.novic_debug_transpose_bf16(src.data_ptr(), R, Cc, ld_src + off, dst.data_ptr(), ld_dst, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :R].cpu(), src[:, :Cc].t().cpu())
    assert bool((dst[:, R:] == 7.0).all())                # nothing written past the valid rows


def test_gradients_with_transposed_operand_copies(monkeypatch):
    """NOVIC_WGRAD_MN=0 - weight- and data-gradient GEMMs on transposed bf16 copies (K-major descriptors), the path before the MN-major
    descriptors - gives the same gradients as the default path up to fp32 atomic order."""
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    embed = synth.synth_embeddings(24, seed=21).to(DEV)
    tgt, pad = synth.synth_targets(24, dims, seed=5)
    grads = []
    for flag in ("1", "0"):
        monkeypatch.setenv("NOVIC_WGRAD_MN", flag)       # read when a handle is created
        model = default_decoder(dims, sd, input_dropout=0.0, layer_dropout=0.0).to(DEV).train()
        _, _, loss_sum, loss_basis, _ = model(embed, tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
        (loss_sum / loss_basis).backward()
        grads.append({k: p.grad.detach().double().cpu() for k, p in model.named_parameters()})
        del model
    monkeypatch.setenv("NOVIC_WGRAD_MN", "1")
    default_decoder(dims, sd).to(DEV)._state(torch.device(DEV))      # restore the process-wide switch
    for k, g in grads[0].items():
        rel = (grads[1][k] - g).norm().item() / max(g.norm().item(), 1e-30)
        assert rel <= 1e-5, (k, rel)


@pytest.mark.parametrize("Mo,No,K", [(128, 512, 19456), (512, 128, 456), (1536, 512, 95), (200, 72, 1000), (6907, 512, 360), (128, 128, 64), (512, 1024, 7)])
def test_wgrad_gemm_from_row_major_operands(Mo, No, K):
    """dw += a^T b straight from the row-major activations (MN-major tcgen05 descriptors; what the training step runs): against an fp64
    product of the same bf16 operands, on top of a non-zero dw.  The columns between Mo / No and the leading dimension hold NaN: they
    may be loaded (a 64-column block is the unit) but must never reach a stored element; rows beyond K do not exist (TMA zero fill)."""
    from novic_b200 import _abi
    lib = _abi.lib()
    lda, ldb = (Mo + 63) // 64 * 64, (No + 63) // 64 * 64 + 64
    g = torch.Generator().manual_seed(Mo + No + K)
    a = torch.full((K, lda), float("nan"), dtype=torch.bfloat16); a[:, :Mo] = (torch.randn(K, Mo, generator=g) * 0.5).to(torch.bfloat16)
    b = torch.full((K, ldb), float("nan"), dtype=torch.bfloat16); b[:, :No] = (torch.randn(K, No, generator=g) * 0.5).to(torch.bfloat16)
    dw0 = torch.randn(Mo, No, generator=g)
    dw = dw0.clone().to(DEV)
    ad, bd = a.to(DEV), b.to(DEV)
    _abi.check(lib.novic_debug_wgrad_mn(ad.data_ptr(), Mo, lda, bd.data_ptr(), No, ldb, K, dw.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = dw0.double() + a[:, :Mo].double().t() @ b[:, :No].double()
    err = (dw.cpu().double() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, K ** 0.5), err


@pytest.mark.parametrize("Mo,No,K", [(128, 512, 19456), (512, 128, 456), (1536, 512, 95), (200, 72, 1000), (6912, 512, 360)])
def test_wgrad_gemm_accumulates_into_fp32(Mo, No, K):
    """dw += a_t b_t^T with split-K vector reductions: against an fp64 product of the same bf16 operands, on top of a non-zero dw."""
    from novic_b200 import _abi
    lib = _abi.lib()
    ld = (K + 63) // 64 * 64
    g = torch.Generator().manual_seed(Mo + No + K)
    a = torch.zeros(Mo, ld, dtype=torch.bfloat16); a[:, :K] = (torch.randn(Mo, K, generator=g) * 0.5).to(torch.bfloat16)
    b = torch.zeros(No, ld, dtype=torch.bfloat16); b[:, :K] = (torch.randn(No, K, generator=g) * 0.5).to(torch.bfloat16)
    dw0 = torch.randn(Mo, No, generator=g)
    dw = dw0.clone().to(DEV)
    ad, bd = a.to(DEV), b.to(DEV)
    _abi.check(lib.novic_debug_wgrad(ad.data_ptr(), Mo, bd.data_ptr(), No, K, ld, dw.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = dw0.double() + a[:, :K].double() @ b[:, :K].double().t()
    err = (dw.cpu().double() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, K ** 0.5), err          # fp32 accumulation of K products of O(0.25) values
