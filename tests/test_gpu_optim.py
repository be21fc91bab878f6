"""Optimizer tail of the training step (train.py:1281-1286, :1108-1119): novic_b200.optim.FusedAdamW
(global-norm clip + AdamW over flat buffers, three library kernels, no host synchronisation) against torch's own
clip_grad_norm_ + torch.optim.AdamW(fused=True) on the same gradients, and the graph-replayed forward + backward against the direct
launches."""
import pytest
import torch

from novic_b200 import default_decoder, synth
from novic_b200.dist import train_step
from novic_b200.optim import FusedAdamW
from tests.golden_util import weight_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DIMS = synth.DecoderDims()


def _torch_twin(model, lr, betas, wd):
    """Separate parameter tensors with the model's values and the reference's two parameter groups (train.py:1108-1115)."""
    twins = [p.detach().clone().requires_grad_(True) for p in model._weight_tensors()]
    one_d = [p for p in twins if p.dim() < 2]
    n_d = [p for p in twins if p.dim() >= 2]
    opt = torch.optim.AdamW([{'params': one_d, 'weight_decay': 0.0}, {'params': n_d, 'weight_decay': wd}], lr=lr, betas=betas, fused=True)
    return twins, opt


@pytest.mark.parametrize("clip", [1.0, 0.0])
def test_fused_adamw_matches_torch_adamw_over_eight_steps(clip):
    model = default_decoder(DIMS, weight_case("lively")).to(DEV).train()
    lr, betas, wd = 1.5e-3, (0.9, 0.95), 0.1
    twins, topt = _torch_twin(model, lr, betas, wd)
    opt = FusedAdamW(model, lr=lr, betas=betas, weight_decay=wd, max_grad_norm=clip)
    params = model._weight_tensors()
    assert all(torch.equal(p, t) for p, t in zip(params, twins))               # re-seating the parameters kept their values
    total = sum(p.numel() for p in params)
    g = torch.Generator(device=DEV).manual_seed(5)
    for step in range(8):
        flat = torch.randn(total + 8, generator=g, device=DEV) * (0.02 if step % 2 else 2e-4)     # clipped and unclipped steps
        off = 0
        for t in twins:
            t.grad = flat[off:off + t.numel()].view(t.shape).clone()
            off += t.numel()
        if clip > 0:
            norm = torch.nn.utils.clip_grad_norm_(twins, max_norm=clip, error_if_nonfinite=True)
        topt.step()
        out = opt.step(flat_grads=flat)
        if clip > 0:
            assert abs(out[0].item() - norm.item()) <= 1e-5 * norm.item()
            assert abs(out[1].item() - min(1.0, clip / (norm.item() + 1e-6))) <= 1e-6
        else:
            assert out[1].item() == 1.0
    worst = max((p - t).abs().max().item() for p, t in zip(params, twins))
    assert worst <= 1e-6, worst
    for p, t in zip(params, twins):
        st, ts = opt.state[p], topt.state[t]
        assert (st["exp_avg"] - ts["exp_avg"]).abs().max().item() <= 1e-7
        assert (st["exp_avg_sq"] - ts["exp_avg_sq"]).abs().max().item() <= 1e-9
        assert float(st["step"]) == float(ts["step"]) == 8.0
    # 1-D tensors (LayerNorm gains) received no weight decay: with zero gradients they would not move at all
    sd = opt.state_dict()
    assert len(sd["param_groups"]) == 2 and sd["param_groups"][0]["weight_decay"] == 0.0


def test_fused_adamw_basis_normalisation_nonfinite_guard_and_state_dict():
    model = default_decoder(DIMS, weight_case("lively")).to(DEV).train()
    opt = FusedAdamW(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, max_grad_norm=1.0)
    params = model._weight_tensors()
    total = sum(p.numel() for p in params)
    g = torch.Generator(device=DEV).manual_seed(9)
    flat = torch.randn(total + 8, generator=g, device=DEV) * 0.5
    stats = torch.tensor([123.0, 250.0, 7.0], device=DEV)
    before = opt.flat_params.clone()
    out = opt.step(flat_grads=flat, stats=stats)
    norm = (flat[:total].double().norm() / 250.0).item()
    assert abs(out[0].item() - norm) <= 1e-5 * norm                              # the gradients are those of loss_sum: divided by loss_basis first
    assert abs(out[2].item() - min(1.0, 1.0 / (norm + 1e-6)) / 250.0) <= 1e-9
    assert not torch.equal(before, opt.flat_params)
    # non-finite gradients: flagged on the device, parameters and moments untouched, the host is told when it asks
    snap, m_snap = opt.flat_params.clone(), opt.flat_exp_avg.clone()
    bad = flat.clone()
    bad[12345] = float("inf")
    opt.step(flat_grads=bad)
    assert torch.equal(snap, opt.flat_params) and torch.equal(m_snap, opt.flat_exp_avg)
    with pytest.raises(RuntimeError):
        opt.grad_norm()
    # state dict round trip into a second optimizer
    sd = opt.state_dict()
    model2 = default_decoder(DIMS, weight_case("lively")).to(DEV).train()
    opt2 = FusedAdamW(model2, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, max_grad_norm=1.0)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2.flat_exp_avg, opt.flat_exp_avg) and torch.equal(opt2.flat_exp_avg_sq, opt.flat_exp_avg_sq) and opt2._step == opt._step
    assert opt2.state[model2._weight_tensors()[3]]["exp_avg"].data_ptr() == opt2.flat_exp_avg[sum(p.numel() for p in model2._weight_tensors()[:3]):].data_ptr()


def test_fused_train_step_learns_like_the_reference_style_loop():
    """dist.train_step with FusedAdamW (single rank) against the reference-style loop (autograd backward, clip_grad_norm_, torch AdamW) from
    the same weights on the same batch, dropout off: the loss trajectories agree (fp32 atomics in the weight gradients make them
    not bit-identical)."""
    embed = synth.synth_embeddings(64, seed=3).to(DEV)
    tgt, pad = synth.synth_targets(64, DIMS, seed=4)
    tgt, pad = tgt.to(DEV), pad.to(DEV)

    def run(fused):
        model = default_decoder(DIMS, weight_case("lively"), input_dropout=0.0, layer_dropout=0.0).to(DEV).train()
        if fused:
            opt = FusedAdamW(model, lr=1.5e-3, betas=(0.9, 0.95), weight_decay=0.1)
        else:
            decay = [p for p in model.parameters() if p.dim() >= 2]
            no_decay = [p for p in model.parameters() if p.dim() < 2]
            opt = torch.optim.AdamW([{'params': no_decay, 'weight_decay': 0.0}, {'params': decay, 'weight_decay': 0.1}], lr=1.5e-3, betas=(0.9, 0.95), fused=True)
        losses, norms = [], []
        for _ in range(8):
            loss, ncorrect, ntok, norm = train_step(model, opt, embed.clone(), tgt, pad, None, noise=None, gradient_clip=1.0)
            losses.append(loss.item())
            norms.append(float(norm))
        model.eval()
        with torch.inference_mode():
            out = model(embed, tgt, pad, None, True, True, False, None)
        return losses, norms, (out[2] / out[3]).item(), int(ntok.item())
    fl, fn, f_eval, f_tok = run(True)
    rl, rn, r_eval, r_tok = run(False)
    assert f_tok == r_tok == int((~pad).sum().item())
    assert fl[-1] < fl[0] - 0.3
    assert max(abs(a - b) for a, b in zip(fl, rl)) <= 0.02 * max(rl), (fl, rl)
    assert max(abs(a - b) / b for a, b in zip(fn, rn)) <= 0.05, (fn, rn)
    assert abs(f_eval - r_eval) <= 0.02 * r_eval          # the inference forward sees the weights the fused step left behind


@pytest.mark.parametrize("p_in,p_layer", [(0.0, 0.0), (0.1, 0.1)])
def test_graph_replay_equals_direct_launches(p_in, p_layer, monkeypatch):
    """The training step replayed from its CUDA graphs (novic_train_fwd_bwd_ex; the inputs staged in the workspace, the dropout seed in
    a device word) gives the gradients of the directly launched step: same seeds -> same masks; fp32 atomic order is the only difference.
    A second replay with another seed must change the masks (the graph does not bake the seed)."""
    from novic_b200 import training
    embed = synth.synth_embeddings(48, seed=3).to(DEV)
    tgt, pad = synth.synth_targets(48, DIMS, seed=4)
    tgt, pad = tgt.to(DEV), pad.to(DEV).view(torch.uint8)

    def grads_of(env, seeds):
        if env is not None:
            monkeypatch.setenv("NOVIC_TRAIN_GRAPHS", env)
        else:
            monkeypatch.delenv("NOVIC_TRAIN_GRAPHS", raising=False)
        model = default_decoder(DIMS, weight_case("eos"), input_dropout=p_in, layer_dropout=p_layer).to(DEV).train()
        outs = []
        for seed in seeds:
            torch.manual_seed(seed)
            loss, correct, pad_out, bucket = training.fwd_bwd(model, embed, tgt, pad, None, 1)
            outs.append((loss.clone(), bucket.flat[:bucket.total].clone()))
        return outs
    direct = grads_of("0", (11, 11, 12))
    replay = grads_of(None, (11, 11, 12))           # first call captures, the next two replay
    for (dl, dg), (rl, rg) in zip(direct, replay):
        assert abs(dl[0].item() - rl[0].item()) <= 1e-4 * abs(dl[0].item()) and dl[1].item() == rl[1].item()
        assert ((dg - rg).norm() / dg.norm()).item() <= 1e-4
    same = ((replay[0][1] - replay[1][1]).norm() / replay[0][1].norm()).item()
    other = ((replay[0][1] - replay[2][1]).norm() / replay[0][1].norm()).item()
    assert same <= 1e-4
    if p_in > 0:
        assert other > 0.05                          # a new seed -> new masks -> clearly different gradients
