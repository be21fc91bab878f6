"""Drop-in `PrefixedIterDecoder` for the reference's decoder seam, backed by libnovic_b200.so.

The reference selects its decoder with `getattr(embedding_decoder, cfg.model)` (infer.py:716) and then only
uses the `EmbeddingDecoder` API (embedding_decoder.py:20-201).  This module mirrors that API - constructor
keyword arguments (embedding_decoder.py:43-75, :633-640), `forward` / `generate` / `generate_beam` signatures
and return tuples, `get_num_params`, and the state-dict key names and shapes - so
that `embedding_decoder.PrefixedIterDecoder = novic_b200.PrefixedIterDecoder` is the whole integration
(see INTEGRATION.md).  The body is not the reference's: parameters are plain holders, and every compute method
calls the CUDA library through ctypes.  There is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import fractions
import math
from typing import Any, Optional

import torch
import torch.nn as nn

from . import _abi, guide

MAX_SEQS_PER_CALL = 1 << 15  # sequences (embeddings x beams) decoded per library call; larger batches are chunked


# ----------------------------------------------------------------------------------------------------------
# Parameter holders: they exist so that state_dict() has the reference's key names; none has a forward().
# ----------------------------------------------------------------------------------------------------------
class _PrefixProjection(nn.Module):            # reference: EmbeddingVectorMLP, embedding_decoder.py:1161-1276
    def __init__(self, embed_dim: int, out_features: int):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(embed_dim, out_features, bias=False))


class _PositionTable(nn.Module):               # reference: LearnedPosEmbedding, embedding_decoder.py:1279-1297
    def __init__(self, max_seq_len: int, dim: int, dropout_prob: float):
        super().__init__()
        self.embedding = nn.Embedding(max_seq_len, dim)
        self.dropout = nn.Dropout(p=dropout_prob)


class _SelfAttention(nn.Module):               # key names of nn.MultiheadAttention(bias=False)
    def __init__(self, dim: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.out_proj = nn.Linear(dim, dim, bias=False)


class _Layer(nn.Module):                       # key names of nn.TransformerEncoderLayer(bias=False)
    def __init__(self, dim: int, ffn_dim: int, dropout_prob: float = 0.0):
        super().__init__()
        # holder of the layer's dropout probability (attention probabilities, both residual branches, feed-forward activation all use
        # layer_dropout in nn.TransformerEncoderLayer); an nn.Dropout so that utils.rescale_dropout (utils.py:177-192) reaches it
        self.dropout = nn.Dropout(p=dropout_prob)
        self.self_attn = _SelfAttention(dim)
        self.linear1 = nn.Linear(dim, ffn_dim, bias=False)
        self.linear2 = nn.Linear(ffn_dim, dim, bias=False)
        self.norm1 = nn.LayerNorm(dim, bias=False)
        self.norm2 = nn.LayerNorm(dim, bias=False)


class _Stack(nn.Module):                       # key names of nn.TransformerEncoder(norm=LayerNorm)
    def __init__(self, dim: int, ffn_dim: int, num_layers: int, dropout_prob: float = 0.0):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(dim, ffn_dim, dropout_prob) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(dim, bias=False)


@dataclasses.dataclass(frozen=True)
class ParamCount:
    """Same fields and `to_str()` as the reference's ParamCount (embedding_decoder.py:1304-1347)."""
    total: int
    used: int
    unused: int
    trained: int
    frozen: int

    def to_str(self) -> str:
        s = f"{self.used} params"
        if self.unused:
            s += f" + {self.unused} unused"
        if self.frozen:
            s += f" where used is {self.trained} trained + {self.frozen} frozen"
        return s

    @staticmethod
    def of(params, unused: int = 0) -> "ParamCount":
        trained = sum(p.numel() for p in params if p.requires_grad)
        frozen = sum(p.numel() for p in params if not p.requires_grad)
        used = trained + frozen - unused
        return ParamCount(total=used + unused, used=used, unused=unused, trained=trained - unused, frozen=frozen)


class EmbeddingDecoder(nn.Module):
    """Interface of the reference's abstract decoder (embedding_decoder.py:20-201)."""

    @classmethod
    def get_target_config_kwargs(cls, **target_kwargs) -> dict[str, Any]:
        raise NotImplementedError

    @classmethod
    def get_data_config_kwargs(cls, **data_kwargs) -> dict[str, Any]:
        raise NotImplementedError

    def get_num_params(self):
        raise NotImplementedError

    def forward(self, embed, target, target_padding, target_weight, calc_loss, calc_correct, only_pred, guide_targets):
        raise NotImplementedError

    def generate(self, embed, collect_logits, calc_loss, temperature, length_alpha, sample_weight, guide_targets, guide_renorm):
        raise NotImplementedError

    def generate_beam(self, embed, topk, temperature, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm):
        raise NotImplementedError

    def precompute_generate_all(self, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm):
        raise NotImplementedError

    def generate_all(self, embed, topk, temperature, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm, precompute=None):
        raise NotImplementedError


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class PrefixedIterDecoder(EmbeddingDecoder):
    """B200-native replacement of embedding_decoder.PrefixedIterDecoder (embedding_decoder.py:617-1079)."""

    @classmethod
    def get_target_config_kwargs(cls, **target_kwargs) -> dict[str, Any]:
        # end token = pad token = 0, no start token, compact ids (embedding_decoder.py:619-627)
        target_kwargs.update(with_start_token=False, with_end_token=True, compact_ids=True)
        return target_kwargs

    @classmethod
    def get_data_config_kwargs(cls, **data_kwargs) -> dict[str, Any]:
        return data_kwargs

    def __init__(self, *, embedder, data_config, mlp_seq_len: int, weight_tying: bool, strictly_causal: bool,
                 enable_nested: bool, vocab_quant: bool, num_end_loss: int, label_smoothing: float, hidden_dim: int,
                 feedfwd_scale: Any, mlp_hidden_layer: str, mlp_hidden_bias: bool, mlp_hidden_norm: bool,
                 mlp_hidden_activation: str, input_dropout: float, num_layers: int, num_heads: int, layer_dropout: float,
                 layer_activation: str, layer_norm_first: bool, layer_bias: bool, logits_bias: bool, init_bias_zero: bool,
                 init_mlp_mode: str, init_mlp_unit_norm: bool, init_tfrm_mode: str, init_tfrm_unit_norm: bool,
                 init_tfrm_unit_postnorm: bool, init_tfrm_proj_layers: bool, init_zero_norm: bool, init_rezero_mode: str):
        super().__init__()
        self.embedder = embedder
        self.target_config = embedder.target_config
        self.target_vocab = embedder.target_vocab
        self.data_config = data_config
        self.embed_dtype = embedder.embed_dtype
        self.embed_dim = embedder.embed_dim
        self.vocab_quant = vocab_quant
        self.num_end_loss = num_end_loss
        self.label_smoothing = label_smoothing
        self.hidden_dim = hidden_dim
        self.feedfwd_scale = fractions.Fraction(feedfwd_scale)
        ffn = self.hidden_dim * self.feedfwd_scale
        if ffn.denominator != 1:
            raise ValueError(f"Feedforward dimension scaler ({self.feedfwd_scale}) must result in an integral feedforward dimension when applied to hidden dimension ({self.hidden_dim})")
        self.feedfwd_dim = ffn.numerator
        self.mlp_seq_len = mlp_seq_len
        self.weight_tying = weight_tying
        self.strictly_causal = strictly_causal
        self.enable_nested = enable_nested
        self.input_dropout = input_dropout
        self.layer_dropout = layer_dropout
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.init_cfg = dict(init_bias_zero=init_bias_zero, init_mlp_mode=init_mlp_mode, init_mlp_unit_norm=init_mlp_unit_norm,
                             init_tfrm_mode=init_tfrm_mode, init_tfrm_unit_norm=init_tfrm_unit_norm,
                             init_tfrm_unit_postnorm=init_tfrm_unit_postnorm, init_tfrm_proj_layers=init_tfrm_proj_layers,
                             init_zero_norm=init_zero_norm)

        # The CUDA kernels implement the reference's default architecture (config/train.yaml:224-308); anything
        # else is refused loudly rather than silently computed differently.
        unsupported = []
        if num_end_loss < 1: unsupported.append("num_end_loss < 1")
        if mlp_seq_len < 1: unsupported.append("mlp_seq_len < 1")
        if hidden_dim != 512: unsupported.append(f"hidden_dim={hidden_dim} (kernels: 512)")
        if self.feedfwd_dim != 128: unsupported.append(f"feedforward dim={self.feedfwd_dim} (kernels: 128)")
        if num_heads != 8: unsupported.append(f"num_heads={num_heads} (kernels: 8)")
        if mlp_hidden_layer != 'none': unsupported.append(f"mlp_hidden_layer={mlp_hidden_layer!r}")
        if not weight_tying: unsupported.append("weight_tying=False")
        if layer_activation != 'gelu': unsupported.append(f"layer_activation={layer_activation!r}")
        if not layer_norm_first: unsupported.append("layer_norm_first=False")
        if layer_bias or logits_bias: unsupported.append("biases")
        if init_rezero_mode != 'none': unsupported.append(f"init_rezero_mode={init_rezero_mode!r}")
        if self.embed_dtype != torch.float32: unsupported.append(f"embed_dtype={self.embed_dtype}")
        if self.embed_dim % 128 != 0: unsupported.append(f"embed_dim={self.embed_dim} (must be a multiple of 128)")
        if num_layers > _abi.NOVIC_MAX_LAYERS: unsupported.append(f"num_layers={num_layers}")
        if unsupported:
            raise ValueError("PrefixedIterDecoder (novic_b200) does not support: " + ", ".join(unsupported))

        E, P, V = hidden_dim, mlp_seq_len, self.target_config.vocab_size
        self.max_seq_len = P + self.target_config.token_length - 1
        self.vocab_size_quant = math.ceil(V / 64) * 64 if vocab_quant else V
        self.embed_mlp = _PrefixProjection(self.embed_dim, P * E)
        self.logits_linear = nn.Linear(E, self.vocab_size_quant, bias=False)
        self.token_embedding = None
        self.pos_embedding = _PositionTable(self.max_seq_len, E, input_dropout)
        self.transformer = _Stack(E, self.feedfwd_dim, num_layers, layer_dropout)
        mask = torch.triu(torch.full((self.max_seq_len, self.max_seq_len), float('-inf'), dtype=self.embed_dtype), diagonal=1)
        if not strictly_causal:
            mask[:P, :P] = 0
        self.register_buffer('causality_mask', mask)
        self.reset_parameters()
        # a checkpoint's vocab_quant rows (V .. ceil64(V)) must be zero, as the reference's loader demands (embedding_decoder.py:437-441)
        self.register_load_state_dict_post_hook(self._verify_unused)

        self._handles: dict[int, dict] = {}  # per CUDA device: library handle, packed weights, workspace

    # ------------------------------------------------------------------------------------------------------
    # Initialisation: same distributions as the reference's default ('balanced') init - SURVEY.md 8c(2)
    # ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def reset_parameters(self) -> None:
        ic = self.init_cfg
        E, L, K, P = self.hidden_dim, self.num_layers, self.feedfwd_dim, self.mlp_seq_len
        if ic['init_mlp_mode'] not in ('default', 'balanced'):
            raise ValueError(f"Unrecognised value for MLP initialisation mode: {ic['init_mlp_mode']}")
        if ic['init_tfrm_mode'] not in ('default', 'open', 'balanced'):
            raise ValueError(f"Unrecognised value for transformer initialisation mode: {ic['init_tfrm_mode']}")
        unit = ic['init_mlp_unit_norm']
        embed_std = 1 / math.sqrt(2 * E) if unit else 1 / math.sqrt(2)
        if ic['init_mlp_mode'] == 'balanced':
            nn.init.normal_(self.embed_mlp.mlp[0].weight, std=embed_std)  # balanced scale 1/sqrt(2), no output bias
        nn.init.normal_(self.logits_linear.weight, std=embed_std)
        if self.vocab_size_quant > self.target_config.vocab_size:
            self.logits_linear.weight[self.target_config.vocab_size:].zero_()
        nn.init.normal_(self.pos_embedding.embedding.weight, std=embed_std)
        f = 1 / math.sqrt(E)
        nominal = f if ic['init_tfrm_unit_norm'] else 1.0
        per_layer = 1 / math.sqrt(2 * L) if ic['init_tfrm_proj_layers'] else 1.0
        if ic['init_tfrm_mode'] != 'default':
            if ic['init_tfrm_mode'] == 'open':
                s_in, s_out, s_f1, s_f2 = f, f, f / math.sqrt(2), f
            else:
                gain = 0.6521 if not (ic['init_tfrm_unit_norm'] or ic['init_zero_norm']) else 0.5
                attn_scale = math.sqrt((1 + nominal ** 4 * (P - 1) / P) / P)
                s_in, s_out, s_f1, s_f2 = f, f / attn_scale, f, 1 / (math.sqrt(K) * gain)
            s_out *= per_layer
            s_f2 *= per_layer
            for layer in self.transformer.layers:
                nn.init.normal_(layer.self_attn.in_proj_weight, std=s_in)
                nn.init.normal_(layer.self_attn.out_proj.weight, std=s_out)
                nn.init.normal_(layer.linear1.weight, std=s_f1)
                nn.init.normal_(layer.linear2.weight, std=s_f2)
        else:
            for layer in self.transformer.layers:
                nn.init.xavier_uniform_(layer.self_attn.in_proj_weight)
        norm_scale = 0.0 if ic['init_zero_norm'] else nominal
        for layer in self.transformer.layers:
            nn.init.constant_(layer.norm1.weight, norm_scale)
            nn.init.constant_(layer.norm2.weight, norm_scale)
        nn.init.constant_(self.transformer.norm.weight, f if ic['init_tfrm_unit_postnorm'] else 1.0)

    def _verify_unused(self, *args, **kwargs):
        V = self.target_config.vocab_size
        if self.vocab_size_quant > V and bool(torch.any(self.logits_linear.weight[V:] != 0)):
            raise ValueError("Unexpected values in the unused portion of a parameter tensor")

    def get_num_params(self):
        unused = (self.vocab_size_quant - self.target_config.vocab_size) * self.hidden_dim
        groups = {
            'Input MLP': ParamCount.of(list(self.embed_mlp.parameters())),
            'Token embed/logits': ParamCount.of(list(self.logits_linear.parameters()), unused=unused),
            'Positional embed': ParamCount.of(list(self.pos_embedding.parameters())),
            'Transformer': ParamCount.of(list(self.transformer.parameters())),
        }
        total = ParamCount.of(list(self.parameters()), unused=unused)
        return total, groups

    # ------------------------------------------------------------------------------------------------------
    # Library plumbing
    # ------------------------------------------------------------------------------------------------------
    def _weight_tensors(self) -> list[torch.Tensor]:
        ts = [self.embed_mlp.mlp[0].weight, self.logits_linear.weight, self.pos_embedding.embedding.weight, self.transformer.norm.weight]
        for layer in self.transformer.layers:
            ts += [layer.self_attn.in_proj_weight, layer.self_attn.out_proj.weight, layer.linear1.weight, layer.linear2.weight,
                   layer.norm1.weight, layer.norm2.weight]
        return ts

    def _state(self, device: torch.device, refresh: bool = False) -> dict:
        if device.type != 'cuda':
            raise RuntimeError("novic_b200.PrefixedIterDecoder computes on CUDA (sm_100a) only; there is no CPU path. "
                               f"Got tensors on {device}.")
        lib = _abi.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        st = self._handles.get(idx)
        with torch.cuda.device(idx):
            if st is None:
                cfg = _abi.NovicCfg(
                    embed_dim=self.embed_dim, hidden_dim=self.hidden_dim, ffn_dim=self.feedfwd_dim, num_layers=self.num_layers,
                    num_heads=self.num_heads, prefix_len=self.mlp_seq_len, vocab_size=self.target_config.vocab_size,
                    token_length=self.target_config.token_length, strictly_causal=int(self.strictly_causal),
                    num_end_loss=self.num_end_loss, ln_eps=1e-5, label_smoothing=float(self.label_smoothing))
                handle = C.c_void_p()
                _abi.check(lib.novic_create(C.byref(cfg), C.byref(handle)))
                st = dict(handle=handle, wbuf=None, wkey=None, ws=None)
                self._handles[idx] = st
            tensors = self._weight_tensors()
            wkey = tuple((t.data_ptr(), 0 if t.is_inference() else t._version) for t in tensors)
            if st['wkey'] != wkey or refresh:
                for t in tensors:
                    if t.device.type != 'cuda' or t.device.index != idx or t.dtype != torch.float32 or not t.is_contiguous():
                        raise RuntimeError("decoder parameters must be contiguous fp32 tensors on the same CUDA device as the input "
                                           f"(found {t.dtype} on {t.device}); call .to(device) first")
                if st['wbuf'] is None:
                    st['wbuf'] = torch.empty(lib.novic_weight_bytes(st['handle']), dtype=torch.uint8, device=device)
                w = _abi.NovicWeights()
                w.embed_mlp, w.tok_embed, w.pos_embed, w.final_norm = (t.data_ptr() for t in tensors[:4])
                for i, layer in enumerate(self.transformer.layers):
                    w.in_proj[i] = layer.self_attn.in_proj_weight.data_ptr()
                    w.out_proj[i] = layer.self_attn.out_proj.weight.data_ptr()
                    w.linear1[i] = layer.linear1.weight.data_ptr()
                    w.linear2[i] = layer.linear2.weight.data_ptr()
                    w.norm1[i] = layer.norm1.weight.data_ptr()
                    w.norm2[i] = layer.norm2.weight.data_ptr()
                stream = torch.cuda.current_stream(idx).cuda_stream
                _abi.check(lib.novic_set_weights(st['handle'], C.byref(w), st['wbuf'].data_ptr(), st['wbuf'].numel(), stream))
                st['wkey'] = wkey
        return st

    def train(self, mode: bool = True):
        # Fused optimizers update parameters without bumping tensor version counters, so the (data_ptr, version) key
        # cannot see their updates: re-pack the bf16 operand copies on every training forward and whenever the mode flips.
        for st in getattr(self, '_handles', {}).values():
            st['wkey'] = None
        return super().train(mode)

    def refresh_weights(self) -> None:
        """Force the next call to re-convert the fp32 parameters (needed only if they were modified in place by an
        operation that does not bump tensor versions while the module stayed in eval mode)."""
        for st in self._handles.values():
            st['wkey'] = None

    def _workspace(self, st: dict, device: torch.device, num_embeds: int, seqs_per_embed: int, rows_per_seq: int) -> torch.Tensor:
        need = _abi.lib().novic_workspace_bytes(st['handle'], num_embeds, seqs_per_embed, rows_per_seq)
        ws = st['ws']
        if ws is None or ws.numel() < need:
            st['ws'] = None
            del ws
            st['ws'] = ws = torch.empty(need, dtype=torch.uint8, device=device)
        return ws

    def _check_embed(self, embed: torch.Tensor) -> torch.Tensor:
        assert embed.ndim == 2 and embed.dtype == self.embed_dtype  # embedding_decoder.py:661
        if embed.shape[1] != self.embed_dim:
            raise ValueError(f"embedding dimension {embed.shape[1]} != {self.embed_dim}")
        return embed.contiguous()

    def __del__(self):
        try:
            lib = _abi.lib()
            for st in getattr(self, '_handles', {}).values():
                lib.novic_destroy(st['handle'])
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    # forward (embedding_decoder.py:659-777)
    # ------------------------------------------------------------------------------------------------------
    def forward(self, embed, target, target_padding, target_weight, calc_loss, calc_correct, only_pred, guide_targets):
        embed = self._check_embed(embed)
        if self.training and torch.is_grad_enabled() and target is not None:
            if guide_targets is not None:
                raise NotImplementedError("guided correctness evaluation is available in evaluation mode only (the training loop never uses it)")
            return self._forward_train(embed, target, target_padding, target_weight, calc_loss, calc_correct, only_pred)
        if guide_targets is not None and calc_correct:
            assert not only_pred  # embedding_decoder.py:755
        if target is None:
            # embedding-only forward (embedding_decoder.py:690-693 with no token positions): logits of the first generated position.
            # A one-column target contributes no token embedding (target[:, :-1] is empty), so the same call computes it.
            assert target_padding is None and target_weight is None
            dummy = torch.zeros((embed.shape[0], 1), dtype=self.target_config.token_dtype, device=embed.device)
            logits, _, _, _, _ = self.forward(embed, dummy, None, None, False, False, only_pred, None)
            return logits, None, None, None, None
        B = embed.shape[0]
        multi = target.ndim == 3
        multi_first = bool(multi and getattr(self.data_config, 'multi_target', False) and getattr(self.data_config, 'multi_first', False))
        if multi:
            if multi_first:  # M x B x C -> B x M x C
                target = target.transpose(0, 1)
                target_padding = None if target_padding is None else target_padding.transpose(0, 1)
                target_weight = None if target_weight is None else target_weight.transpose(0, 1)
            M = target.shape[1]
            assert target.shape[0] == B
            target = target.reshape(B * M, target.shape[-1])
            target_padding = None if target_padding is None else target_padding.reshape(B * M, -1)
            target_weight = None if target_weight is None else target_weight.reshape(B * M)
        else:
            M = 1
        tc = self.target_config
        assert target.dtype == tc.token_dtype and target.ndim == 2 and target.shape[0] == B * M and target.shape[1] >= 1
        assert target_padding is None or (target_padding.dtype == tc.mask_dtype and target_padding.shape == target.shape)
        assert target_weight is None or (target_weight.dtype == self.embed_dtype and target_weight.ndim == 1 and target_weight.shape[0] == target.shape[0])
        if target.dtype != torch.int64:
            raise ValueError("token ids must be int64")
        A, Ct = target.shape
        T = 1 if only_pred else Ct
        V = tc.vocab_size
        dev = embed.device
        st = self._state(dev)
        ws = self._workspace(st, dev, B, M, self.max_seq_len)
        target_c = target.contiguous()
        pad_c = None if target_padding is None else target_padding.contiguous().view(torch.uint8)
        w_c = None if target_weight is None else target_weight.contiguous()
        has_pad = pad_c is not None or w_c is not None
        logits = torch.empty((A, T, V), dtype=torch.float32, device=dev)
        pad_out = torch.empty((A, T), dtype=torch.uint8, device=dev) if has_pad else None
        loss = torch.empty(2, dtype=torch.float32, device=dev) if calc_loss else None
        correct = torch.empty((A, T), dtype=torch.uint8, device=dev) if calc_correct else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            if guide_targets is not None and calc_correct:
                # guided correctness evaluation (embedding_decoder.py:754-760): the trie spans the first Ct columns of the guide targets
                # its own cache (one trie per trimmed target width Ct of an evaluation batch): the decode paths' tries are not evicted
                if not hasattr(self, "_trie_cache_fwd"):
                    self._trie_cache_fwd = guide.TrieCache(max_entries=max(8, tc.token_length))
                trie = self._trie_cache_fwd.get(guide_targets, Ct, V, dev)
                masks = torch.empty(A * Ct * ((V + 31) // 32), dtype=torch.int32, device=dev)
                _abi.check(_abi.lib().novic_forward_guided(st['handle'], embed.data_ptr(), B, M, target_c.data_ptr(), _ptr(pad_c), _ptr(w_c), Ct,
                                                           logits.data_ptr(), _ptr(pad_out), _ptr(loss), _ptr(correct), guide.guide_arg(trie, False),
                                                           masks.data_ptr(), masks.numel() * 4, ws.data_ptr(), ws.numel(), stream))
            else:
                _abi.check(_abi.lib().novic_forward(st['handle'], embed.data_ptr(), B, M, target_c.data_ptr(), _ptr(pad_c), _ptr(w_c), Ct,
                                                    int(bool(only_pred)), logits.data_ptr(), _ptr(pad_out), _ptr(loss), _ptr(correct),
                                                    ws.data_ptr(), ws.numel(), stream))
        out_pad = None if pad_out is None else pad_out.view(torch.bool)
        out_correct = None if correct is None else correct.view(torch.bool)
        loss_sum = loss_basis = None
        if calc_loss:
            loss_sum = loss[0]
            loss_basis = loss[1] if (target_weight is not None) else loss[1].round().to(torch.int64)
        if multi:
            shape = (B, M)
            logits = logits.view(*shape, T, V)
            out_pad = None if out_pad is None else out_pad.view(*shape, T)
            out_correct = None if out_correct is None else out_correct.view(*shape, T)
            if multi_first:
                logits = logits.transpose(0, 1)
                out_pad = None if out_pad is None else out_pad.transpose(0, 1)
                out_correct = None if out_correct is None else out_correct.transpose(0, 1)
        return logits, out_pad, loss_sum, loss_basis, out_correct

    def _forward_train(self, embed, target, target_padding, target_weight, calc_loss, calc_correct, only_pred):
        """Training-mode forward with autograd (train.py:1270-1273).  One fused forward+backward library call; the logits are
        not materialised (the reference's training loop discards them, train.py:1271), so the first return value is None.
        Dropout: input_dropout after the positional embedding (embedding_decoder.py:1297) and layer_dropout on the attention
        probabilities, both residual branches and the activated feed-forward rows (nn.TransformerEncoderLayer), with masks from a
        counter-based hash seeded from torch's CPU generator once per call (reproducible under torch.manual_seed; not torch's own
        dropout stream).  The probabilities are read from the nn.Dropout holders so that utils.rescale_dropout applies."""
        from . import training
        if only_pred:
            raise NotImplementedError("only_pred=True is not supported in training mode")
        B = embed.shape[0]
        multi = target.ndim == 3
        multi_first = bool(multi and getattr(self.data_config, 'multi_target', False) and getattr(self.data_config, 'multi_first', False))
        if multi:
            if multi_first:
                target = target.transpose(0, 1)
                target_padding = None if target_padding is None else target_padding.transpose(0, 1)
                target_weight = None if target_weight is None else target_weight.transpose(0, 1)
            M = target.shape[1]
            target = target.reshape(B * M, target.shape[-1])
            target_padding = None if target_padding is None else target_padding.reshape(B * M, -1)
            target_weight = None if target_weight is None else target_weight.reshape(B * M)
        else:
            M = 1
        tc = self.target_config
        assert target.dtype == tc.token_dtype and target.ndim == 2 and target.shape[0] == B * M and target.shape[1] >= 2
        assert target_padding is None or (target_padding.dtype == tc.mask_dtype and target_padding.shape == target.shape)
        assert target_weight is None or (target_weight.dtype == self.embed_dtype and target_weight.ndim == 1 and target_weight.shape[0] == target.shape[0])
        loss_sum, loss_basis, correct, pad_out = training.train_forward(self, embed, target, target_padding, target_weight, M)
        has_pad = target_padding is not None or target_weight is not None
        out_pad = pad_out.view(torch.bool) if has_pad else None
        out_correct = correct.view(torch.bool) if calc_correct else None
        if target_weight is None:
            loss_basis = loss_basis.detach().round().to(torch.int64)
        if multi:
            T = target.shape[1]
            out_pad = None if out_pad is None else out_pad.view(B, M, T)
            out_correct = None if out_correct is None else out_correct.view(B, M, T)
            if multi_first:
                out_pad = None if out_pad is None else out_pad.transpose(0, 1)
                out_correct = None if out_correct is None else out_correct.transpose(0, 1)
        return None, out_pad, (loss_sum if calc_loss else None), (loss_basis if calc_loss else None), out_correct

    # ------------------------------------------------------------------------------------------------------
    # generate (embedding_decoder.py:779-850)
    # ------------------------------------------------------------------------------------------------------
    def _guide_trie(self, guide_targets, device, check_content: bool = True):
        """W x Cmax guide targets -> cached device trie (novic_b200/guide.py), or None when unguided."""
        if guide_targets is None:
            return None
        assert guide_targets.ndim == 2 and guide_targets.dtype == self.target_config.token_dtype
        if not hasattr(self, "_trie_cache"):
            self._trie_cache = guide.TrieCache()
        return self._trie_cache.get(guide_targets, self.target_config.token_length - 1, self.target_config.vocab_size, device, check_content)

    def _vocab_prior(self, guide_targets, vocab_targets, per_token, scaler, device):
        """(trie, per-edge prior scores) of a beam search with a vocabulary prior (embedding_decoder.py:881-891, :924-936), cached
        per (guide tensor, vocabulary tensor, mode, scaler)."""
        assert vocab_targets.ndim == 2 and vocab_targets.dtype == self.target_config.token_dtype
        if not hasattr(self, "_prior_cache"):
            self._prior_cache = []
        def ident(t):   # identity + version, and the content where no version counter exists (inference tensors)
            if t is None:
                return None
            ver = guide.tensor_version(t)
            return (t.data_ptr(), tuple(t.shape), ver, str(t.device), guide.content_checksum(t) if ver < 0 else None)
        key = (ident(guide_targets), ident(vocab_targets), per_token, scaler, str(device))
        for k, _, v in self._prior_cache:
            if k == key:
                return v
        G, V = self.target_config.token_length - 1, self.target_config.vocab_size
        is_guide = guide_targets is not None and (vocab_targets is guide_targets or (
            vocab_targets.shape == guide_targets.shape and bool(torch.equal(vocab_targets, guide_targets))))     # :885
        gtrie = None if guide_targets is None else guide.build_trie(guide_targets, G, V)
        trie, bias = guide.prior_bias(gtrie, vocab_targets, is_guide, per_token, scaler, G, V)
        value = (trie.to(device), bias.to(device))
        self._prior_cache.append((key, (guide_targets, vocab_targets), value))   # the sources stay alive: their addresses cannot be reused
        if len(self._prior_cache) > 4:
            self._prior_cache.pop(0)
        return value

    def generate(self, embed, collect_logits, calc_loss, temperature, length_alpha, sample_weight, guide_targets, guide_renorm):
        if not temperature > 0:
            raise ValueError("temperature must be positive")
        embed = self._check_embed(embed)
        B = embed.shape[0]
        dev = embed.device
        G = self.target_config.token_length - 1
        V = self.target_config.vocab_size
        st = self._state(dev)
        tok = torch.empty((B, G), dtype=torch.int64, device=dev)
        pad = torch.empty((B, G), dtype=torch.uint8, device=dev)
        score = torch.empty(B, dtype=torch.float32, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        length = torch.empty(B, dtype=torch.float32, device=dev)
        logits = torch.empty((B, G, V), dtype=torch.float32, device=dev) if collect_logits else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib = _abi.lib()
        T = 0
        chunk = min(B, MAX_SEQS_PER_CALL)
        ws = self._workspace(st, dev, chunk, 1, 0)
        trie = self._guide_trie(guide_targets, dev)
        garg = guide.guide_arg(trie, bool(guide_renorm))
        with torch.cuda.device(dev):
            for b0 in range(0, B, chunk):
                n = min(chunk, B - b0)
                t_out = C.c_int32(0)
                _abi.check(lib.novic_generate_greedy(
                    st['handle'], embed[b0:b0 + n].data_ptr(), n, float(temperature), float(length_alpha), tok[b0:b0 + n].data_ptr(),
                    pad[b0:b0 + n].data_ptr(), score[b0:b0 + n].data_ptr(), nll[b0:b0 + n].data_ptr(), length[b0:b0 + n].data_ptr(),
                    None if logits is None else logits[b0:b0 + n].data_ptr(), C.byref(t_out), garg, ws.data_ptr(), ws.numel(), stream))
                T = max(T, t_out.value)
        target = tok[:, :T]
        target_padding = pad.view(torch.bool)[:, :T]
        seq_logits = None if logits is None else logits[:, :T, :]
        if not calc_loss:
            return target, target_padding, seq_logits, None, None, None
        # loss_sum / loss_basis (:838-846): one deterministic reduction kernel of the library, no eager torch arithmetic
        totals = torch.empty(2, dtype=torch.float32, device=dev)
        basis_i = torch.empty(1, dtype=torch.int64, device=dev) if sample_weight is None else None
        if sample_weight is not None:
            assert sample_weight.dtype == torch.float32 and sample_weight.shape == (B,)
            sample_weight = sample_weight.contiguous()
        with torch.cuda.device(dev):
            _abi.check(lib.novic_loss_totals(nll.data_ptr(), length.data_ptr(), _ptr(sample_weight), B, totals.data_ptr(), _ptr(basis_i), stream))
        loss_sum = totals[0]
        loss_basis = basis_i[0] if sample_weight is None else totals[1]
        return target, target_padding, seq_logits, loss_sum, loss_basis, score

    def generate_async(self, embed, temperature=1.0, length_alpha=0.0, guide_targets=None, guide_renorm=False):
        """generate() without its host synchronisation (serving loops, novic_b200/serve.py): everything is enqueued on the current
        stream.  Returns (target B x G int64, target_padding B x G bool, target_score B, T int32 device tensor of one element);
        the reference's outputs are target[:, :T], target_padding[:, :T] once T has been read back.  One library call, so B is
        limited to MAX_SEQS_PER_CALL."""
        if not temperature > 0:
            raise ValueError("temperature must be positive")
        embed = self._check_embed(embed)
        B = embed.shape[0]
        if B > MAX_SEQS_PER_CALL:
            raise ValueError(f"generate_async decodes at most {MAX_SEQS_PER_CALL} embeddings per call (got {B}); split the batch")
        dev = embed.device
        G = self.target_config.token_length - 1
        st = self._state(dev)
        tok = torch.empty((B, G), dtype=torch.int64, device=dev)
        pad = torch.empty((B, G), dtype=torch.uint8, device=dev)
        score = torch.empty(B, dtype=torch.float32, device=dev)
        T = torch.empty(1, dtype=torch.int32, device=dev)
        ws = self._workspace(st, dev, B, 1, 0)
        trie = self._guide_trie(guide_targets, dev, check_content=False)     # no host synchronisation on this path
        garg = guide.guide_arg(trie, bool(guide_renorm))
        with torch.cuda.device(dev):
            _abi.check(_abi.lib().novic_generate_greedy_async(
                st['handle'], embed.data_ptr(), B, float(temperature), float(length_alpha), tok.data_ptr(), pad.data_ptr(), score.data_ptr(),
                None, None, T.data_ptr(), garg, ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return tok, pad.view(torch.bool), score, T

    # ------------------------------------------------------------------------------------------------------
    # generate_beam (embedding_decoder.py:852-984)
    # ------------------------------------------------------------------------------------------------------
    def generate_beam(self, embed, topk, temperature, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm):
        vocab_on = vocab_targets is not None and vocab_scaler != 0                       # :881
        if vocab_on and vocab_scaler < 0:
            raise ValueError("vocab_scaler must be >= 0: a negative scaler gives +inf scores to ids outside the vocabulary (:934-936)")
        if not temperature > 0:
            raise ValueError("temperature must be positive")
        embed = self._check_embed(embed)
        B = embed.shape[0]
        H = int(topk)
        dev = embed.device
        G = self.target_config.token_length - 1
        if H == 1 and vocab_on:
            raise NotImplementedError("a beam of one with a vocabulary prior is not implemented in novic_b200 (use topk >= 2)")
        if self.num_end_loss != 1:
            # with N > 1 the reference's effective padding lags the end token by N - 1 positions (embedding_decoder.py:700-707), so a
            # finished candidate grows N - 1 more tokens whose log-probabilities enter its score before being zeroed (:913, :980)
            raise NotImplementedError("generate_beam with num_end_loss != 1 is not implemented in novic_b200")
        if H == 1:
            # a beam of one is the greedy path; scores coincide (sum of log-probs, length-normalised)
            t, p, _, _, _, s = self.generate(embed, False, True, temperature, length_alpha, None, guide_targets, guide_renorm)
            return t.unsqueeze(1), p.unsqueeze(1), s.unsqueeze(1)
        st = self._state(dev)
        tok = torch.empty((B, H, G), dtype=torch.int64, device=dev)
        pad = torch.empty((B, H, G), dtype=torch.uint8, device=dev)
        score = torch.empty((B, H), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib = _abi.lib()
        T = 0
        chunk = max(1, min(B, MAX_SEQS_PER_CALL // H))
        ws = self._workspace(st, dev, chunk, H, 0)
        if vocab_on:
            trie, bias = self._vocab_prior(guide_targets, vocab_targets, bool(vocab_per_token), float(vocab_scaler), dev)
            garg = guide.guide_arg(trie, bool(guide_renorm) and guide_targets is not None, bias)
        else:
            trie = self._guide_trie(guide_targets, dev)
            garg = guide.guide_arg(trie, bool(guide_renorm))
        with torch.cuda.device(dev):
            for b0 in range(0, B, chunk):
                n = min(chunk, B - b0)
                t_out = C.c_int32(0)
                _abi.check(lib.novic_generate_beam(
                    st['handle'], embed[b0:b0 + n].data_ptr(), n, H, float(temperature), float(length_alpha), tok[b0:b0 + n].data_ptr(),
                    pad[b0:b0 + n].data_ptr(), score[b0:b0 + n].data_ptr(), C.byref(t_out), garg, ws.data_ptr(), ws.numel(), stream))
                T = max(T, t_out.value)
        return tok[:, :, :T], pad.view(torch.bool)[:, :, :T], score

    # ------------------------------------------------------------------------------------------------------
    # generate_all (embedding_decoder.py:986-1079): score every guide target by teacher forcing, return the top-k
    # ------------------------------------------------------------------------------------------------------
    def precompute_generate_all(self, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm):
        """Embedding-independent part of generate_all (embedding_decoder.py:986-1050).  Returns the same 5-tuple shape as the
        reference, except that the third item is the guide trie (or None) where the reference holds its dense
        1 x W x C x V guide-score tensor - callers treat the tuple as opaque (infer.py:535-541)."""
        W, Cmax = guide_targets.shape
        paddings = guide.target_paddings(guide_targets)                                    # :991-994
        C = Cmax - int(paddings.all(dim=0).sum().item())                                   # :996
        paddings = paddings[:, :C].contiguous()
        targets = guide_targets[:, :C].masked_fill(paddings, 0).contiguous()               # :997-998
        trie = self._guide_trie(guide_targets, guide_targets.device) if guide_renorm else None
        if vocab_targets is None or vocab_scaler == 0:
            vocab_scores = None
        else:                                                                              # :1020-1044
            vs = guide.vocab_prior_scores(targets, paddings, vocab_targets[:, :C], bool(vocab_per_token), self.target_config.vocab_size)
            vocab_scores = (vs.to(device=guide_targets.device, dtype=self.embed_dtype) * vocab_scaler).unsqueeze(0)
        if length_alpha == 0:
            alpha_scale = None
        else:                                                                              # :1046-1050
            n = (C - paddings.sum(dim=1)).clamp(min=1).to(self.embed_dtype)
            alpha_scale = n.pow(-length_alpha).unsqueeze(0)
        return targets, paddings, trie, vocab_scores, alpha_scale

    def generate_all(self, embed, topk, temperature, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm, precompute=None):
        if not temperature > 0:
            raise ValueError("temperature must be positive")
        if precompute is None:
            precompute = self.precompute_generate_all(length_alpha=length_alpha, vocab_targets=vocab_targets, vocab_per_token=vocab_per_token,
                                                      vocab_scaler=vocab_scaler, guide_targets=guide_targets, guide_renorm=guide_renorm)
        targets, paddings, trie, vocab_scores, alpha_scale = precompute
        embed = self._check_embed(embed)
        B, K = embed.shape[0], int(topk)
        W, C = targets.shape
        dev = embed.device
        if targets.device != dev:
            targets, paddings = targets.to(dev), paddings.to(dev)
        if trie is not None and trie.child_off.device != dev:
            trie = trie.to(dev)
        st = self._state(dev)
        lib = _abi.lib()
        stream = torch.cuda.current_stream(dev).cuda_stream
        garg = guide.guide_arg(trie, True)
        scores = torch.empty((B, W), dtype=torch.float32, device=dev)
        bc = max(1, min(B, MAX_SEQS_PER_CALL // 16))               # embeddings per call
        mc = max(1, min(W, MAX_SEQS_PER_CALL // bc))                # guide targets per embedding per call
        ws = self._workspace(st, dev, bc, mc, self.max_seq_len)
        pad_u8 = paddings.view(torch.uint8)
        with torch.cuda.device(dev):
            for w0 in range(0, W, mc):
                m = min(mc, W - w0)
                for b0 in range(0, B, bc):
                    n = min(bc, B - b0)
                    tgt = targets[w0:w0 + m].unsqueeze(0).expand(n, -1, -1).reshape(n * m, C).contiguous()      # :1064
                    pad = pad_u8[w0:w0 + m].unsqueeze(0).expand(n, -1, -1).reshape(n * m, C).contiguous()
                    out = torch.empty(n * m, dtype=torch.float32, device=dev)
                    _abi.check(lib.novic_score_targets(st['handle'], embed[b0:b0 + n].data_ptr(), n, m, tgt.data_ptr(), pad.data_ptr(), C,
                                                       float(temperature), garg, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                    scores[b0:b0 + n, w0:w0 + m] = out.view(n, m)
        if vocab_scores is not None:
            scores.sub_(vocab_scores.to(dev))                                              # :1074-1075
        if alpha_scale is not None:
            scores.mul_(alpha_scale.to(dev))                                               # :1076-1077
        topk_scores, topk_indices = torch.topk(scores, k=K, dim=1, largest=True, sorted=True)   # :1079-1083
        idx3 = topk_indices.unsqueeze(2).expand(-1, -1, C)
        topk_targets = targets.expand(B, -1, -1).gather(dim=1, index=idx3)
        topk_paddings = paddings.expand(B, -1, -1).gather(dim=1, index=idx3)
        return topk_targets, topk_paddings, topk_scores

