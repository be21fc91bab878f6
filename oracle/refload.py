"""Import the *unmodified* reference (pallgeuer/novic): from /root/reference in the build container, else from the
byte-for-byte staged copy oracle/_ref/ (oracle/build_ref.py; a git-ignored build output that travels to the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: tests/, bench.py's CPU legs and __graft_entry__.smoke() are the only callers; nothing
under novic_b200/ imports it.  The only shims are a stub `unidecode` module (utils.py:19 imports it, only utils.get_canon uses
it) and a SimpleNamespace standing in for the CLIP embedder, of which the decoder reads four attributes
(embedding_decoder.py:77-86).
"""
from __future__ import annotations

import os
import sys
import types

import torch

STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py: byte-for-byte copies, git-ignored


def _default_root() -> str:
    if "NOVIC_REFERENCE_ROOT" in os.environ:
        return os.environ["NOVIC_REFERENCE_ROOT"]
    if os.path.isfile("/root/reference/embedding_decoder.py"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _default_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "embedding_decoder.py"))


def full_tree() -> bool:
    """True in the build container (the whole reference tree, e.g. embedding_cache.py); the staged copy holds the hot path only."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "embedding_cache.py"))


def import_reference():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "unidecode" not in sys.modules:
        stub = types.ModuleType("unidecode")
        stub.unidecode = lambda s: s
        sys.modules["unidecode"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import embedders, embedding_dataset, embedding_decoder, embedding_noise, infer, utils  # noqa: E401
    return types.SimpleNamespace(embedders=embedders, embedding_dataset=embedding_dataset,
                                 embedding_decoder=embedding_decoder, embedding_noise=embedding_noise,
                                 infer=infer, utils=utils)


# config/train.yaml:250-308 defaults, passed the way infer.py:721-758 builds them
DEFAULT_MODEL_CFG = dict(
    model="PrefixedIterDecoder", vocab_quant=False, num_end_loss=1, label_smoothing=0.0, hidden_dim=512,
    feedfwd_scale="1/4", mlp_seq_len=4, weight_tying=True, strictly_causal=False, enable_nested=False,
    mlp_hidden_layer="none", mlp_hidden_bias=False, mlp_hidden_norm=False, mlp_hidden_activation="gelu",
    input_dropout=0.1, num_layers=6, num_heads=8, layer_dropout=0.1, layer_activation="gelu",
    layer_norm_first=True, layer_bias=False, logits_bias=False, init_bias_zero=True, init_mlp_mode="balanced",
    init_mlp_unit_norm=False, init_tfrm_mode="balanced", init_tfrm_unit_norm=False, init_tfrm_unit_postnorm=True,
    init_tfrm_proj_layers=True, init_zero_norm=False, init_rezero_mode="none",
)


def fake_embedder(ref, vocab_size: int = 6912, token_length: int = 16, embed_dim: int = 1024):
    tc = ref.embedders.TargetConfig(
        vocab_size=vocab_size, token_dtype=torch.int64, mask_dtype=torch.bool, start_token_id=None, end_token_id=0,
        pad_token_id=0, compact_ids=True, compact_map=None, compact_unmap=None, fixed_token_length=False,
        token_length=token_length, use_masks=True)
    return types.SimpleNamespace(target_config=tc, target_vocab=("x",), embed_dtype=torch.float32, embed_dim=embed_dim)


def build_reference_decoder(ref, state_dict=None, *, vocab_size: int = 6912, token_length: int = 16,
                            embed_dim: int = 1024, multi_target: bool = False, use_weights: bool = False, **overrides):
    """Construct the reference's PrefixedIterDecoder through infer.load_decoder_model (infer.py:713-778)."""
    embedder = fake_embedder(ref, vocab_size, token_length, embed_dim)
    dc_dict = dict(use_weights=use_weights, multi_target=multi_target)
    if use_weights:
        dc_dict.update(unit_weights=False)
    if multi_target:
        dc_dict.update(multi_first=False, full_targets=False, fixed_multi_length=True, multi_length=3)
    dc = ref.embedding_dataset.DataConfig.create(dc_dict, use_targets=True)
    cfg = ref.utils.AttrDict({**DEFAULT_MODEL_CFG, **overrides}) if hasattr(ref.utils, "AttrDict") else types.SimpleNamespace(**{**DEFAULT_MODEL_CFG, **overrides})
    ckpt = None if state_dict is None else dict(model_state_dict=state_dict)
    model = ref.infer.load_decoder_model(cfg=cfg, embedder=embedder, data_config=dc, checkpoint=ckpt)
    return model.eval()
