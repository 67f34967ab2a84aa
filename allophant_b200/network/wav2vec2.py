"""Parameter container of the wav2vec2 / XLS-R encoder with the checkpoint's key layout.

The reference wraps Hugging Face's ``Wav2Vec2Model`` (``acoustic_model.py:796-798``) and its
checkpoints therefore carry that module's parameter names.  This module
re-creates exactly those names — nothing else of the Hugging Face implementation — so that
``load_state_dict`` accepts an Allophant ``model_state`` unchanged.  The arithmetic lives in
``allophant_b200.engine`` (CUDA); these modules hold fp32 master parameters only and
raise if called.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch
from torch import nn
from torch.nn import Parameter


@dataclass(frozen=True)
class Wav2Vec2EncoderConfig:
    """The subset of ``transformers.Wav2Vec2Config`` the forward path reads.  Defaults: XLS-R-300M."""

    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    intermediate_size: int = 4096
    layer_norm_eps: float = 1e-5
    feat_extract_norm: str = "layer"  # "layer" (XLS-R, large) | "group" (wav2vec2-base)
    conv_dim: Tuple[int, ...] = (512, 512, 512, 512, 512, 512, 512)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_bias: bool = True
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    do_stable_layer_norm: bool = True
    apply_spec_augment: bool = True
    mask_time_prob: float = 0.075
    mask_time_length: int = 10
    mask_time_min_masks: int = 2
    mask_feature_prob: float = 0.0
    mask_feature_length: int = 10
    mask_feature_min_masks: int = 0
    hidden_dropout: float = 0.1
    attention_dropout: float = 0.1
    feat_proj_dropout: float = 0.1
    activation_dropout: float = 0.0
    layerdrop: float = 0.1
    # feature extractor (preprocessor_config.json)
    sampling_rate: int = 16000
    feature_size: int = 1
    do_normalize: bool = True
    return_attention_mask: bool = True

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


# Hub unreachable offline (the reference calls it unconditionally, acoustic_model.py:787,798):
# the constants of the models the reference was published with are carried here.
KNOWN_MODELS: Dict[str, Wav2Vec2EncoderConfig] = {
    "facebook/wav2vec2-xls-r-300m": Wav2Vec2EncoderConfig(),
    "facebook/wav2vec2-large-xlsr-53": Wav2Vec2EncoderConfig(mask_time_prob=0.075),
    "facebook/wav2vec2-xls-r-1b": Wav2Vec2EncoderConfig(hidden_size=1280, num_hidden_layers=48, intermediate_size=5120),
    # post-LN encoder, GroupNorm feature extractor without conv biases, no attention mask (preprocessor: return_attention_mask
    # false)
    "facebook/wav2vec2-large": Wav2Vec2EncoderConfig(
        feat_extract_norm="group", conv_bias=False, do_stable_layer_norm=False, mask_time_prob=0.05, return_attention_mask=False,
    ),
    # the same ordering at width 768: 12 layers, 12 heads of 64, 3072 feed-forward units, positional conv in 16 groups of 48
    # channels (run as four block-diagonal super groups of 192, engine.PackedEncoder.pos_span); inference only
    "facebook/wav2vec2-base": Wav2Vec2EncoderConfig(
        hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, feat_extract_norm="group", conv_bias=False,
        do_stable_layer_norm=False, mask_time_prob=0.05, return_attention_mask=False,
    ),
}


def encoder_config_for(model_id: str) -> Wav2Vec2EncoderConfig:
    if model_id in KNOWN_MODELS:
        return KNOWN_MODELS[model_id]
    raise ValueError(
        f"Unknown wav2vec2 model id {model_id!r}: allophant_b200 carries the configuration of "
        f"{sorted(KNOWN_MODELS)} (the Hugging Face hub is not consulted)"
    )


class _Params(nn.Module):
    """A module that only owns parameters; its arithmetic is done by the CUDA engine."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} holds parameters only; the forward pass runs in allophant_b200.engine (CUDA)"
        )


class _Affine(_Params):
    """weight/bias pair with LayerNorm/GroupNorm initialisation."""

    def __init__(self, size: int) -> None:
        super().__init__()
        self.weight = Parameter(torch.ones(size))
        self.bias = Parameter(torch.zeros(size))


class _Linear(_Params):
    def __init__(self, in_features: int, out_features: int, std: float = 0.02) -> None:
        super().__init__()
        self.weight = Parameter(torch.empty(out_features, in_features).normal_(0.0, std))
        self.bias = Parameter(torch.zeros(out_features))


class _Conv(_Params):
    def __init__(self, in_channels: int, out_channels: int, kernel: int, bias: bool) -> None:
        super().__init__()
        weight = torch.empty(out_channels, in_channels, kernel)
        nn.init.kaiming_normal_(weight)
        self.weight = Parameter(weight)
        if bias:
            bound = math.sqrt(1.0 / (in_channels * kernel))
            self.bias = Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)


class _ConvLayer(_Params):
    def __init__(self, in_channels: int, out_channels: int, kernel: int, bias: bool, norm: bool) -> None:
        super().__init__()
        self.conv = _Conv(in_channels, out_channels, kernel, bias)
        if norm:
            self.layer_norm = _Affine(out_channels)


class _FeatureExtractor(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        layers = []
        for index, (channels, kernel) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
            in_channels = 1 if index == 0 else cfg.conv_dim[index - 1]
            has_norm = cfg.feat_extract_norm == "layer" or index == 0
            layers.append(_ConvLayer(in_channels, channels, kernel, cfg.conv_bias, has_norm))
        self.conv_layers = nn.ModuleList(layers)


class _FeatureProjection(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.layer_norm = _Affine(cfg.conv_dim[-1])
        self.projection = _Linear(cfg.conv_dim[-1], cfg.hidden_size)


class _WeightNormParams(_Params):
    """``parametrizations.weight.original0`` (g, [1,1,k]) / ``original1`` (v, [O,Cg,k]) of weight_norm(dim=2)."""

    def __init__(self, out_channels: int, group_channels: int, kernel: int) -> None:
        super().__init__()
        std = 2 * math.sqrt(1.0 / (kernel * out_channels))
        v = torch.empty(out_channels, group_channels, kernel).normal_(0.0, std)
        self.original0 = Parameter(v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())
        self.original1 = Parameter(v)


class _Parametrizations(_Params):
    def __init__(self, weight: _WeightNormParams) -> None:
        super().__init__()
        self.weight = weight


class _PosConvInner(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        group_channels = cfg.hidden_size // cfg.num_conv_pos_embedding_groups
        self.bias = Parameter(torch.zeros(cfg.hidden_size))
        self.parametrizations = _Parametrizations(
            _WeightNormParams(cfg.hidden_size, group_channels, cfg.num_conv_pos_embeddings)
        )


class _PosConvEmbed(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.conv = _PosConvInner(cfg)


class _Attention(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.k_proj = _Linear(cfg.hidden_size, cfg.hidden_size)
        self.v_proj = _Linear(cfg.hidden_size, cfg.hidden_size)
        self.q_proj = _Linear(cfg.hidden_size, cfg.hidden_size)
        self.out_proj = _Linear(cfg.hidden_size, cfg.hidden_size)


class _FeedForward(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.intermediate_dense = _Linear(cfg.hidden_size, cfg.intermediate_size)
        self.output_dense = _Linear(cfg.intermediate_size, cfg.hidden_size)


class _EncoderLayer(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.attention = _Attention(cfg)
        self.layer_norm = _Affine(cfg.hidden_size)
        self.feed_forward = _FeedForward(cfg)
        self.final_layer_norm = _Affine(cfg.hidden_size)


class _Encoder(_Params):
    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.pos_conv_embed = _PosConvEmbed(cfg)
        self.layer_norm = _Affine(cfg.hidden_size)
        self.layers = nn.ModuleList([_EncoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class Wav2Vec2Weights(_Params):
    """Parameters of ``Wav2Vec2Model`` under their Hugging Face names (422 tensors for XLS-R-300M)."""

    def __init__(self, cfg: Wav2Vec2EncoderConfig) -> None:
        super().__init__()
        self.config = cfg
        if cfg.mask_time_prob > 0.0 or cfg.mask_feature_prob > 0.0:
            self.masked_spec_embed = Parameter(torch.empty(cfg.hidden_size).uniform_())
        self.feature_extractor = _FeatureExtractor(cfg)
        self.feature_projection = _FeatureProjection(cfg)
        self.encoder = _Encoder(cfg)

    def freeze_feature_encoder(self) -> None:
        for parameter in self.feature_extractor.parameters():
            parameter.requires_grad = False
