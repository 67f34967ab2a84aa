"""Front ends of the from-scratch transformer acoustic model and the length arithmetic of convolutions
(mirrors ``allophant/network/frontend.py``).

The classes here are PARAMETER HOLDERS with the reference's module structure, construction order and ``state_dict`` keys
(``_layer.<i>.weight``, ``_layers.layers.<i>.module._weights.weight`` …); their arithmetic runs in
``network/transformer.py:TransformerPlan`` on the CUDA kernels.  Calling ``forward`` on them raises.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import padding
from ..config import DirectFrontendConfig, DropoutConfig, Glu1dConfig, LayerNormConfig, LinearFrontendConfig, MaxPoolingConfig, SequentialFrontendConfig


def conv_length(kernel_size: int, stride: int = 1, use_padding: bool = True, stft_type: bool = False) -> Callable[[Tensor], Tensor]:
    """Output length of a 1-D convolution (``frontend.py:192-203``)."""
    layer_padding = sum(padding.get_padding(kernel_size, stride, stft_type)) if use_padding else 0

    def padded_length(lengths: Tensor) -> Tensor:
        return torch.div(((lengths + layer_padding) - kernel_size), stride, rounding_mode="floor") + 1

    return padded_length


class _EngineOnly(nn.Module):
    def forward(self, *args: Any, **kwargs: Any) -> Any:
        raise RuntimeError(f"{type(self).__name__} is evaluated by TransformerAcousticModel (CUDA engine)")


class VariableLengthReflectPad(_EngineOnly):
    """``padding.py:24-53``: the index buffers are kept because they are part of the reference's ``state_dict``."""

    def __init__(self, pad: Tuple[int, int]):
        super().__init__()
        self._padding = pad
        left, right = pad
        base = torch.arange(0, right).view(1, 1, -1)
        self.register_buffer("_right_pad_start_indices", base + left)
        self.register_buffer("_right_pad_end_indices", base + 2)
        self.register_buffer("_left_pad_indices", torch.arange(left, 0, -1))

    @property
    def padding(self) -> Tuple[int, int]:
        return self._padding


class Glu1d(_EngineOnly):
    """``frontend.py:98-136``: (reflect) pad -> Conv1d(C -> 2 O, kernel, stride) -> GLU over channels."""

    def __init__(self, input_dimensions: int, output_dimensions: int, kernel_size: int, stride: int = 1, reflect_pad: bool = True):
        super().__init__()
        self._padding = padding.get_padding(kernel_size, stride)
        self._reflect_padding = VariableLengthReflectPad(self._padding) if reflect_pad else None
        self._kernel_size = kernel_size
        self._stride = stride
        self._weights = nn.Conv1d(input_dimensions, output_dimensions * 2, kernel_size=kernel_size, stride=stride)

    @property
    def padding(self) -> Tuple[int, int]:
        return self._padding

    @property
    def kernel_size(self) -> int:
        return self._kernel_size

    @property
    def stride(self) -> int:
        return self._stride


class LengthWrapper(_EngineOnly):
    """``frontend.py:49-79``: a layer plus the function that maps input lengths to output lengths."""

    def __init__(self, module: nn.Module, length_function: Optional[Callable[[Tensor], Tensor]] = None):
        super().__init__()
        self._length_function = length_function
        self.module = module

    def lengths(self, lengths: Tensor) -> Tensor:
        return lengths if self._length_function is None else self._length_function(lengths)


class LengthSequential(_EngineOnly):
    def __init__(self, *args: LengthWrapper):
        super().__init__()
        self.layers = nn.ModuleList(args)

    def lengths(self, lengths: Tensor) -> Tensor:
        for layer in self.layers:
            lengths = layer.lengths(lengths)
        return lengths


class Frontend(_EngineOnly):
    _output_dimensions: int

    @property
    def output_dimensions(self) -> int:
        return self._output_dimensions

    def lengths(self, input_lengths: Tensor) -> Tensor:
        return input_lengths


class DirectFrontend(Frontend):
    def __init__(self, config: DirectFrontendConfig, feature_size: int):
        super().__init__()
        self._output_dimensions = feature_size
        self._dropout = nn.Dropout(config.input_dropout) if config.input_dropout > 0 else None


class LinearFrontend(Frontend):
    """``frontend.py:167-189``: [Dropout,] LayerNorm(features) -> Linear -> LeakyReLU, as ``_layer`` (nn.Sequential)."""

    def __init__(self, config: LinearFrontendConfig, feature_size: int, elementwise_affine: bool = False):
        super().__init__()
        self._output_dimensions = config.neurons
        linear = nn.Linear(feature_size, config.neurons)
        modules: List[nn.Module] = [nn.LayerNorm(feature_size, elementwise_affine=elementwise_affine), linear, nn.LeakyReLU()]
        if config.input_dropout > 0:
            modules.insert(0, nn.Dropout(config.input_dropout))
        self._layer = nn.Sequential(*modules)

    @property
    def layer_norm(self) -> nn.LayerNorm:
        return next(m for m in self._layer if isinstance(m, nn.LayerNorm))

    @property
    def linear(self) -> nn.Linear:
        return next(m for m in self._layer if isinstance(m, nn.Linear))


class Transpose(nn.Module):
    def __init__(self, dimension_a: int, dimension_b: int):
        super().__init__()
        self._dimension_a = dimension_a
        self._dimension_b = dimension_b

    def forward(self, inputs: Tensor) -> Tensor:
        return inputs.transpose(self._dimension_a, self._dimension_b)


class SequentialFrontend(Frontend):
    """``frontend.py:219-276``."""

    def __init__(self, layers: LengthSequential, output_dimensions: int, upscale_factor: float = 1):
        super().__init__()
        self._layers = layers
        self._output_dimensions = output_dimensions
        self._upscale_factor = upscale_factor

    @classmethod
    def from_config(cls, config: SequentialFrontendConfig, feature_size: int) -> "SequentialFrontend":
        layers = []
        previous_output_size = feature_size
        upscale_factor = 1
        for layer in config.layers:
            if isinstance(layer, DropoutConfig):
                layers.append(LengthWrapper(nn.Dropout(layer.rate)))
            elif isinstance(layer, Glu1dConfig):
                module = Glu1d(previous_output_size, layer.out_channels, layer.kernel, layer.stride)
                layers.append(LengthWrapper(module, conv_length(module.kernel_size, module.stride)))
                previous_output_size = layer.out_channels
                upscale_factor *= module.stride
            elif isinstance(layer, LayerNormConfig):
                layers.append(
                    LengthWrapper(nn.Sequential(Transpose(-1, -2), nn.LayerNorm(previous_output_size, elementwise_affine=layer.affine), Transpose(-2, -1)))
                )
            elif isinstance(layer, MaxPoolingConfig):
                # The reference pools with stride = size but declares the lengths of a "same"-padded stride-1 pool
                # (frontend.py:258-259: conv_length(size) -> lengths + 1 for size 2), so behind this layer its frame counts exceed
                # the frames that exist and its own transformer fails on the key-padding mask ("Expected key_padded_mask.shape[1]
                # to be 15, but got 31"; tests/golden/max_pool_reference_behaviour.json holds the unmodified reference's exception).
                # There is no behaviour to reproduce: the configuration is rejected up front with the reason.
                raise NotImplementedError(
                    "max_pool layers cannot be run: the reference declares frame counts for them that exceed the pooled frames "
                    "(allophant/network/frontend.py:258-259) and its own forward pass fails on the attention mask"
                )
            else:
                raise ValueError(f"Unsupported layer config of type: {layer.__class__.__name__}")
        return cls(LengthSequential(*layers), previous_output_size, upscale_factor)

    @property
    def upscale_factor(self) -> float:
        return self._upscale_factor

    def downsampled_lengths(self, lengths: Tensor) -> Tensor:
        return self._layers.lengths(lengths)


def frontend_from_config(frontend_config: Any, feature_size: int, elementwise_affine: bool = False) -> Frontend:
    if isinstance(frontend_config, DirectFrontendConfig):
        return DirectFrontend(frontend_config, feature_size)
    if isinstance(frontend_config, LinearFrontendConfig):
        return LinearFrontend(frontend_config, feature_size, elementwise_affine)
    raise ValueError(f"Unsupported frontend config type {frontend_config.__class__.__name__}")
