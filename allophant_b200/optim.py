"""Optimiser step of the training loop (``allophant/estimator.py:778-791``, ``allophant/config.py:107-173, 213-335``).

``clip_grad_norm_`` / ``FusedAdam`` / ``WarmupScheduler`` / ``OptimizerWrapper`` keep the reference's (and torch's)
interfaces and state layouts — ``FusedAdam.state_dict()`` is interchangeable with ``torch.optim.Adam``'s — while the
arithmetic is three multi-tensor CUDA launches per ~48 parameter tensors (sum of squares, Adam with the clip
coefficient applied on the fly) instead of ~10 small kernels per tensor.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional

import torch
from torch import Tensor

from . import engine, ops
from ._lib import check, lib


def _pointer_table(rows: List[List[Optional[Tensor]]]):
    flat = [0 if t is None else t.data_ptr() for row in rows for t in row]
    return (ctypes.c_void_p * len(flat))(*flat)


def _sizes(tensors: List[Tensor]):
    return (ctypes.c_int64 * len(tensors))(*[t.numel() for t in tensors])


def _checked(tensors: Iterable[Tensor]) -> List[Tensor]:
    result = []
    for tensor in tensors:
        if not tensor.is_cuda:
            raise RuntimeError("allophant_b200.optim runs on CUDA tensors only (no CPU fallback exists)")
        if tensor.dtype != torch.float32 or not tensor.is_contiguous():
            raise ValueError("allophant_b200.optim expects contiguous fp32 parameters and gradients")
        result.append(tensor)
    return result


def sum_of_squares(tensors: List[Tensor]) -> Tensor:
    """fp64 device scalar: the sum of squares of all elements (one multi-tensor launch per 48 tensors)."""
    tensors = _checked(tensors)
    out = torch.empty(1, device=tensors[0].device, dtype=torch.float64)
    check(lib.aph_multi_tensor_sumsq(_pointer_table([[t] for t in tensors]), _sizes(tensors), len(tensors), out.data_ptr(), ops._stream()), "aph_multi_tensor_sumsq")
    return out


def clip_grad_norm_(parameters: Iterable[Tensor], max_norm: float) -> Tensor:
    """``nn.utils.clip_grad_norm_(parameters, max_norm)`` (norm 2): scales the gradients in place by
    ``min(1, max_norm / (total_norm + 1e-6))`` and returns the total norm — without a host synchronisation."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros((), dtype=torch.float32)
    grads = _checked(grads)
    sumsq = sum_of_squares(grads)
    check(lib.aph_multi_tensor_scale(_pointer_table([[g] for g in grads]), _sizes(grads), len(grads), sumsq.data_ptr(), float(max_norm), ops._stream()), "aph_multi_tensor_scale")
    return sumsq.sqrt().float().squeeze(0)


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (no amsgrad; ``weight_decay`` is L2 added to the gradient) as one multi-tensor launch.

    ``step(clip_norm=...)`` additionally folds global-norm clipping into the same pass (the gradients are scaled on
    the fly, ``.grad`` is left untouched).  ``shadow`` maps a parameter to a bf16 tensor of the same shape that
    receives the updated value (the GEMM operand copy)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.shadow: Dict[Tensor, Tensor] = {}
        self.post_step_hooks: List[Any] = []

    def attach_model(self, model: Any) -> None:
        """Lets the Adam kernel write the bf16 GEMM operands of ``model`` (an ``Allophant``) itself: no re-pack pass
        after the step, and the encoder's launch lists (raw pointers into those operands) stay valid."""
        packed = getattr(model._acoustic_model, "_packed", None)
        if packed is None:  # the from-scratch transformer encoder re-packs its (small) operands from the weight generation counter
            return
        self.shadow.update(packed.shadow_map())
        self.post_step_hooks.append(packed.after_fused_step)
        self.__dict__["_tables"] = {}  # the prefilled tables hold the shadow pointers

    # ---- host side of a step.  The arithmetic is three launches per 48 tensors; what used to cost more than those launches run
    # (3.7 ms of Python for ~430 parameters, exposed: the GPU has finished the backward pass by then) is rebuilt only when the
    # set of parameters with a gradient changes: parameter / moment pointers, sizes and bf16 shadows sit in prefilled ctypes
    # tables, a step writes the gradient pointers into them, and the per-parameter `step` tensors torch.optim.Adam keeps are
    # brought up to date lazily (`state_dict()`), from one Python counter per table.
    def _table_for(self, group_index: int, live: List[Tensor]):
        cache = self.__dict__.setdefault("_tables", {})
        key = (group_index, tuple(id(p) for p in live))
        entry = cache.get(group_index)
        if entry is not None and entry["key"] == key and all(p.data_ptr() == ptr for p, ptr in zip(live[:4], entry["first_ptrs"])):
            return entry
        if entry is not None:
            self._flush_steps()
        _checked(live)
        steps = set()
        for p in live:
            state = self.state[p]
            if len(state) == 0:
                state["step"] = torch.tensor(0.0)  # torch.optim.Adam keeps the step as a (host) tensor
                state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps.add(int(state["step"]))
        if len(steps) != 1:
            return None  # parameters at different step counts (a freshly unfrozen layer): the general path below
        n = len(live)
        table = (ctypes.c_void_p * (4 * n))()
        for i, p in enumerate(live):
            state = self.state[p]
            table[4 * i + 0] = p.data_ptr()
            table[4 * i + 2] = state["exp_avg"].data_ptr()
            table[4 * i + 3] = state["exp_avg_sq"].data_ptr()
        entry = dict(
            key=key, first_ptrs=[p.data_ptr() for p in live[:4]], params=live, table=table, grads=(ctypes.c_void_p * n)(), sizes=_sizes(live),
            shadows=_pointer_table([[self.shadow.get(p)] for p in live]), step=steps.pop(), flushed=True,
        )  # fmt: skip
        cache[group_index] = entry
        return entry

    def _flush_steps(self) -> None:
        for entry in self.__dict__.get("_tables", {}).values():
            if entry is not None and not entry["flushed"]:
                for p in entry["params"]:
                    self.state[p]["step"].fill_(float(entry["step"]))
                entry["flushed"] = True

    def state_dict(self):
        self._flush_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        self.__dict__["_tables"] = {}
        super().load_state_dict(state_dict)

    @torch.no_grad()
    def step(self, closure=None, clip_norm: Optional[float] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stream = ops._stream()
        plans = []
        for group_index, group in enumerate(self.param_groups):
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            entry = self._table_for(group_index, live)
            if entry is not None:
                table, grad_table = entry["table"], entry["grads"]
                for i, p in enumerate(live):
                    g = p.grad
                    if g.dtype is not torch.float32 or not g.is_cuda or not g.is_contiguous():
                        raise ValueError("allophant_b200.optim expects contiguous fp32 CUDA gradients")
                    pointer = g.data_ptr()
                    table[4 * i + 1] = pointer
                    grad_table[i] = pointer
            plans.append((group, live, entry))
        if not plans:
            return loss
        sumsq = None
        if clip_norm is not None:
            sumsq = torch.empty(1, device=plans[0][1][0].device, dtype=torch.float64)
            if all(entry is not None for _, _, entry in plans) and len(plans) == 1:
                entry = plans[0][2]
                check(lib.aph_multi_tensor_sumsq(entry["grads"], entry["sizes"], len(entry["params"]), sumsq.data_ptr(), stream), "aph_multi_tensor_sumsq")
            else:
                sumsq = sum_of_squares(_checked([p.grad for _, live, _ in plans for p in live]))
        for group, live, entry in plans:
            beta1, beta2 = group["betas"]
            hyper = (float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]))
            clip = (None if sumsq is None else sumsq.data_ptr(), 0.0 if clip_norm is None else float(clip_norm))
            if entry is not None:
                entry["step"] += 1
                entry["flushed"] = False
                check(
                    lib.aph_multi_tensor_adam(entry["table"], entry["shadows"], entry["sizes"], len(live), *hyper, entry["step"], *clip, stream),
                    "aph_multi_tensor_adam",
                )
                continue
            # general path: parameters of one group at different step counts
            counters = [self.state[p]["step"] for p in live]
            torch._foreach_add_(counters, 1.0)
            by_step: Dict[int, List[Tensor]] = {}
            for p, count in zip(live, torch.stack(counters).tolist()):
                by_step.setdefault(int(count), []).append(p)
            for step, params in by_step.items():
                _checked(params)
                _checked([p.grad for p in params])
                rows = [[p, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]] for p in params]
                shadows = [self.shadow.get(p) for p in params]
                check(
                    lib.aph_multi_tensor_adam(_pointer_table(rows), _pointer_table([[s] for s in shadows]), _sizes(params), len(params), *hyper, step, *clip, stream),
                    "aph_multi_tensor_adam",
                )
        engine.bump_weight_generation()  # parameters changed behind torch's version counters: packed operands are stale
        for hook in self.post_step_hooks:  # ... except the ones this step refreshed itself
            hook()
        return loss


class FusedSGD(torch.optim.Optimizer):
    """``torch.optim.SGD(lr, momentum, weight_decay)`` (dampening 0, no Nesterov: what ``config.py:300-312`` builds) as one
    multi-tensor launch, with the same clipping / bf16-shadow extras as ``FusedAdam``."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, weight_decay: float = 0.0) -> None:
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.shadow: Dict[Tensor, Tensor] = {}
        self.post_step_hooks: List[Any] = []

    attach_model = FusedAdam.attach_model

    @torch.no_grad()
    def step(self, closure=None, clip_norm: Optional[float] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sumsq = None
        if clip_norm is not None:
            grads = [p.grad for group in self.param_groups for p in group["params"] if p.grad is not None]
            if grads:
                sumsq = sum_of_squares(_checked(grads))
        for group in self.param_groups:
            momentum = float(group["momentum"])
            fresh, seasoned = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                first = momentum != 0.0 and "momentum_buffer" not in state
                if first:
                    state["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                (fresh if first else seasoned).append(p)
            for first_step, params in ((1, fresh), (0, seasoned)):
                if not params:
                    continue
                _checked(params)
                _checked([p.grad for p in params])
                rows = [[p, p.grad, self.state[p].get("momentum_buffer")] for p in params]
                shadows = [self.shadow.get(p) for p in params]
                check(
                    lib.aph_multi_tensor_sgd(
                        _pointer_table(rows), _pointer_table([[s] for s in shadows]), _sizes(params), len(params), float(group["lr"]), momentum,
                        float(group["weight_decay"]), first_step, None if sumsq is None else sumsq.data_ptr(),
                        0.0 if clip_norm is None else float(clip_norm), ops._stream(),
                    ),  # fmt: skip
                    "aph_multi_tensor_sgd",
                )
        engine.bump_weight_generation()
        for hook in self.post_step_hooks:
            hook()
        return loss


@dataclass
class WarmupInfo:
    model_size: int


class WarmupScheduler:
    """``config.py:107-173``: the learning-rate warm-up of "Attention is all you need" with an optional plateau."""

    def __init__(self, optimizer: torch.optim.Optimizer, warmup_info: WarmupInfo, warmup_steps: int, constant_steps: int = 0, factor: float = 2) -> None:
        self._optimizer = optimizer
        self._warmup_steps = warmup_steps
        self._constant_steps = constant_steps
        self._steps_until_decay = warmup_steps + constant_steps
        self._factor = factor
        self._model_size = warmup_info.model_size
        self._step = 1
        self._rate_value = self._rate(1)
        self._set_lr(self._rate_value)

    def _set_lr(self, rate: float) -> None:
        for group in self._optimizer.param_groups:
            group["lr"] = rate

    @property
    def last_lr(self) -> float:
        return self._rate_value

    def _rate(self, step: Optional[int] = None) -> float:
        if step is None:
            step = self._step
        if step < self._warmup_steps:
            return self._factor * (self._model_size ** (-0.5) * (step * self._warmup_steps ** (-1.5)))
        if step < self._steps_until_decay:
            return self._factor * (self._model_size ** (-0.5) * (self._warmup_steps ** (-0.5)))
        return self._factor * (self._model_size ** (-0.5) * ((step - self._constant_steps) ** (-0.5)))

    def step(self) -> None:
        self._step += 1
        self._rate_value = self._rate()
        self._set_lr(self._rate_value)

    def state_dict(self) -> Dict[str, Any]:
        return {"warmup_state": {"step": self._step, "rate": self._rate_value}}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self._step = int(state_dict["warmup_state"]["step"])
        self._rate_value = float(state_dict["warmup_state"]["rate"])


class OptimizerWrapper:
    """``config.py:213-276``: optimizer + optional learning-rate schedule, stepped together."""

    def __init__(self, optimizer: torch.optim.Optimizer, warmup_info: WarmupInfo) -> None:
        self._optimizer = optimizer
        self._warmup_info = warmup_info
        self._lr_scheduler: Optional[WarmupScheduler] = None

    def add_schedulers(self, warmup_steps: Optional[int], constant_steps: int = 0, factor: float = 2) -> None:
        self._lr_scheduler = None if warmup_steps is None else WarmupScheduler(self._optimizer, self._warmup_info, warmup_steps, constant_steps, factor)

    @property
    def optimizer(self) -> torch.optim.Optimizer:
        return self._optimizer

    def step(self, clip_norm: Optional[float] = None) -> None:
        if clip_norm is not None and isinstance(self._optimizer, (FusedAdam, FusedSGD)):
            self._optimizer.step(clip_norm=clip_norm)
        else:
            if clip_norm is not None:  # any other optimizer (a plain torch.optim.Adam as in the reference): clip first, as estimator.py:778-791 does
                clip_grad_norm_([p for group in self._optimizer.param_groups for p in group["params"]], clip_norm)
            self._optimizer.step()
        if self._lr_scheduler is not None:
            self._lr_scheduler.step()

    @property
    def param_groups(self) -> List[Dict[Any, Any]]:
        return self._optimizer.param_groups

    def current_learning_rate(self) -> float:
        return self.param_groups[0]["lr"]

    def state_dict(self) -> Dict[str, Any]:
        return {"lr_scheduler": None if self._lr_scheduler is None else self._lr_scheduler.state_dict(), "optimizer": self._optimizer.state_dict()}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self._optimizer.load_state_dict(state_dict["optimizer"])
        if self._lr_scheduler is not None:
            self._lr_scheduler.load_state_dict(state_dict["lr_scheduler"])


def adam_from_config(parameters: Iterable[Tensor], model_size: int, *, model: Any = None, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.98,
                     l2_regularization: float = 0.0, warmup_steps: Optional[int] = 2500, constant_steps: int = 10000, factor: float = 2) -> OptimizerWrapper:
    """``Adam.get_optimizer`` + ``OptimizerWrapper.add_schedulers`` with the defaults of ``default_config.toml:107-121``."""
    adam = FusedAdam(parameters, learning_rate, betas=(beta_1, beta_2), weight_decay=l2_regularization)
    if model is not None:
        adam.attach_model(model)
    wrapper = OptimizerWrapper(adam, WarmupInfo(model_size))
    wrapper.add_schedulers(warmup_steps, constant_steps, factor)
    return wrapper


def sgd_from_config(parameters: Iterable[Tensor], model_size: int, *, model: Any = None, learning_rate: float = 0.01, momentum: float = 0.0,
                    l2_regularization: float = 0.0, warmup_steps: Optional[int] = None, constant_steps: int = 0, factor: float = 2) -> OptimizerWrapper:
    """``SGD.get_optimizer`` (``config.py:300-312``) + ``OptimizerWrapper.add_schedulers``."""
    sgd = FusedSGD(parameters, learning_rate, momentum, l2_regularization)
    if model is not None:
        sgd.attach_model(model)
    wrapper = OptimizerWrapper(sgd, WarmupInfo(model_size))
    wrapper.add_schedulers(warmup_steps, constant_steps, factor)
    return wrapper


def optimizer_from_config(architecture: Any, model: Any) -> OptimizerWrapper:
    """``config.nn.optimizer.get_optimizer(model.parameters(), WarmupInfo(model.d_model))`` + ``add_schedulers(config.nn.lr_schedule)``
    (``estimator.py:982-983``, ``config.py:300-335``) for an ``allophant_b200.config.Architecture``: ``algorithm`` "adam" or
    "sgd", the warm-up schedule when ``lr_schedule`` is of type "warmup"."""
    options = dict(architecture.optimizer or {})
    algorithm = options.pop("algorithm", "adam")
    schedule = dict(architecture.lr_schedule or {})
    if schedule and schedule.get("type", "warmup") != "warmup":
        raise ValueError(f"Unsupported learning rate schedule: {schedule.get('type')!r}")
    schedule_options = dict(
        warmup_steps=schedule.get("warmup_steps") if schedule else None,
        constant_steps=schedule.get("constant_steps", 0),
        factor=schedule.get("factor", 2),
    )
    # ALL parameters, as the reference passes them (estimator.py:982): `freeze_feature_encoder` defaults to true, and what
    # `UnfreezeSchedule.step` unfreezes later must already be in the optimizer; Adam / SGD skip parameters without a `.grad`,
    # and the param_groups then match the reference's, so its optimizer state_dict loads
    parameters = list(model.parameters())
    if algorithm == "adam":
        return adam_from_config(
            parameters, model.d_model, model=model, learning_rate=options.get("learning_rate", 0.01), beta_1=options.get("beta_1", 0.9),
            beta_2=options.get("beta_2", 0.98), l2_regularization=options.get("l2_regularization", 0.0), **schedule_options,
        )  # fmt: skip
    if algorithm == "sgd":
        return sgd_from_config(
            parameters, model.d_model, model=model, learning_rate=options["learning_rate"], momentum=options.get("momentum", 0.0),
            l2_regularization=options.get("l2_regularization", 0.0), **schedule_options,
        )  # fmt: skip
    raise ValueError(f"Unknown optimizer algorithm: {algorithm!r}")
