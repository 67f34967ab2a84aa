"""Headline benchmark: multitask predict + greedy CTC decode, audio-seconds per second.

    python bench.py --gpus N --steps K --warmup W            # B200 path (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): Allophant Multitask, XLS-R-300M shape, random init, batch 32 x 10 s of
synthetic 16 kHz audio per GPU, all 36 attribute heads + composed phoneme head ('es'-sized inventory of 25),
log_softmax + greedy CTC decode of all 37 heads.  One step = one batch through that path.

The JSON line follows the driver's contract: `value` is device-timed with inputs resident in HBM, `e2e` is the
same metric through the public API (Estimator.predict + decode_predictions) with pinned-host inputs copied in and
decoded tokens copied out inside the timed region, `roofline` describes the dominant kernel (the tcgen05 GEMM over
the encoder's linear layers), `cpu_baseline` is the oracle timed on this box's host cores on a bounded sample.
Two sub-records ride on the same line so that the driver's `--gpus N` runs cover them too: `train` (BASELINE configs[2]:
the training step with the NCCL gradient all-reduce, the same per-rank batch on every rank, all-reduce bytes and exposed
milliseconds) and, at N = 1, `membound` (achieved GB/s of the HBM-bound kernels against the measured copy bandwidth).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from typing import Any, Dict, List, Optional, Tuple

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "audio-sec/sec, multitask predict+CTC"
UNIT = "audio-s/s"
BATCH = 32
SECONDS = 10
SAMPLE_RATE = 16000
INVENTORY = 25


def encoder_flops(frames: int, hidden: int = 1024, ffn: int = 4096, layers: int = 24) -> Dict[str, float]:
    """Algorithmic forward FLOPs per utterance."""
    linear = layers * (8 * hidden * hidden + 4 * hidden * ffn) * frames
    attention = layers * 4 * hidden * frames * frames
    return {"linear": float(linear), "attention": float(attention)}


def utterance_flops(samples: int) -> float:
    lengths, length = [], samples
    for kernel, stride in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        length = (length - kernel) // stride + 1
        lengths.append(length)
    conv = 2 * 1 * 10 * 512 * lengths[0] + sum(2 * 512 * k * 512 * l for k, l in zip((3, 3, 3, 3, 2, 2), lengths[1:]))
    frames = lengths[-1]
    proj = 2 * 512 * 1024 * frames
    pos = 2 * 64 * 128 * 1024 * frames
    enc = encoder_flops(frames)
    heads = 2 * 1024 * (36 * 4 + 640) * frames + 2 * 640 * (INVENTORY + 1) * frames
    return conv + proj + pos + enc["linear"] + enc["attention"] + heads


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int) -> None:
        self.index = index
        self.samples: List[List[str]] = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self) -> bool:
        """NVML in-process (nvidia_ml_py): a sample costs microseconds, so even a 150 ms timed region gets a dozen of them;
        the nvidia-smi subprocess below (~100 ms per call) is the fallback."""
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.index
            if visible:
                entries = [entry.strip() for entry in visible.split(",") if entry.strip()]
                if index < len(entries) and entries[index].isdigit():
                    index = int(entries[index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            max_clock = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
            flags = (
                ("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown),
                ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap),
            )
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            return False
        while not self._stop.is_set():
            try:
                clock = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                power = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
                self.samples.append([str(clock), str(max_clock), str(power)] + ["Active" if reasons & bit else "Not Active" for _, bit in flags])
            except Exception:
                pass
            self._stop.wait(0.01)
        return True

    def _run(self) -> None:
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()  # fmt: skip
                if out:
                    self.samples.append([cell.strip() for cell in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> Dict[str, Any]:
        clocks, reasons, max_clock = [], set(), None
        for cells in self.samples:
            try:
                clocks.append(float(cells[0]))
                max_clock = float(cells[1])
            except (ValueError, IndexError):
                continue
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), cells[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        clocks.sort()
        return {
            "sm_mhz": clocks[len(clocks) // 2] if clocks else None,
            "sm_max_mhz": max_clock,
            "reasons": sorted(reasons),
            "samples": len(clocks),
        }


def measured_peaks() -> Dict[str, Any]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as file:
            peaks = json.load(file)
        peaks["source"] = "measured (MEASURED_PEAKS.json)"
        return peaks
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------
# model construction (random init of the named architecture; synthetic inventory — no network, no CSV)
# --------------------------------------------------------------------------------------------------
# the end-to-end stream calls the public API with its CUDA-graph option (one captured graph per input shape, bit-identical to the
# eager step: tests/test_gpu_e2e_parity.py::test_cuda_graph_predict_equals_eager); BENCH_E2E_GRAPH=0 times the eager call
E2E_CUDA_GRAPH = os.environ.get("BENCH_E2E_GRAPH", "1") != "0"
HIERARCHICAL = False  # --hierarchical: BASELINE configs[3] (phoneme head depends on OUTPUT + all 36 attribute heads)


def build_estimator(device: str):
    from allophant_b200.config import Config, PhonemeLayerType
    from allophant_b200.estimator import Estimator, attribute_graph_from_config
    from allophant_b200.phonetic_features import PhoneticAttributeIndexer

    torch.manual_seed(2)
    config = Config.default()
    config.nn.projection.phoneme_layer = PhonemeLayerType.SHARED
    names = [entry.name for entry in config.nn.projection.classes]
    if HIERARCHICAL:
        for entry in config.nn.projection.classes:
            if entry.name == "phoneme":
                entry.dependencies = ["OUTPUT", *[name for name in names if name != "phoneme"]]
    indexer = PhoneticAttributeIndexer.synthetic(max(109, INVENTORY), names, n_categories=3, seed=1, training_inventory=60)
    graph = attribute_graph_from_config(config, indexer)
    estimator = Estimator.from_config(config, 1, SAMPLE_RATE, graph, indexer, device=device, load_pretrained_weights=False)
    inventory = [f"p{index}" for index in range(INVENTORY)]
    return estimator, indexer.composition_feature_matrix(inventory)


def cpu_reference_step(oracle, audio, lengths, tfi) -> None:
    from oracle import restatement

    outputs, frames = oracle.predict(audio, lengths, None, tfi)
    for value in outputs.values():
        restatement.greedy_ctc_decode(value.transpose(1, 0).contiguous(), frames)


def time_cpu_reference(n_utt: int, steps: int, warmup: int, budget_s: float = 1e9) -> Dict[str, Any]:
    """The reference's CPU path (oracle port: HF encoder + restated heads/decoder) on this box's host cores."""
    from oracle import restatement

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    spec = restatement.multitask_spec(n_train_phonemes=60)
    oracle = restatement.OracleModel(spec)
    samples = SECONDS * SAMPLE_RATE
    audio = restatement.synthetic_audio(n_utt, samples, seed=0)
    lengths = torch.full((n_utt,), samples, dtype=torch.long)
    tfi = torch.randint(0, 3, (INVENTORY, 36), generator=torch.Generator().manual_seed(1))
    for _ in range(warmup):
        cpu_reference_step(oracle, audio, lengths, tfi)
    times = []
    for _ in range(steps):
        start = time.perf_counter()
        cpu_reference_step(oracle, audio, lengths, tfi)
        times.append(time.perf_counter() - start)
        if sum(times) > budget_s:
            break
    total = sum(times)
    timed = len(times)
    return {
        "value": n_utt * SECONDS * timed / total,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{timed} x (batch {n_utt} x {SECONDS} s, fp32, all 37 heads + greedy decode), {warmup} warm-up"
        + ("" if timed == steps else f"; stopped after {timed} of {steps} steps (time budget {budget_s:.0f} s)"),
        "ms_per_step": 1000.0 * total / timed,
        "steps_timed": timed,
    }


def run_reference_arm(args) -> None:
    """The reference's CPU path on this arm's own configuration: every step is one whole batch (BATCH x SECONDS, all 37 heads
    + greedy decode) through the oracle port on all host cores.  One warm-up step; timing stops early (and says so in
    `cpu_baseline.sample` / `steps_timed`) once `BENCH_REFERENCE_BUDGET_S` (default 240 s) of CPU time is spent, so that a
    box with few cores still finishes within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    result = time_cpu_reference(BATCH, max(1, args.steps), 1, budget_s=float(os.environ.get("BENCH_REFERENCE_BUDGET_S", "240")))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": result["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": result["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {k: result[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "steps_timed": result["steps_timed"],
        "e2e": {"value": result["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int) -> Dict[str, Any]:
    if HIERARCHICAL:
        name = "BASELINE configs[3]: Allophant Hierarchical (phoneme head over OUTPUT + 36 attribute posteriors, XLS-R-300M shape"
    elif (BATCH, SECONDS, INVENTORY) == (32, 10, 25):
        name = "BASELINE configs[1]: Allophant Multitask (XLS-R-300M shape"
    else:
        name = "BASELINE configs[4] sweep point: Allophant Multitask (XLS-R-300M shape"
    return {
        "workload": f"{name}, random init) inference, "
        f"batch {BATCH} x {SECONDS} s synthetic 16 kHz audio per GPU, 36 attribute heads + composed phoneme head "
        f"(inventory {INVENTORY}), log_softmax + greedy CTC decode of all 37 heads",
        "batch_per_gpu": BATCH,
        "seconds_per_utterance": SECONDS,
        "parallelism": f"dp{n_gpus} (independent utterance shards, no collective)",
        "l2": "per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush between iterations",
        "e2e": "a stream of 10 x steps batches through Estimator.predict + decode_predictions_async: pinned host audio copied in on a "
        "copy stream, decoded tokens copied out, CTCHypothesis lists built on the host while the next two batches compute" + (
            "; predict(cuda_graph=True)" if E2E_CUDA_GRAPH else ""),
    }


def cupti_gemm_time(step, is_encoder_linear) -> Optional[Tuple[int, float]]:
    """(launches, summed kernel milliseconds) of the encoder-linear GEMMs of one UN-BRACKETED step, from CUPTI kernel
    records (torch.profiler): every `ops.run_gemm` call launches exactly one `gemm_bf16_kernel` on the one stream of the
    step, so the k-th kernel record belongs to the k-th call; `is_encoder_linear(args)` picks the same launches the event
    brackets time.  None when CUPTI is not available or the records do not line up."""
    from allophant_b200 import ops

    flags: List[bool] = []
    original = ops.run_gemm

    def recording_gemm(gemm_args):
        flags.append(bool(is_encoder_linear(gemm_args)))
        original(gemm_args)

    try:
        from torch.profiler import ProfilerActivity, profile

        import warnings

        torch.cuda.synchronize()
        ops.run_gemm = recording_gemm
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    step()
                    torch.cuda.synchronize()
        finally:
            ops.run_gemm = original
        kernels = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gemm_bf16_kernel" in e.name),
                         key=lambda e: e.time_range.start)  # fmt: skip
    except Exception:
        return None
    if len(kernels) != len(flags) or not any(flags):
        return None
    total_us = sum(event.time_range.end - event.time_range.start for event, flag in zip(kernels, flags) if flag)
    return sum(flags), total_us / 1000.0


def kernel_only_roofline(step, is_encoder_linear, flops: float, peak: float, step_ms: float) -> Optional[Dict[str, Any]]:
    """Cross-check of `roofline.achieved` without the event brackets: a CUDA event between two launches keeps the GPU front
    end from overlapping the launch of the next kernel with the execution of the previous one, so every bracketed GEMM pays
    its full launch latency (10-16 us for a kernel whose parameters hold four tensor maps)."""
    measured = cupti_gemm_time(step, is_encoder_linear)
    if measured is None:
        return None
    count, total_ms = measured
    achieved = flops / (total_ms / 1000.0) / 1e12
    return {
        "source": "CUPTI kernel records of one un-bracketed step (torch.profiler), the same launches as the event brackets",
        "launches": count, "avg_launch_ms": total_ms / max(1, count), "achieved": achieved, "frac": achieved / peak,
        "share_of_step": total_ms / step_ms,
    }  # fmt: skip


# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    from allophant_b200 import ops
    from allophant_b200.dataset_processing import Batch
    from allophant_b200.predictions import decode_predictions_async

    process = setup_process()
    world, rank, local_rank, device, distributed = process
    if distributed:
        import torch.distributed as dist

    estimator, tfi = build_estimator(device)
    tfi_dev = tfi.to(device)
    samples = SECONDS * SAMPLE_RATE
    generator = torch.Generator().manual_seed(rank)
    host_audio = (0.1 * torch.randn(BATCH, samples, generator=generator)).pin_memory()
    host_lengths = torch.full((BATCH,), samples, dtype=torch.long).pin_memory()
    host_languages = torch.zeros(BATCH, dtype=torch.long).pin_memory()
    resident = Batch(host_audio.to(device), host_lengths.to(device), host_languages.to(device))

    def device_step():
        predictions = estimator.predict(resident, tfi_dev)
        cache = predictions._decode_cache
        return ops.ctc_greedy_collapse(cache["argmax"], cache["maxlp"], cache["frames32"], cache["n_utt"], cache["seq"],
                                       cache["argmax"].shape[0] * cache["n_utt"], 0)  # fmt: skip

    copy_stream = torch.cuda.Stream(device=device)

    def e2e_launch():
        """Host -> device copy of one batch from pinned memory (on a copy stream, so it overlaps the previous batch's
        kernels), predict, greedy decode and the device -> host copy of the decoded tokens: everything is enqueued,
        nothing synchronises."""
        with torch.cuda.stream(copy_stream):
            batch = Batch(host_audio, host_lengths, host_languages).to(device, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record()
        torch.cuda.current_stream().wait_event(copied)
        for tensor in (batch.audio_features, batch.lengths, batch.language_ids):
            tensor.record_stream(torch.cuda.current_stream())
        predictions = estimator.predict(batch, tfi_dev, cuda_graph=E2E_CUDA_GRAPH)
        return decode_predictions_async(predictions)

    def e2e_stream(steps: int):
        """A streaming client of the public API: while the GPU works on batch i+1 the host turns the copied-back
        tokens of batch i into CTCHypothesis lists.  Every batch is copied in, computed, copied out and decoded."""
        in_flight: List[Any] = []
        hypotheses = None
        trace = os.environ.get("BENCH_E2E_TRACE") == "1"
        host = {"launch": 0.0, "wait": 0.0, "assemble": 0.0}
        marks = []
        for _ in range(steps):
            t0 = time.perf_counter()
            if trace:
                begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                begin.record()
            in_flight.append(e2e_launch())
            if trace:
                end.record()
                marks.append((begin, end))
            host["launch"] += time.perf_counter() - t0
            if len(in_flight) > 2:  # two batches in flight (the pinned result buffers rotate over three slots): 17.4 -> 16.8 ms/step
                pending = in_flight.pop(0)
                t1 = time.perf_counter()
                if trace:
                    pending._event.synchronize()
                t2 = time.perf_counter()
                hypotheses = pending.result()
                host["wait"] += t2 - t1
                host["assemble"] += time.perf_counter() - t2
        while in_flight:
            hypotheses = in_flight.pop(0).result()
        if trace:
            torch.cuda.synchronize()
            busy = sum(a.elapsed_time(b) for a, b in marks) / steps
            span = marks[0][0].elapsed_time(marks[-1][1]) / steps
            print(f"[e2e trace] steps {steps}: GPU busy {busy:.2f} ms/step, GPU span {span:.2f} ms/step; host launch {host['launch'] / steps * 1e3:.2f}, "
                  f"wait {host['wait'] / steps * 1e3:.2f}, assemble {host['assemble'] / steps * 1e3:.2f} ms/step", file=sys.stderr, flush=True)  # fmt: skip
        return hypotheses

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps: int) -> float:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(steps):
            step()
        end.record()
        barrier()
        elapsed = torch.tensor([start.elapsed_time(end)], device=device)
        if distributed:
            dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        return float(elapsed.item())

    for _ in range(max(3, args.warmup)):
        device_step()
    ops.reset_launch_count()
    if os.environ.get("BENCH_NO_CLOCK_SAMPLER") == "1":  # diagnostic: is the nvidia-smi poller perturbing what follows?
        elapsed_ms = timed(device_step, args.steps)
        clocks = {}
    else:
        with ClockSampler(local_rank) as sampler:
            elapsed_ms = timed(device_step, args.steps)
        clocks = sampler.summary()
    launches = ops.launch_count() // args.steps

    e2e_stream(2)
    barrier()
    e2e_steps = 10 * args.steps  # a stream of batches: the one-batch pipeline fill is amortised over 10K batches
    wall_start = time.perf_counter()
    result = e2e_stream(e2e_steps)  # the last result() waits for the last device-to-host copy
    assert len(result) == 37 and len(result["phoneme"]) == BATCH
    barrier()
    e2e_seconds = torch.tensor([time.perf_counter() - wall_start], device=device)
    if distributed:
        dist.all_reduce(e2e_seconds, op=dist.ReduceOp.MAX)
    d2h_bytes = 0
    cache = estimator.predict(resident, tfi_dev)._decode_cache
    n_seq = cache["argmax"].shape[0] * cache["n_utt"]
    d2h_bytes = n_seq * cache["seq"] * 4 * 2 + n_seq * 8
    h2d_bytes = host_audio.numel() * 4 + host_lengths.numel() * 8 + host_languages.numel() * 8

    audio_seconds = world * BATCH * SECONDS * args.steps
    value = audio_seconds / (elapsed_ms / 1000.0)
    e2e_value = world * BATCH * SECONDS * e2e_steps / float(e2e_seconds.item())

    # ---- roofline of the dominant kernel: the tcgen05 GEMM over the encoder's 24 x {QKV, out, FFN1, FFN2} ----
    roofline = None
    cpu_baseline = None
    if rank == 0:
        length = samples
        for kernel, stride in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
            length = (length - kernel) // stride + 1
        frames = length
        gemm_events: List[Any] = []
        original = ops.run_gemm

        def timed_gemm(gemm_args):
            if gemm_args.k in (1024, 4096) and gemm_args.n in (1024, 3072, 4096) and gemm_args.mode == 0:
                start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                start.record()
                original(gemm_args)
                end.record()
                gemm_events.append((start, end))
            else:
                original(gemm_args)

        ops.run_gemm = timed_gemm
        try:
            device_step()
            torch.cuda.synchronize()
        finally:
            ops.run_gemm = original
        bracketed_ms = sum(start.elapsed_time(end) for start, end in gemm_events)
        # The same 96 launches (same argument structs, same operands and outputs, in the order of the step) replayed back to
        # back between ONE pair of events: an event between two launches keeps the front end from overlapping the next kernel's
        # launch (and its programmatic-dependent-launch prologue) with the kernel before it, 4-8 us per launch that the real step
        # does not pay.  The per-launch brackets are kept as `bracketed`.
        encoder_gemms: List[Any] = []

        def recording_gemm(gemm_args):
            if gemm_args.k in (1024, 4096) and gemm_args.n in (1024, 3072, 4096) and gemm_args.mode == 0:
                encoder_gemms.append(gemm_args)
            original(gemm_args)

        ops.run_gemm = recording_gemm
        try:
            device_step()
            torch.cuda.synchronize()
        finally:
            ops.run_gemm = original
        replay_rounds = 5
        for gemm_args in encoder_gemms:
            original(gemm_args)
        replay_start, replay_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        replay_start.record()
        for _ in range(replay_rounds):
            for gemm_args in encoder_gemms:
                original(gemm_args)
        replay_end.record()
        torch.cuda.synchronize()
        gemm_ms = replay_start.elapsed_time(replay_end) / replay_rounds
        flops = encoder_flops(frames)["linear"] * BATCH
        peaks = measured_peaks()
        achieved = flops / (gemm_ms / 1000.0) / 1e12
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        roofline = {
            "kernel": "aph::gemm_bf16_kernel (encoder QKV / out-proj / FFN1 / FFN2, 96 launches per step)",
            "bound": "tensor",
            "achieved": achieved,
            "peak": peak,
            "unit": "TFLOP/s",
            "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over one layer's QKV / out-proj / FFN1 /
            # FFN2 launches of the `ncu --set full` capture in profiles/r02_ncu_summary.md (89.9 / 118.1 / 124.3 / 263.5 MB: the
            # plan now pads the batch to a 512-frame length bucket, round 1: 83.4 / 110.8 / 117.3 / 249.7): equal to the algorithmic
            # operand + residual bytes, i.e. no re-reads
            "traffic": 148.9e6,
            "traffic_unit": "bytes per launch",
            "traffic_source": "constant from profiles/r02_ncu_summary.md (one `ncu --set full` capture of this step; not re-measured per run)",
            "peak_source": f"bf16_tflops_sustained, {peaks['source']}",
            "launches": len(encoder_gemms),
            "avg_launch_ms": gemm_ms / max(1, len(encoder_gemms)),
            "share_of_step": gemm_ms / (elapsed_ms / args.steps),
            "timing": f"CUDA events around {replay_rounds} back-to-back replays of the step's {len(encoder_gemms)} encoder GEMM launches (same argument structs and buffers)",
            "bracketed": {
                "source": "one CUDA event pair around every encoder GEMM launch inside a whole step (includes the launch gap the bracket itself creates)",
                "launches": len(gemm_events),
                "avg_launch_ms": bracketed_ms / max(1, len(gemm_events)),
                "achieved": flops / (bracketed_ms / 1000.0) / 1e12,
                "frac": flops / (bracketed_ms / 1000.0) / 1e12 / peak,
            },
        }
        # the CUPTI cross-check runs with programmatic dependent launch off: with it a kernel's record starts while its
        # predecessor still drains (the wait is inside the kernel), which would inflate every duration
        pdl_was = ops.set_pdl(False)
        try:
            kernel_only = kernel_only_roofline(
                device_step, lambda g: g.k in (1024, 4096) and g.n in (1024, 3072, 4096) and g.mode == 0, flops, peak, elapsed_ms / args.steps
            )
        finally:
            ops.set_pdl(pdl_was)
        if kernel_only is not None:
            kernel_only["note"] = "measured with programmatic dependent launch off (records of overlapping kernels would include the wait)"
            roofline["kernel_only"] = kernel_only
        if not args.skip_cpu_baseline and world == 1:  # rank 0 at N = 1 only: a bounded sample (~10 s of CPU work) of the same workload
            baseline = time_cpu_reference(8, 3, 1)
            cpu_baseline = {k: baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    ragged = None
    if rank == 0 and world == 1 and not args.skip_ragged:
        ragged = measure_ragged_stream(estimator, tfi_dev, device, value)

    # ---- sub-records: the training / data-parallel path (BASELINE configs[2]) and the HBM-bound kernels, measured in this same run
    del estimator, resident
    torch.cuda.empty_cache()
    train = None
    if not args.skip_train:
        record = measure_training(process, args.train_steps, 3, False, detail=True)
        if record is not None:
            train = {key: record[key] for key in ("metric", "value", "unit", "ms_per_step", "steps", "gpu_launches", "e2e", "roofline")}
            train["config"] = record["config"]
    membound = None
    if rank == 0 and world == 1 and not args.skip_membound:
        membound = measure_membound(device)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "config": workload_config(world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "tflops_per_gpu": utterance_flops(samples) * BATCH / (elapsed_ms / args.steps / 1000.0) / 1e12,
        "train": train,
        "membound": membound,
        "ragged": ragged,
    }
    print(json.dumps(line), flush=True)


def measure_ragged_stream(estimator, tfi_dev, device: str, fixed_shape_value: float) -> Dict[str, Any]:
    """A stream of batches as `MaxFrameBatchSampler` (allophant/batching.py:94-139) forms them — utterances of U[3 s, 15 s] in
    random order, batched until batch size x longest utterance would exceed the frame budget of the fixed-shape workload
    (32 x 10 s) — through the same `Estimator.predict` + greedy decode step.  Every batch has its own (utterances, padded length):
    the model pads each up to a 64-frame length bucket and runs it from a launch list carved out of one shared workspace arena
    (allophant_b200.engine.bucket_samples / WorkspaceArena).  Reported: throughput over the VALID audio after one warm-up pass
    over the stream, launch lists built and arena re-allocations DURING the timed passes (both must be 0), and the ratio to the
    fixed-shape number of this run."""
    from allophant_b200 import ops
    from allophant_b200.batching import MaxFrameBatchSampler
    from allophant_b200.dataset_processing import Batch

    generator = torch.Generator().manual_seed(11)
    n_utterances = 640
    lengths = torch.randint(3 * SAMPLE_RATE, 15 * SAMPLE_RATE + 1, (n_utterances,), generator=generator)
    order = torch.randperm(n_utterances, generator=generator).tolist()
    budget = BATCH * SECONDS * SAMPLE_RATE
    batches = []
    for indices in MaxFrameBatchSampler(order, budget, lengths):
        batch_lengths = lengths[indices]
        longest = int(batch_lengths.max())
        audio = 0.1 * torch.randn(len(indices), longest, generator=generator)
        audio *= (torch.arange(longest)[None, :] < batch_lengths[:, None]).float()
        batches.append(Batch(audio.to(device), batch_lengths.to(device), torch.zeros(len(indices), dtype=torch.long, device=device)))
    acoustic = estimator.model.acoustic_model

    def run_stream() -> None:
        for batch in batches:
            predictions = estimator.predict(batch, tfi_dev)
            cache = predictions._decode_cache
            ops.ctc_greedy_collapse(cache["argmax"], cache["maxlp"], cache["frames32"], cache["n_utt"], cache["seq"],
                                    cache["argmax"].shape[0] * cache["n_utt"], 0)  # fmt: skip

    run_stream()  # warm-up: every bucket of the stream gets its launch list, the arena reaches its size
    torch.cuda.synchronize()
    builds, arena = acoustic.plan_builds, acoustic._arena.buffer
    allocated = torch.cuda.memory_allocated()
    passes = 2
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(passes):
        run_stream()
    end.record()
    torch.cuda.synchronize()
    seconds = start.elapsed_time(end) / 1000.0
    valid = float(lengths.sum()) / SAMPLE_RATE * passes
    padded = sum(batch.audio_features.shape[0] * batch.audio_features.shape[1] for batch in batches) / SAMPLE_RATE * passes
    shapes = {(batch.audio_features.shape[0], batch.audio_features.shape[1]) for batch in batches}
    return {
        "metric": "audio-sec/sec of valid audio, multitask predict+CTC over a MaxFrameBatchSampler stream",
        "value": valid / seconds,
        "unit": UNIT,
        "padded_audio_value": padded / seconds,
        "batches_per_pass": len(batches),
        "distinct_batch_shapes": len(shapes),
        "launch_lists": len(acoustic._plans),
        "launch_lists_built_while_timed": acoustic.plan_builds - builds,
        "arena_reallocated_while_timed": acoustic._arena.buffer is not arena,
        "arena_bytes": int(acoustic._arena.capacity()),
        "device_memory_growth_while_timed_bytes": int(torch.cuda.memory_allocated() - allocated),
        "bucket_frames": acoustic.bucket_frames,
        "padded_over_fixed_shape": (padded / seconds) / fixed_shape_value,
        "valid_over_fixed_shape": (valid / seconds) / fixed_shape_value,
        "workload": f"{n_utterances} utterances U[3 s, 15 s], frame budget {BATCH} x {SECONDS} s per batch, {passes} timed passes after one warm-up pass",
    }


# --------------------------------------------------------------------------------------------------
# HBM-bound kernels at BASELINE sizes: achieved GB/s (algorithmic bytes) against the measured copy bandwidth
# --------------------------------------------------------------------------------------------------
def measure_membound(device: str) -> List[Dict[str, Any]]:
    from allophant_b200 import ops

    peak = float(measured_peaks()["hbm_gbs"])
    records: List[Dict[str, Any]] = []

    def timeit(fn, iters=20, warm=3) -> float:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(iters):
            fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / iters

    def report(kernel: str, shape: str, ms: float, nbytes: float, note: str = "") -> None:
        gbs = nbytes / ms / 1e6
        records.append({"kernel": kernel, "shape": shape, "us": ms * 1e3, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "peak_gbs": peak,
                        "frac": gbs / peak, "note": note})  # fmt: skip

    def synthetic_labels(frames, n_classes, seed, fraction=0.25):
        generator = torch.Generator().manual_seed(seed)
        lengths = (frames.double() * fraction).floor().long().clamp_min(1)
        labels = torch.zeros(len(frames), int(lengths.max()), dtype=torch.long)
        for row, length in enumerate(lengths.tolist()):
            labels[row, :length] = torch.randint(1, n_classes, (length,), generator=generator)
        return labels, lengths

    # configs[3]: the wide composed phoneme head, 128 x 499 frames x 3184 classes (0.81 GB of logits: far beyond L2)
    rows, width = 128 * 499, 3184
    logits = torch.randn(rows, width, device=device)
    out = torch.empty_like(logits)
    am = torch.empty(rows, dtype=torch.int32, device=device)
    mx = torch.empty(rows, device=device)
    report("log_softmax_wide (+argmax, max)", "63872 x 3184 fp32", timeit(lambda: ops.log_softmax_wide(logits, width, rows, width, out, width, am, mx)), 2.0 * rows * width * 4)
    report("torch.log_softmax (library, same shape)", "63872 x 3184 fp32", timeit(lambda: torch.log_softmax(logits, -1), iters=10), 2.0 * rows * width * 4, "library bar")
    del logits, out

    # configs[1]: the 36 narrow heads packed [15968 x 144]
    rows2 = 32 * 499
    packed = torch.randn(rows2, 144, device=device)
    col = torch.arange(0, 144, 4, dtype=torch.int32, device=device)
    wid = torch.full((36,), 4, dtype=torch.int32, device=device)
    off = torch.arange(36, dtype=torch.int64, device=device) * rows2 * 4
    out2 = torch.empty(rows2 * 144, device=device)
    am2 = torch.empty(36, rows2, dtype=torch.int32, device=device)
    mx2 = torch.empty(36, rows2, device=device)
    report("log_softmax_heads 36 heads (+argmax, max)", "15968 x 144 fp32", timeit(lambda: ops.log_softmax_heads(packed, 144, rows2, 0, 144, col, wid, off, 36, out2, am2, mx2)),
           2.0 * rows2 * 144 * 4 + 36 * rows2 * 8, "18 MB: latency-bound at this size")

    # LayerNorm of the residual stream: rotate over 4 inputs so that the 65 MB rows do not stay in the 126 MB L2
    xs = [torch.randn(rows2, 1024, device=device) for _ in range(4)]
    g, b = torch.ones(1024, device=device), torch.zeros(1024, device=device)
    o16 = torch.empty(rows2, 1024, device=device, dtype=torch.bfloat16)

    def layer_norms():
        for x in xs:
            ops.layernorm_rows(x, rows2, 1024, 1024, g, b, 1e-5, out_bf16=o16, ld_bf16=1024)

    report("layernorm_rows fp32 -> bf16", "15968 x 1024", timeit(layer_norms) / 4, rows2 * 1024 * 6.0)
    del xs
    rows3 = 32 * 15999
    xb = torch.randn(rows3, 512, device=device).bfloat16()
    g5, b5 = torch.ones(512, device=device), torch.zeros(512, device=device)
    report("layernorm_rows + GELU bf16 in place (conv layer 1)", "511968 x 512", timeit(lambda: ops.layernorm_rows(xb, rows3, 512, 512, g5, b5, 1e-5, gelu=True, out_bf16=xb, ld_bf16=512), iters=10), rows3 * 512 * 4.0)
    del xb

    # conv0 fused (waveform normalisation + conv 1->512 k10 s5 + LayerNorm + GELU): 32 x 160000 samples -> [32, 31999, 512] bf16
    audio = torch.randn(32, 160000, device=device) * 0.1
    lengths = torch.full((32,), 160000, dtype=torch.int64, device=device)
    stats = torch.empty(32, 3, dtype=torch.float64, device=device)
    mr = torch.empty(32, 2, device=device)
    ops.wave_stats(audio, lengths, stats, mr)
    w0 = torch.randn(512, 10, device=device) * 0.3
    o0 = torch.empty(32, 31999, 512, device=device, dtype=torch.bfloat16)
    report("conv0 + waveform norm + LayerNorm + GELU", "32 x 160000 -> 32 x 31999 x 512 bf16", timeit(lambda: ops.conv0_ln_gelu(audio, lengths, mr, w0, g5, g5, b5, 1e-5, o0), iters=10),
           32 * 160000 * 4.0 + 32 * 31999 * 512 * 2.0)
    del o0, audio

    # CTC, configs[2] shape per GPU (8 utterances) and the whole global batch (64): 36 heads c=4 + phoneme c=501, T' = 749
    for n_utt in (8, 64):
        frames = 749
        input_lengths = torch.randint(150, frames + 1, (n_utt,), generator=torch.Generator().manual_seed(4))
        input_lengths[0] = frames
        classes = [4] * 36 + [501]
        log_probs, labels_l, lens_l = [], [], []
        for head, c in enumerate(classes):
            log_probs.append(torch.log_softmax(torch.randn(frames, n_utt, c, device=device), -1))
            labels, label_lengths = synthetic_labels(input_lengths, c, seed=head)
            labels_l.append(labels.to(device))
            lens_l.append(label_lengths.to(device))
        il = input_lengths.to(device)
        problem = ops.CtcProblem(log_probs, labels_l, lens_l, il, batch_first=False, need_grad=True)
        scale = torch.ones(len(classes), device=device)
        valid = int(input_lengths.sum())
        lp_bytes = float(sum(valid * c * 4 for c in classes))
        alpha_bytes = float(valid * problem.s_pad * 4 * len(classes))
        note = "latency-bound recursion over T' sequential frames: bytes / time is far from the HBM roofline by construction"
        report("ctc_forward (alpha kept)", f"{n_utt} utt x 37 heads, T' <= 749", timeit(problem.forward, iters=10), lp_bytes + alpha_bytes, note)
        report("ctc_backward (beta + gradient)", f"{n_utt} utt x 37 heads, T' <= 749", timeit(lambda: problem.backward(scale), iters=10), 2 * lp_bytes + alpha_bytes, note)

    am3 = torch.randint(0, 4, (37, rows2), dtype=torch.int32, device=device)
    mx3 = torch.randn(37, rows2, device=device)
    fl = torch.full((32,), 499, dtype=torch.int32, device=device)
    report("ctc_greedy_collapse", "37 heads x 32 utt x 499", timeit(lambda: ops.ctc_greedy_collapse(am3, mx3, fl, 32, 499, 37 * 32, 0)), 37 * rows2 * 8.0, "4.7 MB: latency-bound")
    torch.cuda.empty_cache()
    return records


# --------------------------------------------------------------------------------------------------
# --workload train: BASELINE configs[2] — forward + multi-head CTC + backward (+ NCCL gradient all-reduce)
# --------------------------------------------------------------------------------------------------
TRAIN_BATCH = 8          # utterances per GPU (64 over 8 GPUs)
TRAIN_LANGUAGES = 34
TRAIN_PHONES = 500       # shared phones P and phoneme classes Q of the synthetic allophone layer


def build_training_estimator(device: str):
    """Default architecture (Multitask + allophone layer + composed phone embeddings) over a synthetic inventory."""
    import numpy as np

    from allophant_b200.config import Config
    from allophant_b200.estimator import Estimator, attribute_graph_from_config
    from allophant_b200.phonetic_features import AllophoneData, ArticulatoryAttributes, LanguageAllophoneMappings, PhoneticAttributeIndexer

    torch.manual_seed(2)
    config = Config.default()
    names = [entry.name for entry in config.nn.projection.classes]
    features = [name for name in names if name != "phoneme"]
    rng = np.random.default_rng(1)
    table = rng.integers(0, 3, size=(TRAIN_PHONES, len(features)))
    table[:3, :] = np.arange(3)[:, None]
    phones = [f"ph{index}" for index in range(TRAIN_PHONES)]
    shared = ArticulatoryAttributes(phones, features, table, {f: ["0", "1", "2"] for f in features})
    allophones = {}
    for language in range(TRAIN_LANGUAGES):  # identity plus up to two more allophones per phoneme, 40-phoneme inventories
        inventory = sorted(rng.choice(TRAIN_PHONES, size=40, replace=False).tolist())
        allophones[language] = {
            int(phoneme): sorted({int(phoneme), *[int(p) for p in rng.choice(TRAIN_PHONES, size=2, replace=False)]}) for phoneme in inventory
        }
    mappings = LanguageAllophoneMappings(allophones, [f"l{index}" for index in range(TRAIN_LANGUAGES)], phones)
    indexer = PhoneticAttributeIndexer(
        shared, [f"p{index}" for index in range(TRAIN_PHONES)], features, features + ["phoneme"], mappings, AllophoneData(shared)
    )
    graph = attribute_graph_from_config(config, indexer)
    estimator = Estimator.from_config(config, 1, SAMPLE_RATE, graph, indexer, device=device, load_pretrained_weights=False)
    return estimator, allophones


def setup_process():
    """(world, rank, local_rank, device, distributed) of this process; joins the NCCL group when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: allophant_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    distributed = world > 1
    if distributed:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(device))
    return world, rank, local_rank, device, distributed


def measure_training(process, steps: int, warmup: int, eval_arithmetic: bool, detail: bool = True) -> Optional[Dict[str, Any]]:
    """BASELINE configs[2] on this process group: forward + 37-head CTC + backward (+ overlapped NCCL gradient all-reduce) +
    global-norm clip + Adam + warm-up LR.  EVERY rank runs the SAME per-rank batch (8 utterances, U[3 s, 15 s], seed 3), so
    the per-rank work does not depend on N and the weak-scaling curve isolates the exchange step.  Returns the record on
    rank 0 (None elsewhere): device-timed step, end-to-end step, launches, the GEMM roofline and — for N > 1 — the all-reduce
    bytes and its EXPOSED time (step with the reducer minus the same step without it, both max over ranks)."""
    from allophant_b200 import ops, optim
    from allophant_b200.dataset_processing import Batch
    from allophant_b200.distributed import GradientReducer, attach_gradient_reducer, global_label_count
    from allophant_b200.loss_functions import multi_head_ctc_loss

    world, rank, local_rank, device, distributed = process
    if distributed:
        import torch.distributed as dist

    estimator, allophones = build_training_estimator(device)
    model = estimator.model
    # train() mode: HF dropout 0.1 / attention dropout 0.1 / LayerDrop 0.1 / SpecAugment 0.075 are
    # applied by the CUDA path with counter-based masks; --eval-arithmetic times the deterministic arithmetic instead
    model.train(not eval_arithmetic)
    wire = torch.bfloat16 if os.environ.get("BENCH_ALLREDUCE_BF16") == "1" else None  # experiment: gradients as bf16 on the wire
    reducer = GradientReducer(wire_dtype=wire) if distributed else None
    attach_gradient_reducer(model, reducer)

    # variable-length batch: U[3 s, 15 s], sorted, zero padded to the longest utterance; the same on every rank
    generator = torch.Generator().manual_seed(3)
    seconds = 3.0 + 12.0 * torch.rand(TRAIN_BATCH, generator=generator)
    lengths = (seconds * SAMPLE_RATE).long().sort(descending=True).values
    samples = int(lengths.max())
    audio = 0.1 * torch.randn(TRAIN_BATCH, samples, generator=generator)
    audio = audio * (torch.arange(samples)[None, :] < lengths[:, None])
    languages = torch.randint(0, TRAIN_LANGUAGES, (TRAIN_BATCH,), generator=generator)
    host_audio, host_lengths, host_languages = audio.pin_memory(), lengths.pin_memory(), languages.pin_memory()

    frames = model.downsampled_lengths(lengths)
    head_classes = {name: 4 for name in model.classes if name != "phoneme"}
    head_classes["phoneme"] = TRAIN_PHONES + 1
    names = [name for name in model.classes]
    labels_host, label_lengths_host = {}, {}
    for name in names:
        head_lengths = (frames.double() * 0.25).floor().long()
        head_labels = torch.zeros(TRAIN_BATCH, int(head_lengths.max()), dtype=torch.long)
        for row, length in enumerate(head_lengths.tolist()):
            if name == "phoneme":  # labels from the utterance's own language inventory (absent phonemes are fully masked)
                inventory = torch.tensor(sorted(allophones[int(languages[row])]), dtype=torch.long) + 1
                head_labels[row, :length] = inventory[torch.randint(0, len(inventory), (length,), generator=generator)]
            else:
                head_labels[row, :length] = torch.randint(1, head_classes[name], (length,), generator=generator)
        labels_host[name], label_lengths_host[name] = head_labels.pin_memory(), head_lengths.pin_memory()
    resident = Batch(host_audio.to(device), host_lengths.to(device), host_languages.to(device))
    labels_dev = {name: value.to(device) for name, value in labels_host.items()}
    label_lengths_dev = {name: value.to(device) for name, value in label_lengths_host.items()}
    parameters = list(model.parameters())  # all of them, as the reference hands them to the optimizer (estimator.py:982)

    # default_config.toml:107-121: Adam(0.9, 0.98), lr 1e-3 under the warm-up schedule; global-norm clipping folded into the Adam pass
    optimizer = optim.adam_from_config(parameters, model.d_model, model=model)

    def step(batch, labels, label_lengths):
        for parameter in parameters:
            parameter.grad = None
        predictions = model(batch)
        predictions.outputs.pop("phone", None)
        order = list(predictions.outputs)
        losses = multi_head_ctc_loss(
            [predictions.outputs[name] for name in order], [labels[name] for name in order], predictions.lengths, [label_lengths[name] for name in order]
        )
        count = global_label_count([label_lengths[name] for name in order])  # all-reduced scalar: the loss normaliser of the WHOLE batch
        loss = losses.sum() / count.to(losses.dtype)
        loss.backward()
        optimizer.step(clip_norm=1.0)
        return loss

    def device_step():
        return step(resident, labels_dev, label_lengths_dev)

    def e2e_step():
        batch = Batch(host_audio, host_lengths, host_languages).to(device, non_blocking=True)
        labels = {name: value.to(device, non_blocking=True) for name, value in labels_host.items()}
        label_lengths = {name: value.to(device, non_blocking=True) for name, value in label_lengths_host.items()}
        return float(step(batch, labels, label_lengths).item())

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(count: int) -> float:
        """milliseconds for `count` steps: barrier + synchronize on both sides, device events, max over ranks"""
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(count):
            device_step()
        end.record()
        barrier()
        elapsed = torch.tensor([start.elapsed_time(end)], device=device)
        if distributed:
            dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        return float(elapsed.item())

    for _ in range(max(3, warmup)):
        loss = device_step()
    if not torch.isfinite(loss):
        raise RuntimeError(f"training loss is not finite: {float(loss)}")
    ops.reset_launch_count()
    collectives_before = (reducer.issued, reducer.bytes) if reducer is not None else (0, 0)
    with ClockSampler(local_rank) as sampler:
        elapsed_ms = timed_steps(steps)
    collectives = None
    if reducer is not None:
        collectives = {
            "collectives_per_step": (reducer.issued - collectives_before[0]) // steps,
            "bytes_per_step": (reducer.bytes - collectives_before[1]) // steps,
            "wire_dtype": "bf16" if wire is not None else "f32",
        }
    launches = ops.launch_count() // steps
    clocks = sampler.summary()
    if reducer is not None:
        # the same steps WITHOUT the exchange (every rank keeps its local gradients): the difference is what the all-reduce
        # adds to the step, i.e. the part of it that the backward pass does not hide
        attach_gradient_reducer(model, None)
        device_step()
        local_ms = timed_steps(steps)
        attach_gradient_reducer(model, reducer)
        device_step()
        collectives["step_ms_without_allreduce"] = local_ms / steps
        collectives["exposed_ms_per_step"] = (elapsed_ms - local_ms) / steps

    for _ in range(2):
        e2e_step()
    barrier()
    wall_start = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_seconds = torch.tensor([time.perf_counter() - wall_start], device=device)
    if distributed:
        dist.all_reduce(e2e_seconds, op=dist.ReduceOp.MAX)
    audio_seconds = world * float(lengths.sum()) / SAMPLE_RATE * steps
    h2d_bytes = host_audio.numel() * 4 + 16 * TRAIN_BATCH + sum(v.numel() * 8 for v in labels_host.values()) + sum(
        v.numel() * 8 for v in label_lengths_host.values()
    )

    # ---- roofline of the dominant kernel: every rank runs one more step (it contains collectives), rank 0 times its GEMMs
    roofline = None
    if detail:
        # kernels are timed one at a time: the weight-gradient stream is off for these two extra steps (records of overlapping
        # kernels include the time they share the SMs, so their sum would exceed the step)
        from allophant_b200 import engine as _engine

        overlap_before = _engine.set_backward_overlap(False)
        gemm_events: List[Any] = []
        original = ops.run_gemm
        sizes = (1024, 3072, 4096)

        def is_encoder_linear(gemm_args) -> bool:
            return gemm_args.mode == 0 and gemm_args.n in sizes and (
                (not gemm_args.b_mn_major and gemm_args.k in sizes) or (gemm_args.b_mn_major and not gemm_args.a_mn_major and gemm_args.k_seq in sizes)
                or (gemm_args.a_mn_major and gemm_args.a_rows in sizes)
            )

        def timed_gemm(gemm_args):
            if is_encoder_linear(gemm_args):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                original(gemm_args)
                ev1.record()
                gemm_events.append((ev0, ev1))
            else:
                original(gemm_args)

        if rank == 0:
            ops.run_gemm = timed_gemm
        try:
            device_step()
            torch.cuda.synchronize()
        finally:
            ops.run_gemm = original
        if rank == 0:
            plan_frames = int(frames.max())
            gemm_ms = sum(a.elapsed_time(b) for a, b in gemm_events)
            flops = 3.0 * encoder_flops(plan_frames)["linear"] * TRAIN_BATCH  # forward + dgrad + wgrad over the padded frame count
            skipped = list(model._heads.last_regularisation["plan"].skipped)  # LayerDrop decisions of this very step
            flops *= (len(skipped) - sum(skipped)) / max(1, len(skipped))
            peaks = measured_peaks()
            achieved = flops / (gemm_ms / 1000.0) / 1e12
            peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
            roofline = {
                "kernel": "aph::gemm_bf16_kernel (encoder linears: forward, data-gradient and weight-gradient forms)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"bf16_tflops_sustained, {peaks['source']}", "launches": len(gemm_events),
                "avg_launch_ms": gemm_ms / max(1, len(gemm_events)), "share_of_step": gemm_ms / (elapsed_ms / steps),
            }  # fmt: skip
        # one more step on every rank (collectives inside): rank 0 reads the un-bracketed kernel durations from CUPTI
        if rank == 0:
            unbracketed = cupti_gemm_time(device_step, is_encoder_linear)
            if unbracketed is not None and roofline is not None:
                count, total_ms = unbracketed
                skipped = list(model._heads.last_regularisation["plan"].skipped)
                step_flops = 3.0 * encoder_flops(int(frames.max()))["linear"] * TRAIN_BATCH * (len(skipped) - sum(skipped)) / max(1, len(skipped))
                roofline["kernel_only"] = {
                    "source": "CUPTI kernel records of one un-bracketed step (torch.profiler), the same launches as the event brackets",
                    "launches": count, "avg_launch_ms": total_ms / max(1, count), "achieved": step_flops / (total_ms / 1000.0) / 1e12,
                    "frac": step_flops / (total_ms / 1000.0) / 1e12 / roofline["peak"], "share_of_step": total_ms / (elapsed_ms / steps),
                }  # fmt: skip
        else:
            device_step()
            torch.cuda.synchronize()
        _engine.set_backward_overlap(overlap_before)
        if roofline is not None:
            roofline["note"] = (
                "GEMM launches timed with the weight-gradient stream off (APH_BWD_OVERLAP): one kernel at a time; `share_of_step` relates "
                "their sum to the timed step, which runs with the stream on"
            )
    attach_gradient_reducer(model, None)
    del optimizer, estimator, model
    torch.cuda.empty_cache()
    if distributed:
        dist.barrier()
    if rank != 0:
        return None
    return {
        "metric": "audio-sec/sec, multitask training step (forward + multi-head CTC + backward + gradient all-reduce + optimizer)",
        "value": audio_seconds / (elapsed_ms / 1000.0),
        "unit": UNIT,
        "n_gpus": world,
        "steps": steps,
        "warmup": max(3, warmup),
        "ms_per_step": elapsed_ms / steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[2]: Allophant Multitask (XLS-R-300M shape, random init, allophone layer over "
            f"{TRAIN_LANGUAGES} languages x {TRAIN_PHONES} phones, feature extractor frozen) training step: forward + 37-head CTC + backward"
            f"{' + overlapped NCCL gradient all-reduce' if distributed else ''} + global-norm clip + Adam + warm-up LR; {TRAIN_BATCH} utterances per GPU "
            "(the same batch on every rank), U[3 s, 15 s], " + ("eval()-mode arithmetic (no dropout / LayerDrop / SpecAugment)" if eval_arithmetic else
                                "train() mode: hidden/attention/feature-projection dropout 0.1, LayerDrop 0.1, SpecAugment 0.075 x 10 frames"),
            "batch_per_gpu": TRAIN_BATCH,
            "audio_seconds_per_gpu": float(lengths.sum()) / SAMPLE_RATE,
            "padded_seconds": samples / SAMPLE_RATE,
            "parallelism": f"dp{world}",
            "allreduce": collectives,
            "l2": "per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush between iterations",
        },
        "clocks": clocks,
        "e2e": {"value": audio_seconds / float(e2e_seconds.item()), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": None,
    }


def run_train_arm(args) -> None:
    process = setup_process()
    line = measure_training(process, args.steps, args.warmup, args.eval_arithmetic)
    if process[4]:
        import torch.distributed as dist

        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main() -> None:
    global BATCH, SECONDS, INVENTORY, HIERARCHICAL
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=["b200", "reference"], default="b200")
    parser.add_argument("--skip-cpu-baseline", action="store_true")
    parser.add_argument("--skip-train", action="store_true", help="predict workload: leave out the `train` sub-record (configs[2] step on the same ranks)")
    parser.add_argument("--skip-ragged", action="store_true", help="predict workload: leave out the `ragged` sub-record (MaxFrameBatchSampler stream, N = 1)")
    parser.add_argument("--skip-membound", action="store_true", help="predict workload: leave out the `membound` sub-record (HBM-bound kernels, N = 1)")
    parser.add_argument("--train-steps", type=int, default=5, help="timed steps of the `train` sub-record")
    parser.add_argument("--eval-arithmetic", action="store_true", help="train workload: eval()-mode arithmetic (no dropout / LayerDrop / SpecAugment)")
    parser.add_argument("--batch", type=int, default=BATCH, help="utterances per GPU (predict workload)")
    parser.add_argument("--seconds", type=int, default=SECONDS, help="seconds per utterance (predict workload)")
    parser.add_argument("--inventory", type=int, default=INVENTORY, help="phonemes of the target inventory (composed phoneme head)")
    parser.add_argument("--hierarchical", action="store_true", help="BASELINE configs[3]: hierarchical phoneme head")
    parser.add_argument("--workload", choices=["predict", "train"], default="predict",
                        help="predict = BASELINE configs[1] (the headline, default); train = configs[2] training step")  # fmt: skip
    args = parser.parse_args()
    BATCH, SECONDS, INVENTORY, HIERARCHICAL = args.batch, args.seconds, args.inventory, args.hierarchical
    hang_dump = float(os.environ.get("BENCH_HANG_DUMP", "0"))
    if hang_dump > 0:  # debugging aid: dump every thread's Python stack if the run is still going after this many seconds
        import faulthandler

        faulthandler.dump_traceback_later(hang_dump, exit=True)
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "train":
        run_train_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
