"""TEST INFRASTRUCTURE — training-step golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_training

For every case the reference's own ``Allophant`` (built through ``oracle/reference_shim.py`` exactly like
``oracle/make_golden.py``) runs ``model(batch)`` (predict=False, ``eval()`` so that dropout / LayerDrop /
SpecAugment are off), the reference's ``CTCWrapper`` per head and the step arithmetic of
``allophant/estimator.py:708-738`` (``loss = sum_heads ctc / sum label lengths``; ``backward()``).  The
restatement's ``OracleModel.training_step`` must agree (loss 1e-6, every gradient 1e-4 of its norm) before the
fixture is written.  Stored: loss, per-head CTC sums, labels, and a fingerprint of EVERY gradient (norm, sum, 32
samples) plus the full tensor for the small ones — full encoder gradients would be >100 MB per case.
"""
from __future__ import annotations

import os
import sys
from typing import Any, Dict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import make_golden, restatement  # noqa: E402

CASES = ["multitask_2layer", "hierarchical_2layer", "allophones_2layer"]
FULL_TENSOR_LIMIT = 4096  # elements


def run_case(name: str) -> Dict[str, Any]:
    case = make_golden.CASES[name]
    spec = make_golden.build_spec(case)
    ref, model, _ = make_golden.reference_model(spec)
    model.eval()
    lengths = torch.tensor(case["lengths"], dtype=torch.long)
    n_utt, samples = len(lengths), int(lengths.max())
    audio = restatement.synthetic_audio(n_utt, samples, seed=0) * restatement.mask_sequence(lengths)
    language_ids = torch.arange(n_utt) % max(1, len(spec.allophones or {0: 0}))
    batch = ref.batching.Batch(audio, lengths, language_ids)

    # ---- the reference's training step (estimator.py:708-738), eval()-mode arithmetic
    for parameter in model.parameters():
        parameter.grad = None
    predictions = model(batch)
    predictions.outputs.pop("phone", None)
    frames = predictions.lengths
    head_classes = {key: value.shape[-1] for key, value in predictions.outputs.items()}
    labels, label_lengths = restatement.training_labels(spec, head_classes, frames, language_ids)
    ctc = ref.loss_functions.CTCWrapper()
    total = torch.tensor(0, dtype=torch.float32)
    per_head, normaliser = {}, 0
    for key, output in predictions.outputs.items():
        head_loss = ctc(output, labels[key], frames, label_lengths[key])
        per_head[key] = float(head_loss)
        normaliser += int(label_lengths[key].sum())
        total = total + head_loss
    loss = total / normaliser
    loss.backward()
    reference_grads = {key: parameter.grad.detach().clone() for key, parameter in model.named_parameters() if parameter.grad is not None}
    assert all(np.isfinite(v) and v > 0 for v in per_head.values()), per_head

    # ---- the restatement must reproduce it
    oracle = restatement.OracleModel(spec)
    oracle_loss, oracle_heads, oracle_grads = oracle.training_step(audio, lengths, labels, label_lengths, language_ids)
    assert abs(float(oracle_loss) - float(loss)) <= 1e-6 * abs(float(loss)), (float(oracle_loss), float(loss))
    assert sorted(oracle_grads) == sorted(reference_grads), sorted(set(oracle_grads) ^ set(reference_grads))[:8]
    worst = 0.0
    for key, value in reference_grads.items():
        scale = float(value.norm())
        deviation = float((oracle_grads[key] - value).norm()) / max(scale, 1e-12)
        worst = max(worst, deviation)
        assert deviation < 1e-4, f"{key}: restatement gradient deviates by {deviation:.2e} of its norm"

    fixture = dict(
        case=name,
        case_config=case,
        lengths=lengths,
        language_ids=language_ids,
        frames=frames.clone(),
        labels=labels,
        label_lengths=label_lengths,
        loss=float(loss),
        per_head=per_head,
        normaliser=normaliser,
        gradient_summaries={key: restatement.gradient_summary(value) for key, value in reference_grads.items()},
        small_gradients={key: value for key, value in reference_grads.items() if value.numel() <= FULL_TENSOR_LIMIT},
        frozen=[key for key, parameter in model.named_parameters() if parameter.grad is None],
        restatement_max_deviation=worst,
        versions=dict(torch=torch.__version__, transformers=__import__("transformers").__version__),
    )
    print(f"[{name}] loss={float(loss):.6f} heads={len(per_head)} grads={len(reference_grads)} frozen={len(fixture['frozen'])} "
          f"restatement deviation={worst:.2e}")  # fmt: skip
    return fixture


def main() -> None:
    for name in sys.argv[1:] or CASES:
        fixture = run_case(name)
        path = os.path.join(make_golden.GOLDEN_DIR, f"training_{name}.pt")
        torch.save(fixture, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
