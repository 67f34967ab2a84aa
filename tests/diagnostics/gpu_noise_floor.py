"""GPU diagnostic: spread of the eval()-mode gradient deviation (CUDA bf16 path vs fp32 oracle) over different audio seeds
of the small golden models — the noise floor the train()-mode tolerance has to be read against."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import restatement  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_training import _training_step, norm_err  # noqa: E402


def main() -> None:
    from allophant_b200.dataset_processing import Batch

    name = sys.argv[1] if len(sys.argv) > 1 else "allophones_2layer"
    fixture = helpers.load_golden(f"training_{name}")
    spec = helpers.spec_for_case(fixture["case_config"])
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    lengths = fixture["lengths"]
    for seed in range(8):
        audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=seed) * restatement.mask_sequence(lengths)
        batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
        model.eval()
        loss, _, _ = _training_step(model, batch, fixture)
        reference_loss, _, reference = oracle.training_step(audio, lengths, fixture["labels"], fixture["label_lengths"], fixture["language_ids"])
        worst = {}
        for pname, parameter in model.named_parameters():
            if parameter.grad is None or pname not in reference or float(reference[pname].norm()) < 1e-7:
                continue
            worst[pname] = norm_err(parameter.grad, reference[pname])
        ranked = sorted(worst.items(), key=lambda item: -item[1])[:3]
        print(f"{name} audio seed {seed}: loss {float(loss):.5f} / {float(reference_loss):.5f}  "
              + ", ".join(f"{k.split('._model.')[-1].replace('_projection._layers.', '')}={v:.3e}" for k, v in ranked), flush=True)  # fmt: skip


if __name__ == "__main__":
    main()
