"""GPU diagnostic: gradient deviation of one train()-mode step against the oracle with ONE regularisation site enabled at
a time (same masks on both sides).  A site whose deviation stands out has a mask or arithmetic mismatch; uniform
deviations at the eval()-mode level are bf16 noise."""
import dataclasses
import sys

import torch

sys.path.insert(0, ".")
from oracle import restatement  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_training import _training_step, norm_err  # noqa: E402


def main() -> None:
    from allophant_b200 import engine
    from allophant_b200.dataset_processing import Batch

    name = sys.argv[1] if len(sys.argv) > 1 else "allophones_2layer"
    fixture = helpers.load_golden(f"training_{name}")
    spec = helpers.spec_for_case(fixture["case_config"])
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0) * restatement.mask_sequence(lengths)
    batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
    sites = {
        "none": {},
        "hidden": dict(hidden_dropout=0.1),
        "attention": dict(attention_dropout=0.1),
        "feat_proj": dict(feat_proj_dropout=0.1),
        "spec": dict(mask_time_prob=0.075),
        "all": dict(hidden_dropout=0.1, attention_dropout=0.1, feat_proj_dropout=0.1, mask_time_prob=0.075),
    }
    original = engine.Stochastic.from_config
    rate = model._projection._acoustic_model_dropout
    for label, overrides in list(sites.items()) + [("input", {})]:
        engine.Stochastic.from_config = classmethod(lambda cls, cfg, seed, o=overrides: cls(seed, **o))
        rate.p = 0.2 if label in ("input", "all") else 0.0
        for seed in (31, 32):
            model.train()
            model._heads.skip_layers_override = [False, False]
            torch.manual_seed(seed)
            loss, _, _ = _training_step(model, batch, fixture)
            state = model._heads.last_regularisation
            model.eval()
            stochastic, plan = state["stochastic"], state["plan"]
            cfg = plan.cfg
            masks = helpers.regularisation_masks(
                stochastic, plan.n_utt, plan.seq, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, plan.skipped,
                plan.spec_mask.cpu() if plan.spec_active else None,
            )  # fmt: skip
            blocks = {0: cfg.num_hidden_layers, **{column: index for index, column in model._heads.hidden_blocks.items()}}
            masks["classifier_input"] = {
                blocks[column]: helpers.keep_mask(drop, plan.rows, cfg.hidden_size).view(plan.n_utt, plan.seq, cfg.hidden_size)
                for column, drop in state["input_dropout"].items()
            }
            reference_loss, _, reference = oracle.training_step(
                audio, lengths, fixture["labels"], fixture["label_lengths"], fixture["language_ids"], regularisation=masks
            )
            worst = {}
            for pname, parameter in model.named_parameters():
                if parameter.grad is None or pname not in reference or float(reference[pname].norm()) < 1e-7:
                    continue
                worst[pname] = norm_err(parameter.grad, reference[pname])
            ranked = sorted(worst.items(), key=lambda item: -item[1])[:3]
            print(f"{name} {label:10s} seed {seed}: loss {float(loss):.5f} / {float(reference_loss):.5f}  "
                  + ", ".join(f"{k.split('._model.')[-1].replace('_projection._layers.', '')}={v:.3e}" for k, v in ranked), flush=True)  # fmt: skip
    engine.Stochastic.from_config = original


if __name__ == "__main__":
    main()
