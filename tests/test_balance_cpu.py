"""CPU: oracle/balance.py (the reference trainer's gradient balancing, trainer/hw_with_style_trainer.py:340-377) against
what the UNMODIFIED trainer did in curriculum slots 1 -> 2 (tests/golden/trainer_balance.npz, written by
`python -m oracle.make_trainer_golden balance`): four stashed gradient sets, multipliers [0.6, 0.5, 0.4, 0.75], the
gradient before and after balancing for 40 parameter tensors across the model's parts — including tensors whose own
mean |gradient| is exactly zero (the :354-359 replacement rule)."""
import numpy as np
import torch

from oracle import balance as obal


def test_balance_oracle_matches_the_reference_trainer(golden_dir):
    gold = np.load(f"{golden_dir}/trainer_balance.npz")
    names = gold["names"].tolist()
    assert len(names) >= 30 and int(gold["n_zero_mean"]) >= 1
    mult = gold["multipliers"].tolist()
    assert mult == [0.6, 0.5, 0.4, 0.75]                     # config :100, `balance_var_x`
    main = [torch.from_numpy(gold[f"D/{n}"].copy()) for n in names]
    sets = [[torch.from_numpy(gold[f"R{k}/{n}"]) if f"R{k}/{n}" in gold.files else None for n in names] for k in range(4)]
    for n, d in zip(names, main):                            # the stored statistics are those of the full tensors
        assert abs(float(d.abs().mean()) - float(gold[f"meanD/{n}"])) <= 1e-7 * max(1.0, float(gold[f"meanD/{n}"]))
    out = obal.balance(main, sets, mult, fill=torch.tensor(float(gold["fill"])))
    used_fill = 0
    for n, got in zip(names, out):
        ref = torch.from_numpy(gold[f"after/{n}"])
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) <= 1e-6 * scale + 1e-12, n
        used_fill += int(float(gold[f"meanD/{n}"]) == 0 and scale > 0)
    assert used_fill >= 1                                    # the zero-mean rule was exercised and changed a gradient


def test_balance_full_statistics():
    """abs_means on a whole parameter list: None entries, exact zeros, and the fill value."""
    g = [torch.tensor([1.0, -3.0]), None, torch.zeros(3), torch.tensor([0.5])]
    means, fill = obal.abs_means(g)
    assert means[1] is None and float(means[0]) == 2.0 and float(means[2]) == 0.0 and abs(float(fill) - 1.25) < 1e-7
    saved = [[torch.tensor([2.0, 2.0]), None, torch.ones(3), None]]
    out = obal.balance([t if t is None else t.clone() for t in g], saved, [0.5])
    assert torch.allclose(out[0], torch.tensor([1.0, -3.0]) + 0.5 * torch.tensor([2.0, 2.0]) * (2.0 / 2.0))
    assert torch.allclose(out[2], 0.5 * torch.ones(3) * 1.25)       # zero-mean tensor takes the fill value
    assert torch.equal(out[3], g[3])
