"""TEST INFRASTRUCTURE — restatement of the reference trainer's gradient balancing (`balance_loss`, after
https://arxiv.org/abs/1903.00277; trainer/hw_with_style_trainer.py:340-377): every stashed gradient set R_k is added to the
gradient D left by the last backward, per parameter tensor, rescaled to D's mean magnitude and weighted by x_k:

    p.grad += x_k * R_k * (mean|D| / mean|R_k|)          for every set k and every tensor with R_k present, mean|R_k| != 0

where a tensor whose mean|D| is exactly 0 uses the average of the non-zero mean|D| over all tensors instead (:354-359).
mean|D| is taken ONCE, before any set is added (:343-352).  SURVEY.md §8 f2: this is what a flat-buffer kernel pair
(segmented |g| sums, then a fused rescale-add) has to reproduce.  Pinned by tests/golden/trainer_balance.npz."""
import torch


def abs_means(grads):
    """mean|g| per tensor (None where there is no gradient) and the :354-359 replacement value for exact zeros."""
    means = [None if g is None else g.abs().mean() for g in grads]
    nz = [m for m in means if m is not None and m != 0]
    fill = (sum(nz) / len(nz)) if nz else None
    return means, fill


def balance(main, saved_sets, multipliers, means=None, fill=None):
    """main: list of tensors or None (D, modified in place like p.grad); saved_sets: list of such lists (R_k);
    multipliers: x_k per set (`balance_var_x`).  `means` / `fill` may be given when only a subset of the model's tensors is
    passed (they are statistics of the full tensors / of all tensors)."""
    if means is None:
        means, fill_all = abs_means(main)
        fill = fill_all if fill is None else fill
    means = [m if (m is None or m != 0 or fill is None) else fill for m in means]
    for x, saved in zip(multipliers, saved_sets):
        for i, (R, D) in enumerate(zip(saved, main)):
            if R is None:
                continue
            r = R.abs().mean()
            if r != 0:
                D += x * R * (means[i] / r)
    return main
