"""TEST INFRASTRUCTURE — deterministic synthetic inputs shared by the golden
generator, the tests and bench.py (numpy RandomState: identical on every box)."""
import numpy as np


from bench_inputs import ctc_case, gen_case, gen_noise, gen_noise_shapes, hwr_case  # noqa: E402,F401  (input builders, numpy only)


def state_dict_from_seed(make_module, seed):
    """Random-init weights, reproducible from the seed (torch CPU generator)."""
    import torch
    torch.manual_seed(seed)
    m = make_module()
    # BatchNorm affine/running stats and biases away from their trivial init, so parity is not vacuous
    g = torch.Generator().manual_seed(seed + 1)
    sd = m.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif "batchnorm" in k or (k.startswith("cnn1d") and k.split(".")[1] in ("1", "4", "7", "10")):
            if k.endswith("weight"):
                v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
            elif k.endswith("bias"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif "noise" in k and k.endswith("weight_orig"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.25)   # make the noise path matter
        elif k.endswith("conv1.0.bias") or k.endswith("out.0.conv.bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return m, sd


DISC_SITES = (("convs1.3", 0.05, 2), ("convs3.4", 0.05, 4), ("convs4.0", 0.025, 2), ("convs4.4", 0.025, 4),
              ("convs4.7", 0.025, 4), ("convs4.11", 0.025, 4))   # Dropout2d site, p, channels / dim


def disc_masks(B, seed, dim=64, p_scale=4.0):
    """Keep-masks [B,C] of the discriminator's Dropout2d layers (model/discriminator_ap.py:89,105,116-127).
    p_scale inflates the drop probability of the MASK DRAW (not the 1/(1-p) rescale) so that small batches really
    contain dropped channels."""
    r = np.random.RandomState(seed)
    return {site: (r.rand(B, c * dim) >= p * p_scale).astype(np.float32) for site, p, c in DISC_SITES}


def perturb_disc(sd, seed):
    """GroupNorm affine parameters away from (1, 0), in place, so that parity is not vacuous."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for k in ("in_conv.1", "convs3.1"):
        sd[k + ".weight"].copy_(torch.rand(sd[k + ".weight"].shape, generator=g) * 0.5 + 0.75)
        sd[k + ".bias"].copy_(torch.randn(sd[k + ".bias"].shape, generator=g) * 0.1)
    return sd


def random_state_dict(net, seed, n_class=80, golden_dir=None):
    """Random-init weights of one of the path's networks ('gen', 'hwr', 'disc', 'enc') built from the key/shape fixture
    that oracle/make_golden.py took from the UNMODIFIED reference's `state_dict()` (tests/golden/<net>.npz:
    state_dict_keys) — so the baseline arms of bench.py need neither /root/reference nor the product's modules.
    Magnitudes follow the reference's initialisers in spirit (fan-in scaled weights, unit norm scales, N(0,1) EqualLR /
    FusedUpsample weights, unit spectral-norm vectors); the values only have to keep the step finite: it is timed, not
    compared."""
    import os
    import torch
    if golden_dir is None:
        golden_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    keys = np.load(os.path.join(golden_dir, f"{net}.npz"))["state_dict_keys"]
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for item in keys:
        k, shp = str(item).split(":")
        shape = tuple(int(d) for d in shp.split("x")) if shp else ()
        if net == "hwr" and k.startswith("cnn1d.12."):
            shape = (n_class,) + shape[1:]
        if net == "gen" and k == "conv.0.conv1.weight":
            shape = (n_class + 128,) + shape[1:]
        if k.endswith("num_batches_tracked"):
            v = torch.zeros(shape, dtype=torch.int64)
        elif k.endswith("running_var"):
            v = torch.ones(shape)
        elif k.endswith("running_mean"):
            v = torch.zeros(shape)
        elif k.endswith(("weight_u", "weight_v")):
            v = torch.randn(shape, generator=g)
            v = v / v.norm().clamp_min(1e-12)
        elif net == "gen" and len(shape) == 4 and shape[1:] == (1, 3, 3) and shape[0] > 1:
            v = (torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16).expand(shape).clone()   # Blur buffers
        elif k.endswith("weight_orig") and "noise" in k:
            v = torch.full(shape, 0.01)
        elif k.endswith("weight_orig") or (net == "gen" and k.endswith("conv1.0.weight")):
            v = torch.randn(shape, generator=g)                      # EqualLR / FusedUpsample: N(0,1), scaled at run time
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            v = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif k.endswith("weight"):
            v = torch.ones(shape)                                    # normalisation scales
        elif "adain" in k and k.endswith("style.bias"):
            v = torch.cat((torch.ones(shape[0] // 2), torch.zeros(shape[0] // 2)))     # gamma 1, beta 0 (pure_gen.py:59-60)
        else:
            v = torch.zeros(shape)
        sd[k] = v
    return sd
