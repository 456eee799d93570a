"""Thin python entry points of the fused memory-bound kernels (hwg_fused.cu)."""
import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_LRELU  # noqa: F401


def linear(x, W, b, act=ACT_NONE, slope=0.0):
    B, K = x.shape
    O = W.size(0)
    out = torch.empty((B, O), device=x.device, dtype=torch.float32)
    _lib.call("hwg_linear_f32", x.data_ptr(), W.data_ptr(), _lib.ptr(b), out.data_ptr(), B, K, O, act, slope,
              _lib.stream())
    return out


def linear_bwd(x, y, gy, W, act=ACT_NONE, slope=0.0, need_gx=True, gW=None, gb=None, accumulate=False):
    """Backward of y = act(x W^T + b) (fp32).  Returns (gx or None, gW, gb); gW/gb may be caller-provided buffers
    (with accumulate=True they are added to — e.g. views of a flat gradient buffer)."""
    B, K = x.shape
    O = W.size(0)
    gx = torch.zeros((B, K), device=x.device, dtype=torch.float32) if need_gx else None    # the kernel adds slices
    if gW is None:
        gW = torch.empty((O, K), device=x.device, dtype=torch.float32)
        gb = torch.empty(O, device=x.device, dtype=torch.float32)
        accumulate = False
    _lib.call("hwg_linear_bwd_f32", x.data_ptr(), _lib.ptr(y), gy.data_ptr(), W.data_ptr(), B, K, O, act, slope,
              _lib.ptr(gx), gW.data_ptr(), gb.data_ptr(), int(accumulate), _lib.stream())
    return gx, gW, gb


def pixelnorm(x):
    out = torch.empty_like(x)
    _lib.call("hwg_pixelnorm_f32", x.data_ptr(), out.data_ptr(), x.size(0), x.size(1), _lib.stream())
    return out


def gen_pack_input(content, style, Cp):
    """content [T,B,C] fp32 (any strides), style [B,S] fp32 -> [B,1,T,Cp] bf16."""
    T, B, C = content.shape
    S = 0 if style is None else style.size(1)
    x = torch.empty((B, 1, T, Cp), device=content.device, dtype=torch.bfloat16)
    _lib.call("hwg_gen_pack_input", content.data_ptr(), content.stride(0), content.stride(1), content.stride(2),
              _lib.ptr(style), T, B, C, S, Cp, x.data_ptr(), _lib.stream())
    return x


def adain_coeffs(stats, gamma, beta, gb_stride, N, C, HW, eps=1e-5, save=False):
    coef = torch.empty((N, C, 2), device=stats.device, dtype=torch.float32)
    sv = torch.empty((N, C, 2), device=stats.device, dtype=torch.float32) if save else None
    _lib.call("hwg_adain_coeffs", stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), gb_stride, N, C, HW, eps,
              coef.data_ptr(), _lib.ptr(sv), _lib.stream())
    return (coef, sv) if save else coef


def bn_coeffs(stats, N, C, count_per_n, weight, bias, running_mean, running_var, momentum, eps, use_batch_stats):
    dev = weight.device
    coef = torch.empty((C, 2), device=dev, dtype=torch.float32)
    save = torch.empty((C, 2), device=dev, dtype=torch.float32)
    _lib.call("hwg_bn_coeffs", _lib.ptr(stats), N, C, count_per_n, _lib.ptr(weight), _lib.ptr(bias),
              _lib.ptr(running_mean), _lib.ptr(running_var), momentum, eps, int(use_batch_stats), coef.data_ptr(),
              save.data_ptr(), _lib.stream())
    return coef, save


def scale_shift_act(x, coef, per_sample, act=ACT_NONE, slope=0.0, out=None):
    """x [N,H,W,C] bf16 -> act(a*x+b); in place when out is None."""
    N, H, W, C = x.shape
    out = x if out is None else out
    _lib.call("hwg_scale_shift_act", x.data_ptr(), out.data_ptr(), coef.data_ptr(), int(per_sample), N, H * W, C, act,
              slope, _lib.stream())
    return out


def blur_noise_act_stats(x, noise, noise_w, stats, act=ACT_LRELU, slope=0.2, seed=0, subseq=0, seed_dev=None):
    N, H, W, C = x.shape
    y = torch.empty_like(x)
    _lib.call("hwg_blur_noise_act_stats", x.data_ptr(), y.data_ptr(), N, H, W, C, _lib.ptr(noise), _lib.ptr(noise_w),
              seed, subseq, _lib.ptr(seed_dev), act, slope, _lib.ptr(stats), _lib.stream())
    return y


def gen_output(x, coef, w, b0):
    """b0: 1-element fp32 CUDA tensor (read on the device: no host sync, CUDA-graph friendly)."""
    N, H, W, C = x.shape
    out = torch.empty((N, 1, H, W), device=x.device, dtype=torch.float32)
    _lib.call("hwg_gen_output", x.data_ptr(), coef.data_ptr(), w.data_ptr(), b0.data_ptr(), N, H * W, C, out.data_ptr(),
              _lib.stream())
    return out


def hwr_stem(img, w, b):
    N, _, H, W = img.shape
    Cout = w.size(0)
    y = torch.empty((N, H // 2, W // 2, Cout), device=img.device, dtype=torch.bfloat16)
    _lib.call("hwg_hwr_stem", img.data_ptr(), w.data_ptr(), b.data_ptr(), N, H, W, Cout, y.data_ptr(), _lib.stream())
    return y


def maxpool_nhwc(x, k, s, p):
    N, H, W, C = x.shape
    Ho = (H + 2 * p[0] - k[0]) // s[0] + 1
    Wo = (W + 2 * p[1] - k[1]) // s[1] + 1
    y = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.bfloat16)
    _lib.call("hwg_maxpool_nhwc", x.data_ptr(), y.data_ptr(), N, H, W, C, k[0], k[1], s[0], s[1], p[0], p[1], Ho, Wo,
              _lib.stream())
    return y


# ---- backward passes (hwg_fused_bwd.cu) ----------------------------------------------------------------
class ZeroArena:
    """Pre-zeroed fp32 accumulators for one pass, carved out of ONE zero-filled allocation (one memset instead of one
    fill launch per accumulator)."""

    def __init__(self, numel, device):
        self.buf = torch.zeros(numel, device=device, dtype=torch.float32)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.buf.numel():      # sized too small by the caller: fall back to a fresh fill
            return torch.zeros(shape, device=self.buf.device, dtype=torch.float32)
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += -(-n // 4) * 4
        return v


def _zeros(arena, shape, device):
    return arena.take(*shape) if arena is not None else torch.zeros(shape, device=device, dtype=torch.float32)


def logsoftmax_bwd(g, lp, Cp, arena=None):
    """g, lp [T,B,C] fp32 -> (gz [B,1,T,Cp] bf16 NHWC, dbias [C] fp32)."""
    T, B, C = lp.shape
    gz = torch.empty((B, 1, T, Cp), device=lp.device, dtype=torch.bfloat16)
    db = _zeros(arena, (C,), lp.device)
    _lib.call("hwg_logsoftmax_bwd", g.data_ptr(), lp.data_ptr(), T, B, C, Cp, gz.data_ptr(), db.data_ptr(), _lib.stream())
    return gz, db


def _is_peer(sync):
    return hasattr(sync, "allreduce_")          # dp.PeerExchange (in-kernel NVLink exchange) vs a torch process group


def bn_bwd(g, z, coef, save, weight, relu=True, arena=None, sync_group=None, sync_key=None, use_batch=True):
    """BatchNorm(+ReLU) backward on NHWC bf16: returns (gz bf16, dweight [C], dbias [C], dconv_bias [C]);
    dweight/dbias are strided views of the [C,2] sums.
    sync_group: what the forward normalised JOINT batch statistics with (SyncBN) — a dp.PeerExchange (one
    hwg_peer_allreduce_f32 launch, slot `sync_key`) or a torch.distributed group (NCCL all-reduce): the
    (sum gy, sum gy*xhat) pair that enters the input gradient is summed over the ranks and divided by the global row
    count; the parameter gradients stay the local sums (the gradient all-reduce averages those).
    use_batch=False: the forward normalised with the RUNNING statistics (eval mode under autograd) — mean and variance
    are constants, so the input gradient is sc*gy: the apply pass gets zeroed sums and no exchange runs."""
    C = z.size(-1)
    rows = z.numel() // C
    sums = _zeros(arena, (C, 2), z.device)
    _lib.call("hwg_bn_bwd_reduce", g.data_ptr(), z.data_ptr(), coef.data_ptr(), save.data_ptr(), rows, C, int(relu),
              sums.data_ptr(), _lib.stream())
    gsums, grows = sums, rows
    if not use_batch:
        gsums = _zeros(arena, (C, 2), z.device)
    elif sync_group is not None:
        gsums = torch.empty_like(sums)
        if _is_peer(sync_group):
            _lib.call("hwg_peer_allreduce_f32", sums.data_ptr(), gsums.data_ptr(), 2 * C, *sync_group.args(sync_key),
                      _lib.stream())
            grows = rows * sync_group.world
        else:
            import torch.distributed as dist
            gsums.copy_(sums)
            dist.all_reduce(gsums, group=sync_group)
            grows = rows * dist.get_world_size(sync_group)
    gz = torch.empty_like(z)
    dcb = _zeros(arena, (C,), z.device)
    _lib.call("hwg_bn_bwd_apply", g.data_ptr(), z.data_ptr(), coef.data_ptr(), save.data_ptr(), weight.data_ptr(),
              gsums.data_ptr(), rows, grows, C, int(relu), gz.data_ptr(), dcb.data_ptr(), _lib.stream())
    return gz, sums[:, 1], sums[:, 0], dcb


def bn_coeffs_synced(stats, N, C, count_per_n, weight, bias, running_mean, running_var, momentum, eps, group, key=None):
    """bn_coeffs over the JOINT batch of the ranks (SyncBN; SURVEY 8e coupling 1), equal shards per rank.
    group = dp.PeerExchange: ONE launch folds the local per-(n,c) sums, adds them over the ranks through peer memory
    and writes coefficients / running statistics (hwg_bn_coeffs_peer).  group = torch.distributed group: fold,
    NCCL all-reduce of [C,2] floats, hwg_bn_coeffs with the global element count."""
    if _is_peer(group):
        dev = weight.device
        coef = torch.empty((C, 2), device=dev, dtype=torch.float32)
        save = torch.empty((C, 2), device=dev, dtype=torch.float32)
        _lib.call("hwg_bn_coeffs_peer", stats.data_ptr(), N, C, count_per_n * N * group.world, _lib.ptr(weight),
                  _lib.ptr(bias), _lib.ptr(running_mean), _lib.ptr(running_var), momentum, eps, coef.data_ptr(),
                  save.data_ptr(), *group.args(key), _lib.stream())
        return coef, save
    import torch.distributed as dist
    tot = stats.view(N, C, 2).sum(0, keepdim=True).contiguous()
    dist.all_reduce(tot, group=group)
    return bn_coeffs(tot, 1, C, count_per_n * N * dist.get_world_size(group), weight, bias, running_mean, running_var,
                     momentum, eps, True)


def relu_maxpool_bwd(ga, c, k, s, p, arena=None):
    """ga [N,Ho,Wo,C], c [N,H,W,C] (post-ReLU, pre-pool) -> (gc [N,H,W,C] bf16, dbias [C])."""
    N, H, W, C = c.shape
    Ho, Wo = ga.size(1), ga.size(2)
    gc = torch.empty_like(c)
    db = _zeros(arena, (C,), c.device)
    _lib.call("hwg_relu_maxpool_bwd", ga.data_ptr(), c.data_ptr(), N, H, W, C, k[0], k[1], s[0], s[1], p[0], p[1], Ho, Wo,
              gc.data_ptr(), db.data_ptr(), _lib.stream())
    return gc, db


def hwr_stem_bwd(img, w, b, ga, arena=None):
    N, _, H, W = img.shape
    Cout = w.size(0)
    dw = _zeros(arena, (Cout, 9), img.device)
    db = _zeros(arena, (Cout,), img.device)
    _lib.call("hwg_hwr_stem_bwd", img.data_ptr(), w.data_ptr(), b.data_ptr(), ga.data_ptr(), N, H, W, Cout, dw.data_ptr(),
              db.data_ptr(), _lib.stream())
    return dw, db


def adain_lrelu_bwd(g, a, save, coef, slope=0.2, noise=None, seed=0, subseq=0, row_subseq=False, seed_dev=None,
                    sums=None, dch=None):
    """Backward of x_next = AdaIN(LeakyReLU(y)), y = pre + nw*z.  g, a [N,H,W,C] bf16.
    Returns (gy bf16, dgamma [N,C], dbeta [N,C], dbias [C] = sum gy, dnoise_w [C] = sum gy*z).
    sums [N,C,2] / dch [C,2]: optional pre-zeroed fp32 accumulators (slices of a workspace)."""
    N, H, W, C = a.shape
    if sums is None:
        sums = torch.zeros((N, C, 2), device=a.device, dtype=torch.float32)
    _lib.call("hwg_adain_bwd_reduce", g.data_ptr(), a.data_ptr(), save.data_ptr(), N, H * W, C, sums.data_ptr(),
              _lib.stream())
    gy = torch.empty_like(a)
    if dch is None:
        dch = torch.zeros((C, 2), device=a.device, dtype=torch.float32)
    _lib.call("hwg_adain_bwd_apply", g.data_ptr(), a.data_ptr(), save.data_ptr(), coef.data_ptr(), sums.data_ptr(),
              N, H, W, C, slope, _lib.ptr(noise), seed, subseq, _lib.ptr(seed_dev), int(row_subseq), gy.data_ptr(),
              dch.data_ptr(),
              _lib.stream())
    return gy, sums[:, :, 1], sums[:, :, 0], dch[:, 0], dch[:, 1]


def gen_output_bwd(g_out, out, a, coef, w, dwb=None):
    """Returns (gx [N,H,W,C] bf16, dw [C], db0 []).  dwb: optional pre-zeroed fp32 [C+1] accumulator."""
    N, H, W, C = a.shape
    gx = torch.empty_like(a)
    if dwb is None:
        dwb = torch.zeros(C + 1, device=a.device, dtype=torch.float32)
    _lib.call("hwg_gen_output_bwd", g_out.data_ptr(), out.data_ptr(), a.data_ptr(), coef.data_ptr(), w.data_ptr(),
              N, H * W, C, gx.data_ptr(), dwb.data_ptr(), _lib.stream())
    return gx, dwb[:C], dwb[C]


def hwr_stem_bwd_image(img, w, b, ga):
    """ga [N,H/2,W/2,64] bf16 -> gradient w.r.t. the image [N,1,H,W] fp32 (conv0 + ReLU + MaxPool backward, fused)."""
    N, _, H, W = img.shape
    g = torch.zeros((N, 1, H, W), device=img.device, dtype=torch.float32)
    _lib.call("hwg_hwr_stem_bwd_image", img.data_ptr(), w.data_ptr(), b.data_ptr(), ga.data_ptr(), N, H, W, w.size(0),
              g.data_ptr(), _lib.stream())
    return g


def hwr_stem_bwd_expand(img, w, b, ga):
    """ga [N,H/2,W/2,Cout] bf16 -> gradient w.r.t. conv0's output [N,H,W,Cout] bf16."""
    N, _, H, W = img.shape
    Cout = w.size(0)
    gc0 = torch.empty((N, H, W, Cout), device=img.device, dtype=torch.bfloat16)
    _lib.call("hwg_hwr_stem_bwd_expand", img.data_ptr(), w.data_ptr(), b.data_ptr(), ga.data_ptr(), N, H, W, Cout,
              gc0.data_ptr(), _lib.stream())
    return gc0
