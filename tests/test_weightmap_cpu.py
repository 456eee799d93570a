"""CPU: the hwg_linear_map job geometry (weightmap.ConvMap) reproduces the torch re-layouts of every generator
convolution flavour — forward pack, dgrad pack — and the wgrad unpack is the exact adjoint of the forward pack."""
import numpy as np
import pytest
import torch

from handwriting_line_generation_b200 import weightmap as wm
from tests import ref_map, ref_pack


def _close(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.equal(a.float(), b.float()) or (a.float() - b.float()).abs().max() <= 4e-3 * b.float().abs().max()


CASES = [
    ("conv3x3", lambda: wm.map_conv3x3(32, 16), (32, 16, 3, 3)),
    ("initial", lambda: wm.map_initial(22, 32), (22, 32, 4, 3)),          # Ci=22 -> padded operands
    ("vert_up", lambda: wm.map_vert_up(16, 32), (16, 32, 3, 3)),
    ("fused_up", lambda: wm.map_fused_up(32, 16, 0.0589), (32, 16, 3, 3)),
]


@pytest.mark.parametrize("name,mk,shape", CASES, ids=[c[0] for c in CASES])
def test_pack_matches_torch(name, mk, shape):
    torch.manual_seed(0)
    w = torch.randn(shape)
    m = mk()
    t = wm.JobTable()
    if name == "initial":
        cin_pad = 64
        ref_f = ref_pack.initial_fwd(w, cin_pad)
        dst_f = torch.full_like(ref_f, 7.0)
        C = m.Co
        m.add_pack_fwd(t, w, dst_f, out_off=[kx * 4 * C * cin_pad + r * C * cin_pad for r in range(4) for kx in range(3)],
                       Cip=cin_pad)
        ref_d = ref_pack.initial_dgrad(w)
    elif name == "conv3x3":
        ref_f, ref_d = ref_pack.conv3x3_fwd(w), ref_pack.conv3x3_dgrad(w)
        dst_f = torch.full_like(ref_f, 7.0)
        m.add_pack_fwd(t, w, dst_f)
    elif name == "vert_up":
        ref_f = torch.stack(ref_pack.vert_up_fwd(w), 0).reshape(12, m.Co, m.Ci)
        ref_d = ref_pack.vert_up_dgrad(w)
        dst_f = torch.full_like(ref_f, 7.0)
        m.add_pack_fwd(t, w, dst_f)
        assert wm.vert_taps(0) == [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1)]
        assert wm.vert_taps(1) == [(0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]
    else:
        ref_f, taps = ref_pack.fused_up_fwd(w, 0.0589)
        assert taps == wm.fused_taps()
        ref_d = ref_pack.fused_up_dgrad(w, 0.0589)
        dst_f = torch.full_like(ref_f, 7.0)
        m.add_pack_fwd(t, w, dst_f)
    dst_d = torch.full(m.dgrad_shape(), 7.0, dtype=torch.bfloat16)
    m.add_pack_dgrad(t, w, dst_d)
    ref_map.run_jobs_cpu(t)
    _close(dst_f, ref_f)
    _close(dst_d, ref_d)


@pytest.mark.parametrize("name,mk,shape", CASES, ids=[c[0] for c in CASES])
def test_unpack_is_adjoint_of_pack(name, mk, shape):
    """<pack(w), dW> == <w, unpack(dW)> for random w, dW (fp32 copies of the maps)."""
    torch.manual_seed(1)
    m = mk()
    w = torch.randn(shape)
    Cip = m.Cip
    # forward pack in fp32 through the same job (dst fp32)
    t = wm.JobTable()
    pf = torch.zeros((m.Tf, m.Co, Cip))
    t.add(w, pf, R=m.Co, C=m.Ci, Cp=Cip, s_r=m.s_co, s_c=m.s_ci, d_r=Cip, d_c=1, M=m.Af,
          out_off=[i * m.Co * Cip for i in range(m.Tf)])
    dW = torch.randn((m.Tf, m.Co, Cip))
    g = torch.zeros(shape)
    m.add_unpack_wgrad(t, dW, g)
    ref_map.run_jobs_cpu(t)
    lhs = (pf * dW).sum().item()
    rhs = (w * g).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs)), (lhs, rhs)


def test_sum_job_and_accumulate():
    t = wm.JobTable()
    st = torch.randn(5, 8, 2)                 # per-sample (sum, sumsq) pairs
    out = torch.ones(8)
    t.add(st, out, R=1, C=8, s_r=0, s_c=2, d_r=0, d_c=1, M=None, nin=5, in_stride=16, accumulate=True, scale=0.5)
    ref_map.run_jobs_cpu(t)
    assert torch.allclose(out, 1 + 0.5 * st[:, :, 0].sum(0), atol=1e-6)
