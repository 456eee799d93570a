"""Job tables for hwg_linear_map (include/hwg_b200.h): every weight re-layout of a module in one launch.

A convolution parameter w (kernel positions k = 0..K-1 contiguous, channel strides s_co / s_ci) and its `ConvMap`
matrices describe three linear maps:
  pack forward   Pf[t][co][ci] = sum_k Af[k][t] w[co,ci,k]      bf16, the operand of hwg_conv_fprop
  pack dgrad     Pd[t][ci][co] = sum_k Ad[k][t] w[co,ci,k]      bf16, the operand of the input-gradient convolution
  unpack wgrad   g_w[co,ci,k]  = sum_t Af[k][t] dW[t][co][ci]   fp32, the adjoint of the forward pack applied to
                                                                 hwg_conv_wgrad's tap-major output
The reference does the same re-parameterisations with ATen ops inside forward (EqualLR hook pure_gen.py:222-241,
FusedUpsample pure_gen.py:259-271) and leaves the adjoints to autograd.
"""
import ctypes

import numpy as np
import torch

from . import _lib

MAP_MAX = 16


class MapJob(ctypes.Structure):
    """struct hwgMapJob (include/hwg_b200.h)."""
    _fields_ = [
        ("src_off", ctypes.c_int64), ("dst_off", ctypes.c_int64), ("M", ctypes.c_void_p),
        ("nin", ctypes.c_int32), ("nout", ctypes.c_int32),
        ("R", ctypes.c_int32), ("C", ctypes.c_int32), ("Rp", ctypes.c_int32), ("Cp", ctypes.c_int32),
        ("dst_bf16", ctypes.c_int32), ("accumulate", ctypes.c_int32),
        ("scale", ctypes.c_float), ("flags", ctypes.c_int32),
        ("in_stride", ctypes.c_int64),
        ("s_r", ctypes.c_int64), ("s_c", ctypes.c_int64), ("d_r", ctypes.c_int64), ("d_c", ctypes.c_int64),
        ("in_off", ctypes.c_int64 * MAP_MAX), ("out_off", ctypes.c_int64 * MAP_MAX),
        ("scale_dev", ctypes.c_void_p),
    ]


class JobTable:
    """Collects jobs on the host, uploads the table once, launches it with `run`.

    src/dst of a job are either tensors (absolute addresses; the table must be rebuilt if they are re-allocated) or
    integer BYTE offsets relative to the bases passed to run()."""

    def __init__(self):
        self.jobs = []          # python dicts (also read by the CPU checker in tests/)
        self._dev = None
        self._keep = []

    def add(self, src, dst, *, R, C, s_r, s_c, d_r, d_c, M=None, in_off=None, out_off=None, nin=None, in_stride=0,
            Rp=None, Cp=None, dst_bf16=False, accumulate=False, scale=1.0, scale_dev=None):
        if M is not None:
            M = np.ascontiguousarray(M, dtype=np.float32)
            nin, nout = M.shape
            assert nin <= MAP_MAX and nout <= MAP_MAX
            in_off = list(range(nin)) if in_off is None else list(in_off)
            assert len(in_off) == nin
        else:
            nout = 1
            assert nin is not None and in_stride != 0
            in_off = []
        out_off = [0] if out_off is None else list(out_off)
        assert len(out_off) == nout
        self.jobs.append(dict(src=src, dst=dst, M=M, nin=nin, nout=nout, R=R, C=C, Rp=Rp or R, Cp=Cp or C,
                              dst_bf16=bool(dst_bf16), accumulate=bool(accumulate), scale=float(scale), scale_dev=scale_dev,
                              in_stride=in_stride, s_r=s_r, s_c=s_c, d_r=d_r, d_c=d_c, in_off=in_off, out_off=out_off))
        self._dev = None

    def finalize(self, device):
        ms = [j["M"].reshape(-1) for j in self.jobs if j["M"] is not None]
        mflat = torch.from_numpy(np.concatenate(ms) if ms else np.zeros(1, np.float32)).to(device)
        arr = (MapJob * len(self.jobs))()
        moff = 0
        per_block = _lib.load().hwg_map_items_per_block()
        tab = []
        for ji, (a, j) in enumerate(zip(arr, self.jobs)):
            a.flags = 0
            for bit, key in enumerate(("src", "dst")):
                v = j[key]
                if isinstance(v, torch.Tensor):      # absolute address
                    setattr(a, key + "_off", v.data_ptr())
                    a.flags |= 1 << bit
                else:                                # byte offset relative to the base passed to run()
                    setattr(a, key + "_off", int(v))
            if j["M"] is not None:
                a.M = mflat.data_ptr() + 4 * moff
                moff += j["M"].size
            else:
                a.M = None
            a.nin, a.nout, a.R, a.C, a.Rp, a.Cp = j["nin"], j["nout"], j["R"], j["C"], j["Rp"], j["Cp"]
            a.dst_bf16, a.accumulate, a.scale, a.in_stride = int(j["dst_bf16"]), int(j["accumulate"]), j["scale"], j["in_stride"]
            a.scale_dev = None if j.get("scale_dev") is None else j["scale_dev"].data_ptr()
            a.s_r, a.s_c, a.d_r, a.d_c = j["s_r"], j["s_c"], j["d_r"], j["d_c"]
            for i, v in enumerate(j["in_off"]):
                a.in_off[i] = v
            for i, v in enumerate(j["out_off"]):
                a.out_off[i] = v
            tab += [(ji, b) for b in range(-(-(j["Rp"] * j["Cp"]) // per_block))]
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self._tab = torch.tensor(tab, dtype=torch.int32).reshape(-1, 2).to(device)
        self._dev, self._keep = raw, [mflat]
        return self

    def run(self, src_base=None, dst_base=None):
        assert self._dev is not None, "JobTable.finalize() first"
        _lib.call("hwg_linear_map", self._dev.data_ptr(), self._tab.data_ptr(), self._tab.size(0),
                  _lib.ptr(src_base), _lib.ptr(dst_base), _lib.stream())


class ConvMap:
    """Linear relation between one convolution parameter and its kernel-side operands (see module docstring).
    Af [K][Tf], Ad [K][Td] numpy; taps_f / taps_d = tap lists of the forward / dgrad launches."""

    def __init__(self, Co, Ci, K, transposed_param, Af, Ad, taps_d, d_in_stride=(1, 1)):
        self.Co, self.Ci, self.K = Co, Ci, K
        self.s_co, self.s_ci = (K, Co * K) if transposed_param else (Ci * K, K)
        self.Af, self.Ad = np.asarray(Af, np.float32), np.asarray(Ad, np.float32)
        self.Tf, self.Td = self.Af.shape[1], self.Ad.shape[1]
        self.taps_d, self.d_in_stride = taps_d, d_in_stride
        self.Cip, self.Cop = -(-Ci // 16) * 16, -(-Co // 16) * 16

    # -- jobs ---------------------------------------------------------------------------------------------
    def add_pack_fwd(self, table, w, dst, out_off=None, Cip=None, scale_dev=None, Cop=None):
        """dst bf16 [Tf][Co][Cip] unless out_off (element offsets per tap) says otherwise; Cop > Co zero-pads the
        output-channel rows ([Tf][Cop][Cip]); scale_dev: 1-element device tensor multiplied in at run time."""
        Cip = Cip or self.Cip
        rows = Cop or self.Co
        out_off = [t * rows * Cip for t in range(self.Tf)] if out_off is None else out_off
        table.add(w, dst, R=self.Co, C=self.Ci, Rp=rows, Cp=Cip, s_r=self.s_co, s_c=self.s_ci, d_r=Cip, d_c=1, M=self.Af,
                  out_off=out_off, dst_bf16=True, scale_dev=scale_dev)

    def add_pack_dgrad(self, table, w, dst, scale_dev=None):
        """dst bf16 [Td][Cip16][Cop]: the transposed tap matrices."""
        Rp = -(-self.Ci // 16) * 16
        table.add(w, dst, R=self.Ci, C=self.Co, Rp=Rp, Cp=self.Cop, s_r=self.s_ci, s_c=self.s_co, d_r=self.Cop, d_c=1,
                  M=self.Ad, out_off=[t * Rp * self.Cop for t in range(self.Td)], dst_bf16=True, scale_dev=scale_dev)

    def dgrad_shape(self):
        return (self.Td, -(-self.Ci // 16) * 16, self.Cop)

    def add_unpack_wgrad(self, table, dw_src, g_dst, Cip=None, accumulate=False, co_rows=None, scale_dev=None):
        """dw_src fp32 [Tf][co_rows][Cip] (hwg_conv_wgrad layout; co_rows >= Co when the launch padded the output
        channels) -> g_dst in the parameter's layout."""
        Cip = Cip or self.Cip
        co_rows = co_rows or self.Co
        table.add(dw_src, g_dst, R=self.Co, C=self.Ci, s_r=Cip, s_c=1, d_r=self.s_co, d_c=self.s_ci,
                  M=self.Af.T.copy(), in_off=[t * co_rows * Cip for t in range(self.Tf)],
                  out_off=list(range(self.K)), accumulate=accumulate, scale_dev=scale_dev)


# ---- the generator's convolution flavours (model/pure_gen.py) -------------------------------------------------
TAPS3x3 = [(i - 1, j - 1) for i in range(3) for j in range(3)]


def map_conv3x3(Co, Ci):
    """nn.Conv2d 3x3 pad 1 (pure_gen.py:197): taps in row-major kernel order; dgrad = transposed matrices, negated taps."""
    return ConvMap(Co, Ci, 9, False, np.eye(9), np.eye(9), [(-dh, -dw) for dh, dw in TAPS3x3])


def map_conv_taps(Co, Ci, taps):
    """Stride-1 nn.Conv2d / nn.Conv1d with K = len(taps) kernel positions (cnn_only_hwr.py:31,78-90): identity maps;
    dgrad = transposed matrices with negated taps."""
    K = len(taps)
    return ConvMap(Co, Ci, K, False, np.eye(K), np.eye(K), [(-dh, -dw) for dh, dw in taps])


def map_initial(Ci, Co):
    """nn.ConvTranspose2d(Ci, Co, (4,3), padding=(0,1)) on H=1 (pure_gen.py:161-163): output row r is a 1x3 correlation
    with flipped kx: out[r,x] = sum_kx in[x+1-kx] W[:, :, r, kx].  Forward taps per row (0, 1-kx), t = r*3+kx;
    dgrad: 12-tap convolution of gy [B,4,T,C] with taps (r, kx-1)."""
    return ConvMap(Co, Ci, 12, True, np.eye(12), np.eye(12), [(r, kx - 1) for r in range(4) for kx in range(3)])


def vert_src(par, kh):
    return (par + kh - 1) // 2


def vert_taps(par):
    dhs = sorted({vert_src(par, kh) for kh in range(3)})
    return [(dh, kw - 1) for dh in dhs for kw in range(3)]


def map_vert_up(Co, Ci):
    """nn.Upsample((2,1), nearest) + Conv2d 3x3 (pure_gen.py:176-186) = two row-parity convolutions on the
    un-upsampled input; kernel rows that read the same source row are summed.  t = par*6 + j*3 + kw.
    dgrad: gy rows r = 2*h + par feed input row h: 12 taps (dh in -1..2, stride (2,1) on gy)."""
    Af = np.zeros((9, 12), np.float32)
    for par in (0, 1):
        dhs = sorted({vert_src(par, kh) for kh in range(3)})
        for kh in range(3):
            j = dhs.index(vert_src(par, kh))
            for kw in range(3):
                Af[kh * 3 + kw, par * 6 + j * 3 + kw] = 1
    comb = {-1: [2], 0: [1, 2], 1: [0, 1], 2: [0]}      # gy row offset -> kernel rows that land there
    Ad = np.zeros((9, 12), np.float32)
    taps_d = []
    for d, dh in enumerate((-1, 0, 1, 2)):
        for kw in range(3):
            taps_d.append((dh, 1 - kw))
            for kh in comb[dh]:
                Ad[kh * 3 + kw, d * 3 + kw] = 1
    return ConvMap(Co, Ci, 9, False, Af, Ad, taps_d, (2, 1))


FUSED_SEL = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}   # output parity -> [(input offset, 4x4 kernel index)]


def fused_taps():
    """Forward tap list of FusedUpsample as four output-parity 2x2 convolutions, t = ((py*2+px)*2+a)*2+b."""
    return [(dh, dw) for py in (0, 1) for px in (0, 1) for dh, _ in FUSED_SEL[py] for dw, _ in FUSED_SEL[px]]


def fused_phases():
    return [(py, px) for py in (0, 1) for px in (0, 1) for _ in range(4)]


def map_fused_up(Ci, Co, multiplier):
    """FusedUpsample (pure_gen.py:250-279): w4[ky,kx] = multiplier/4 * sum of the (up to four) 3x3 entries (i,j) with
    i in {ky-1,ky}, j in {kx-1,kx}; conv_transpose2d(stride 2, pad 1).  dgrad: 16-tap stride-2 convolution of gy."""
    A = np.zeros((9, 16), np.float32)
    for ky in range(4):
        for kx in range(4):
            for i in (ky - 1, ky):
                for j in (kx - 1, kx):
                    if 0 <= i < 3 and 0 <= j < 3:
                        A[i * 3 + j, ky * 4 + kx] = multiplier / 4
    order = [ky * 4 + kx for py in (0, 1) for px in (0, 1) for _, ky in FUSED_SEL[py] for _, kx in FUSED_SEL[px]]
    return ConvMap(Co, Ci, 9, True, A[:, order], A, [(ky - 1, kx - 1) for ky in range(4) for kx in range(4)], (2, 2))
