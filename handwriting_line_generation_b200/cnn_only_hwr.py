"""Drop-in for the reference recognizer `CNNOnlyHWR` (model/cnn_only_hwr.py:7-117).

Same constructor signature, module names and `state_dict` keys as the reference (so the released
`hwr.*` snapshots load).  `forward` runs on libhwg_b200: a fused stem (conv0+ReLU+pool), tcgen05
implicit-GEMM convolutions with bias/ReLU or BatchNorm-statistics epilogues, NHWC bf16 pool and
BatchNorm-apply passes, and a head whose epilogue does the log-softmax and stores [T,B,C] fp32
directly (the reference's permute at :105 costs nothing here).  The torch modules are parameter
containers; their forward is never called.
"""
import torch
import torch.nn as nn

from . import _lib, conv, ops, weightmap
from ._lib import ACT_LOGSOFTMAX, ACT_NONE, ACT_RELU

_PADS = [1, 1, 1, 1, 1, 0, 0]
_NM = [64, 128, 256, 256, 512, 512, 512]
# 1-D head: (conv index, bn index, padding, dilation) — cnn_only_hwr.py:78-89
_CNN1D = [(0, 1, 2, 2), (3, 4, 4, 4), (6, 7, 0, 1), (9, 10, 8, 8)]


class CNNOnlyHWR(nn.Module):
    def __init__(self, nclass, nc=1, cnnOutSize=512, nh=512, leakyRelu=False, norm='group', small=False,
                 pad=False):
        super().__init__()
        if norm != 'batch':
            raise NotImplementedError("only norm='batch' (config \"hwr\": \"CNNOnly batchnorm\") is implemented")
        if leakyRelu or nc != 1:
            raise NotImplementedError("leakyRelu / nc != 1 are not used by the reference configs")
        if pad == 'less':
            self.pad = nn.ZeroPad2d((32 if small else 64,) * 2 + (0, 0))
        elif pad:
            self.pad = nn.ZeroPad2d((64 if small else 128,) * 2 + (0, 0))
        else:
            self.pad = None
        self.small = small
        cnn = nn.Sequential()

        def convRelu(i, bn=False):
            n_in = nc if i == 0 else _NM[i - 1]
            cnn.add_module('conv{0}'.format(i), nn.Conv2d(n_in, _NM[i], 3, 1, _PADS[i]))
            if bn:
                cnn.add_module('batchnorm{0}'.format(i), nn.BatchNorm2d(_NM[i]))
            cnn.add_module('relu{0}'.format(i), nn.ReLU(True))

        convRelu(0)
        if not small:
            cnn.add_module('pooling{0}'.format(0), nn.MaxPool2d(2, 2))
        convRelu(1)
        cnn.add_module('pooling{0}'.format(1), nn.MaxPool2d(2, 2))
        convRelu(2, True)
        convRelu(3)
        cnn.add_module('pooling{0}'.format(2), nn.MaxPool2d((2, 2), (2, 1), (0, 1)))
        convRelu(4, True)
        convRelu(5)
        cnn.add_module('pooling{0}'.format(3), nn.MaxPool2d((2, 2), (2, 1), (0, 1)))
        convRelu(6, True)
        self.cnn = cnn
        size1d = 512
        self.cnn1d = nn.Sequential(
            nn.Conv1d(size1d, size1d, 3, 1, 2, 2), nn.BatchNorm1d(size1d), nn.ReLU(True),
            nn.Conv1d(size1d, size1d, 3, 1, 4, 4), nn.BatchNorm1d(size1d), nn.ReLU(True),
            nn.Conv1d(size1d, size1d, 3, 1, 0, 1), nn.BatchNorm1d(size1d), nn.ReLU(True),
            nn.Conv1d(size1d, size1d, 3, 1, 8, 8), nn.BatchNorm1d(size1d), nn.ReLU(True),
            nn.Conv1d(size1d, nclass, 3, 1, 0, 1),
            nn.LogSoftmax(dim=1))
        self.nclass = nclass
        self.saved_features = None
        self._save = False
        self._cache_key, self._cache = None, None
        self._plan, self._plan_ptrs, self._bwd_plans = None, None, {}
        # data parallelism: set to a dp.PeerExchange (in-kernel exchange over NVLink peer memory) or a
        # torch.distributed group (NCCL all-reduce per layer) to normalise train-mode BatchNorm with the statistics of
        # the JOINT batch of the ranks (the reference is single-process, so its batch statistics span the whole
        # batch); None = per-rank statistics (what DistributedDataParallel does by default)
        self.sync_bn_group = None

    def setup_save_features(self):
        """cnn_only_hwr.py:109-117 hooks cnn[15] (conv5, whose in-place ReLU has run by the time anyone
        reads it): here the post-ReLU conv5 activation is exported as [B,512,H,W] fp32 on request."""
        self._save = True
        self.saved_features = [None]

    # -- derived weights --------------------------------------------------------------------------
    def conv_layers(self):
        """(key, parameter module, tap list) of every tensor-core convolution, in forward order."""
        t3, t3p0 = conv.conv_taps(3, 3, 1, 1), conv.conv_taps(3, 3, 0, 0)
        out = [(f"w{i}", getattr(self.cnn, f"conv{i}"), t3 if i <= 4 else t3p0) for i in range(1, 7)]
        for ci, _, pad, dil in _CNN1D + [(12, None, 0, 1)]:
            out.append((f"v{ci}", self.cnn1d[ci], conv.conv_taps(1, 3, 0, pad, 1, dil)))
        return out

    def _build_plan(self):
        """Persistent kernel-side operand buffers + ONE hwg_linear_map job table that (re)fills them from the fp32
        parameters (weightmap.py): tap-major bf16 forward and dgrad operands of the eleven tensor-core convolutions."""
        dev = self.cnn.conv0.weight.device
        t = weightmap.JobTable()
        c = {"maps": {}, "dgrad": {}}
        for key, mod, taps in self.conv_layers():
            w = mod.weight
            assert w.is_contiguous() and w.dtype == torch.float32
            m = weightmap.map_conv_taps(w.size(0), w.size(1), taps)
            c[key] = torch.empty((m.Tf, m.Co, m.Cip), device=dev, dtype=torch.bfloat16)
            d = torch.empty(m.dgrad_shape(), device=dev, dtype=torch.bfloat16)
            m.add_pack_fwd(t, w, c[key])
            m.add_pack_dgrad(t, w, d)
            c["maps"][key] = m
            c["dgrad"][key] = (d, m.taps_d)
            c[("b" if key[0] == "w" else "c") + key[1:]] = mod.bias.detach()
        c["w0"] = self.cnn.conv0.weight.detach().reshape(64, 9)        # stem: fp32, read directly (a view)
        c["b0"] = self.cnn.conv0.bias.detach()
        t.finalize(dev)
        return {"table": t, "c": c}

    def _packed(self):
        key = tuple((p.data_ptr(), p._version) for p in _lib.params(self))
        if self._cache_key == key:
            return self._cache
        ptrs = tuple(k[0] for k in key)
        if self._plan is None or self._plan_ptrs != ptrs:
            self._plan, self._plan_ptrs = self._build_plan(), ptrs
            self._bwd_plans = {}
        self._plan["table"].run()
        self._cache_key, self._cache = key, self._plan["c"]
        return self._cache

    def _bn(self, y, stats, bn):
        """BatchNorm (+ReLU) on an NHWC bf16 activation whose per-(n,c) sums came from the conv epilogue."""
        N, H, W, C = y.shape
        use_batch = self.training or not bn.track_running_stats
        momentum = 0.1 if bn.momentum is None else bn.momentum
        if use_batch and self.training and self.sync_bn_group is not None:
            coef, save = ops.bn_coeffs_synced(stats, N, C, H * W, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                              bn.running_var, momentum, bn.eps, self.sync_bn_group, key=("f", id(bn)))
        else:
            coef, save = ops.bn_coeffs(stats, N, C, H * W, bn.weight.detach(), bn.bias.detach(),
                                       bn.running_mean if (self.training or not use_batch) else None,
                                       bn.running_var if (self.training or not use_batch) else None,
                                       momentum, bn.eps, use_batch)
        if self.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        return ops.scale_shift_act(y, coef, False, ACT_RELU), save

    def forward(self, input, style=None):
        _lib.require_cuda(input)
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in _lib.params(self))):
            from .autograd_hwr import hwr_apply  # backward pass lives there
            return hwr_apply(self, input)
        return self._forward_impl(input)

    def _forward_impl(self, input):
        if self.pad is not None:
            input = self.pad(input)
        if self.small:
            raise NotImplementedError("small=True (no first pooling) is not implemented")
        c = self._packed()
        x = input.float().contiguous()
        B, _, H, W = x.shape
        dev = x.device
        need_stats = self.training

        def stats_for(C):
            return torch.zeros((B, C, 2), device=dev, dtype=torch.float32) if need_stats else None

        t3 = conv.conv_taps(3, 3, 1, 1)
        t3p0 = conv.conv_taps(3, 3, 0, 0)
        a = ops.hwr_stem(x, c["w0"], c["b0"])                                         # [B,H/2,W/2,64]
        a = conv.conv_fprop(a, c["w1"], t3, a.size(1), a.size(2), bias=c["b1"], act=ACT_RELU)
        a = ops.maxpool_nhwc(a, (2, 2), (2, 2), (0, 0))                                 # [B,H/4,W/4,128]
        st = stats_for(256)
        a = conv.conv_fprop(a, c["w2"], t3, a.size(1), a.size(2), bias=c["b2"], stats=st)
        a, _ = self._bn(a, st, self.cnn.batchnorm2)
        a = conv.conv_fprop(a, c["w3"], t3, a.size(1), a.size(2), bias=c["b3"], act=ACT_RELU)
        a = ops.maxpool_nhwc(a, (2, 2), (2, 1), (0, 1))                                 # [B,H/8,W/4+1,256]
        st = stats_for(512)
        a = conv.conv_fprop(a, c["w4"], t3, a.size(1), a.size(2), bias=c["b4"], stats=st)
        a, _ = self._bn(a, st, self.cnn.batchnorm4)
        a = conv.conv_fprop(a, c["w5"], t3p0, a.size(1) - 2, a.size(2) - 2, bias=c["b5"], act=ACT_RELU)
        if self._save:
            self.saved_features[0] = a.permute(0, 3, 1, 2).float()
        a = ops.maxpool_nhwc(a, (2, 2), (2, 1), (0, 1))
        st = stats_for(512)
        a = conv.conv_fprop(a, c["w6"], t3p0, a.size(1) - 2, a.size(2) - 2, bias=c["b6"], stats=st)
        a, _ = self._bn(a, st, self.cnn.batchnorm6)
        if a.size(1) != 1:
            # the reference folds a residual height into channels (view(b,-1,w)); its Conv1d(512,...) then
            # fails unless h == 1, i.e. the image height is 64
            raise RuntimeError(f"CNNOnlyHWR expects 64-px-high images (conv height {a.size(1)} != 1)")
        for ci, bi, pad, dil in _CNN1D:
            Wi = a.size(2)
            Wo = Wi + 2 * pad - 2 * dil
            st = stats_for(512)
            a = conv.conv_fprop(a, c[f"v{ci}"], conv.conv_taps(1, 3, 0, pad, 1, dil), 1, Wo, bias=c[f"c{ci}"], stats=st)
            a, _ = self._bn(a, st, self.cnn1d[bi])
        T = a.size(2) - 2
        C = self.nclass
        out = torch.empty((T, B, C), device=dev, dtype=torch.float32)
        conv.conv_fprop(a, c["v12"], conv.conv_taps(1, 3, 0, 0), 1, T, bias=c["c12"], act=ACT_LOGSOFTMAX,
                        out_view=(out, C, 0, B * C, 0))
        return out
