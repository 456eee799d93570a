"""TEST INFRASTRUCTURE — CPU interpreter of hwg_linear_map job tables (include/hwg_b200.h semantics), used to check
the job geometry against the ATen re-layouts without a GPU and the CUDA kernel against this on the GPU."""
import numpy as np
import torch


def _flat(t):
    assert t.is_contiguous()
    return t.view(-1)


def run_jobs_cpu(table, src_base=None, dst_base=None):
    """Applies every job of a weightmap.JobTable to CPU tensors (in place on the dst tensors)."""
    for j in table.jobs:
        if isinstance(j["src"], torch.Tensor):
            src, so = _flat(j["src"]).float().numpy(), 0
        else:
            src, so = _flat(src_base).float().numpy(), j["src"] // 4
        if isinstance(j["dst"], torch.Tensor):
            dst_t, do = _flat(j["dst"]), 0
        else:
            dst_t = _flat(dst_base)
            do = j["dst"] // dst_t.element_size()
        out = dst_t.float().numpy().copy()
        rows, cols = np.arange(j["Rp"])[:, None], np.arange(j["Cp"])[None, :]
        pad = (rows >= j["R"]) | (cols >= j["C"])
        sb = np.where(pad, 0, so + rows * j["s_r"] + cols * j["s_c"])
        db = do + rows * j["d_r"] + cols * j["d_c"]
        if j["M"] is None:
            acc = sum(src[sb + i * j["in_stride"]] for i in range(j["nin"]))
            vals = [np.where(pad, 0.0, j["scale"] * acc)]
        else:
            v = np.stack([np.where(pad, 0.0, src[np.where(pad, 0, sb + o)]) for o in j["in_off"]], -1)   # [Rp,Cp,nin]
            res = (v.astype(np.float32) @ j["M"]) * np.float32(j["scale"])
            vals = [res[..., o] for o in range(j["nout"])]
        for o, val in zip(j["out_off"], vals):
            if j["accumulate"]:
                np.add.at(out, (db + o)[~pad], val[~pad])
            else:
                out[db + o] = val
        dst_t.copy_(torch.from_numpy(out).to(dst_t.dtype))
