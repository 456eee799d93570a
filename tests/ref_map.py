"""TEST INFRASTRUCTURE — CPU interpreter of hwg_linear_map job tables (include/hwg_b200.h semantics), used to check
the job geometry against the ATen re-layouts without a GPU and the CUDA kernel against this on the GPU."""
import numpy as np
import torch


def _flat(t):
    """(flat view of the tensor's whole storage, element offset of t in it): jobs address memory as data_ptr + explicit
    strides, so t may be any view (e.g. one kernel row of a weight)."""
    t = t.detach()
    n = t.untyped_storage().nbytes() // t.element_size()
    return torch.as_strided(t, (n,), (1,), 0), t.storage_offset()


def run_jobs_cpu(table, src_base=None, dst_base=None):
    """Applies every job of a weightmap.JobTable to CPU tensors (in place on the dst tensors)."""
    for j in table.jobs:
        if isinstance(j["src"], torch.Tensor):
            src, so = _flat(j["src"])
            src = src.float().numpy()
        else:
            src, so = _flat(src_base)
            src, so = src.float().numpy(), so + j["src"] // 4
        if isinstance(j["dst"], torch.Tensor):
            dst_t, do = _flat(j["dst"])
        else:
            dst_t, do = _flat(dst_base)
            do += j["dst"] // dst_t.element_size()
        out = dst_t.float().numpy().copy()
        scale = np.float32(j["scale"]) * (np.float32(1.0) if j.get("scale_dev") is None else np.float32(j["scale_dev"].item()))
        rows, cols = np.arange(j["Rp"])[:, None], np.arange(j["Cp"])[None, :]
        pad = (rows >= j["R"]) | (cols >= j["C"])
        sb = np.where(pad, 0, so + rows * j["s_r"] + cols * j["s_c"])
        db = do + rows * j["d_r"] + cols * j["d_c"]
        if j["M"] is None:
            acc = sum(src[sb + i * j["in_stride"]] for i in range(j["nin"]))
            vals = [np.where(pad, 0.0, scale * acc)]
        else:
            v = np.stack([np.where(pad, 0.0, src[np.where(pad, 0, sb + o)]) for o in j["in_off"]], -1)   # [Rp,Cp,nin]
            res = (v.astype(np.float32) @ j["M"]) * scale
            vals = [res[..., o] for o in range(j["nout"])]
        touched = []
        for o, val in zip(j["out_off"], vals):
            if j["accumulate"]:
                np.add.at(out, (db + o)[~pad], val[~pad])
                touched.append((db + o)[~pad].reshape(-1))
            else:
                out[db + o] = val
                touched.append((db + o).reshape(-1))
        # write back ONLY what the job addresses, as the kernel does: the destination may share its storage with regions
        # another thread is working on (the data-parallel tests all-reduce slices of the flat gradient buffer in the
        # background while later jobs fill other slices)
        idx = np.concatenate(touched)                  # duplicates carry the same value
        dst_t[torch.from_numpy(idx)] = torch.from_numpy(out[idx]).to(dst_t.dtype)
