#!/usr/bin/env python
"""bench.py — throughput of the hot path on N B200s (driver contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default workload (`--workload gan_train`, bench_gan_train.py) = BASELINE.json's metric on configs[2] / configs[3]: the
reference's balanced optimizer step of the GAN curriculum (trainer :300-391, :724-748) — generator forward, frozen recognizer
+ CTC and frozen discriminator, the two losses back-propagated separately through the generator and stashed, second generator
forward + Encoder2 perceptual loss, per-tensor gradient balancing, clip + Adam; gradient all-reduce at N > 1 — at a GLOBAL
batch of 128 lines of 64x1024 px (strong scaling: 128 / N lines per GPU).
`--workload gen_infer` (this file) = configs[1]: pure_gen generator inference, batch 32 per GPU, T_s=256;
`--workload hwr_train` (bench_hwr_train.py) = configs[0]: recognizer + CTC train step, batch 8 per GPU.
A "step" is one pass of that path over one batch of synthetic input (bench_inputs.py).

  value     lines/s with the inputs already resident in HBM (device-timed, CUDA events)
  e2e       the same through the public API (SpacedGenerator.forward) from pinned HOST buffers:
            H2D of content+style and D2H of the generated images inside the timed region
  roofline  the dominant kernel (conv_fprop_kernel, tensor-bound): algorithmic conv FLOPs per
            step / summed CUDA-event time of its launches, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the reference's CPU implementation of the same path (oracle port: torch fp32 on all
            host threads) on a bounded sample of the workload
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


# ----------------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# ----------------------------------------------------------------------------------------------
GEN = dict(n_class=80, style_dim=128, dim=256, T=256, B=32)


def gen_conv_layers(T, n_in=208, dim=256):
    """Algorithmic forward conv FLOPs of SpacedGenerator per line, layer by layer (2*pixels*Cout*Cin*taps of the
    reference's layers; transposed convs counted on their input pixels), with the activation bytes a
    layer must move at bf16 (input read once + output written once) and the kernel that serves it."""
    c = [dim, dim // 2, dim // 4, dim // 8, dim // 16]
    L = []
    H, W = 1, T
    L.append(("b0.conv1", 2 * (H * W) * n_in * c[0] * 12, (H * W * 256 + 4 * W * c[0]) * 2, "conv_fprop_kernel"))
    H = 4
    L.append(("b0.conv2", 2 * (H * W) * c[0] * c[0] * 9, 2 * H * W * c[0] * 2, "conv_fprop_kernel"))
    for i in (1, 2):                               # upsample(2,1) + conv3x3, conv2
        L.append((f"b{i}.conv1", 2 * (2 * H * W) * c[i - 1] * c[i] * 9, (H * W * c[i - 1] + 2 * H * W * c[i]) * 2,
                  "conv_fprop_kernel"))
        H *= 2
        L.append((f"b{i}.conv2", 2 * (H * W) * c[i] * c[i] * 9, 2 * H * W * c[i] * 2, "conv_fprop_kernel"))
    for i in (3, 4):                               # FusedUpsample 4x4 s2 (on input pixels), conv2
        L.append((f"b{i}.conv1", 2 * (H * W) * c[i - 1] * c[i] * 16, (H * W * c[i - 1] + 4 * H * W * c[i]) * 2,
                  "conv_small_kernel"))
        H, W = 2 * H, 2 * W
        L.append((f"b{i}.conv2", 2 * (H * W) * c[i] * c[i] * 9, 2 * H * W * c[i] * 2, "conv_small_kernel"))
    return L, 2 * (H * W) * c[4]                   # + the 1x1 output conv (fused into gen_output_kernel)


def gen_conv_flops_per_line(T, n_in=208, dim=256):
    L, out = gen_conv_layers(T, n_in, dim)
    return sum(l[1] for l in L) + out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_generator_lines_per_s(sample_B, reps, seed=0):
    from oracle import gen as ogen, synth
    from handwriting_line_generation_b200.pure_gen import SpacedGenerator  # parameter container only (CPU)
    torch.manual_seed(seed)
    sd = SpacedGenerator(GEN["n_class"], GEN["style_dim"], GEN["dim"], n_style_trans=6, emb_dropout=False,
                         append_style=True, small=False).state_dict()
    content, style = synth.gen_case(GEN["T"], sample_B, GEN["n_class"], GEN["style_dim"], seed)
    c, s = torch.from_numpy(content), torch.from_numpy(style)
    shapes = synth.gen_noise_shapes(GEN["T"], sample_B, GEN["dim"])
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            noise = [torch.randn(sh) for sh in shapes]          # the reference draws its noise inside forward
            ogen.generator_forward(sd, c, s, noise)
            if i:                                                # first call warms the thread pool / allocator
                times.append(time.perf_counter() - t0)
    return sample_B / statistics.median(times), times


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sample_B = 16
    t0 = time.time()
    lps, times = cpu_generator_lines_per_s(sample_B, max(1, args.steps))
    sample = (f"{len(times)} timed passes of the generator forward on a {sample_B}-line slice of the batch "
              f"(T_s={GEN['T']}, 64x{4 * GEN['T']} px), torch fp32 on {cores} host threads, {time.time() - t0:.1f}s")
    line = {"impl": "reference", "metric": "generated lines/sec", "value": lps, "unit": "lines/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(GEN["B"]),
            "cpu_baseline": {"value": lps, "unit": "lines/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": lps, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(B):
    return {"workload": "BASELINE configs[1]: pure_gen SpacedGenerator inference (char-spec content + random style "
                        f"vectors -> 64x{4 * GEN['T']} px lines), T_s={GEN['T']}, IAM charset (80 classes), "
                        "random-init weights, synthetic text",
            "batch_per_gpu": B, "line_px": [64, 4 * GEN["T"]],
            "l2": "no explicit flush: the bf16 activations one step streams (~0.7 GB at B=32) exceed the 126 MB L2; "
                  "weights (4 MB) stay cached, as in production",
            "noise": "NoiseInjection N(0,1) drawn in-kernel (counter-based hash + Box-Muller), a fresh seed every step",
            "execution": "one CUDA graph per step (fixed shapes), replayed; gpu_launches = kernels inside the replayed graphs"}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import conv as hconv
    import bench_inputs as synth  # input builders (numpy)

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl ours (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, T = GEN["B"], GEN["T"]
    torch.manual_seed(0)   # identical weights on every rank
    model = pkg.SpacedGenerator(GEN["n_class"], GEN["style_dim"], GEN["dim"], n_style_trans=6, emb_dropout=False,
                                append_style=True, small=False).to(dev).eval()
    n_sets = 4
    host_sets = []
    for i in range(n_sets):
        content, style = synth.gen_case(T, B, GEN["n_class"], GEN["style_dim"], 1000 * rank + i)
        host_sets.append((torch.from_numpy(content).pin_memory(), torch.from_numpy(style).pin_memory()))
    dev_sets = [(c.to(dev), s.to(dev)) for c, s in host_sets]
    out_host = [torch.empty((B, 1, 64, 4 * T), dtype=torch.float32).pin_memory() for _ in range(2)]
    torch.manual_seed(1234 + rank)  # per-rank noise streams

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    from handwriting_line_generation_b200 import graphs

    def fwd(c, s):
        with torch.no_grad():
            return model(c, s)

    # one eager step counts this path's kernel launches (a graph replay makes none on the host)
    for i in range(2):
        fwd(*dev_sets[i])
    torch.cuda.synchronize()
    n0 = pkg._lib.launch_count()
    fwd(*dev_sets[0])
    launches_per_step = pkg._lib.launch_count() - n0
    # the step is captured once in a CUDA graph (fixed shapes; a device-side counter re-seeds the noise every replay)
    graphed = graphs.GraphedStep(fwd, list(dev_sets[0]), modules=[model], warmup=max(3, args.warmup))
    in_c, in_s = graphed.static_in

    def step_device(i):
        c, s = dev_sets[i % n_sets]
        return graphed(c, s)           # device-to-device copy into the graph's input buffers + replay

    def step_e2e(i):
        hc, hs = host_sets[i % n_sets]
        in_c.copy_(hc, non_blocking=True)      # H2D from pinned memory straight into the graph's inputs
        in_s.copy_(hs, non_blocking=True)
        graphed.graph.replay()
        out_host[i % 2].copy_(graphed.static_out, non_blocking=True)   # D2H of the generated lines

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(max(3, args.warmup)):
        step_device(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t0 = time.time()
    ms = timed(step_device, args.steps)
    t1 = time.time()
    launches = launches_per_step * args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    # ---- rooflines: CUDA events around every convolution launch of a few eager steps (same kernels, same stream) ----
    prof = []
    hconv.PROFILE = prof
    psteps = min(args.steps, 5)
    barrier()
    for i in range(psteps):
        fwd(*dev_sets[i % n_sets])
    barrier()
    hconv.PROFILE = None
    peaks = load_peaks()
    layers, _ = gen_conv_layers(T)
    kern = {}
    for e0, e1, fl, kind, by, *_ in prof:
        k = kern.setdefault(kind, {"ms": 0.0, "launches": 0, "issued_flop": 0.0, "bytes": 0.0})
        k["ms"] += e0.elapsed_time(e1) / psteps
        k["launches"] += 1 / psteps
        k["issued_flop"] += fl / psteps
        k["bytes"] += by / psteps
    for kind, k in kern.items():
        k["algorithmic_gflop_per_step"] = sum(l[1] for l in layers if l[3] == kind) * B / 1e9
        k["algorithmic_mb_per_step"] = sum(l[2] for l in layers if l[3] == kind) * B / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    lines = B * world
    ms_step = ms / args.steps
    value = lines / (ms_step * 1e-3)
    e2e_value = lines / (ms_e2e / args.steps * 1e-3)
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1:
        torch.set_num_threads(cores)
        tb = time.time()
        lps, times = cpu_generator_lines_per_s(16, 25)
        cpu = {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{len(times)} timed generator forwards on a 16-line half of the batch (T_s={T}), torch fp32, "
                         f"{time.time() - tb:.1f}s of CPU work"}

    def roof(kind):
        k = kern[kind]
        if kind == "conv_small_kernel":   # HBM-bound layers: activation bytes (read once + written once) / kernel time
            ach = k["algorithmic_mb_per_step"] * 1e6 / (k["ms"] * 1e-3) / 1e9
            r = {"bound": "hbm", "kernel": kind, "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                 "frac": ach / peaks["hbm"], "peak_source": f"{peaks['src']} HBM copy bandwidth",
                 "algorithmic_mb_per_step": k["algorithmic_mb_per_step"]}
        else:
            ach = k["algorithmic_gflop_per_step"] * 1e9 / (k["ms"] * 1e-3) / 1e12
            r = {"bound": "tensor", "kernel": kind, "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                 "frac": ach / peaks["tf_sust"],
                 "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside the step)",
                 "algorithmic_gflop_per_step": k["algorithmic_gflop_per_step"],
                 "issued_gflop_per_step": k["issued_flop"] / 1e9}
        r.update({"traffic": None, "launches_per_step": round(k["launches"]), "kernel_ms_per_step": k["ms"],
                  "share_of_step": k["ms"] / ms_step})
        return r

    top = max(kern, key=lambda kk: kern[kk]["ms"])
    line = {
        "metric": "generated lines/sec", "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(B),
        "e2e": {"value": e2e_value, "unit": "lines/s", "h2d_bytes_per_step": int(T * B * GEN["n_class"] * 4 + B * GEN["style_dim"] * 4),
                "d2h_bytes_per_step": int(B * 64 * 4 * T * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof(top),
        "roofline_other_kernels": [roof(kk) for kk in kern if kk != top],
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gan_train", choices=["gan_train", "gen_infer", "hwr_train"],
                    help="gan_train (default) = the GAN 'gen' lesson train step on the hot path (BASELINE metric); "
                         "gen_infer = BASELINE configs[1]; hwr_train = configs[0] recognizer+CTC train step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "gan_train":
        import bench_gan_train
        bench_gan_train.main(args, rank, world, local_rank, load_peaks, ClockSampler)
    elif args.workload == "hwr_train":
        import bench_hwr_train
        bench_hwr_train.main(args, rank, world, local_rank, load_peaks, ClockSampler)
    elif args.impl == "reference":
        if args.steps > 40:
            args.steps = 40   # bounded: each step is a 16-line half batch on the CPU (~0.4 s)
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
