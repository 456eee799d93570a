"""bench.py --workload hwr_train: BASELINE configs[0] — CNNOnlyHWR + CTC train step (forward, CTC loss, backward
through every layer, Adam) on synthetic 64x1024 IAM-shaped lines, batch 8 per GPU, 80 classes, 60-char targets.
Same JSON contract as bench.py; data-parallel with the NCCL gradient all-reduce of
handwriting_line_generation_b200.dp.GradBuckets when WORLD_SIZE > 1."""
import json
import os
import statistics
import time

import numpy as np
import torch

HWR = dict(B=8, W=1024, S=60, C=80)
GF_FWD_PER_LINE = 24.661  # forward conv GFLOP per 64x1024 line; fwd+dgrad+wgrad = 3x


def config(B):
    return {"workload": "BASELINE configs[0]: cnn_only_hwr CTC recognizer train step (fwd + CTC loss + bwd + Adam) on "
                        f"synthetic 64x{HWR['W']} lines, IAM charset ({HWR['C']} classes), {HWR['S']}-char targets, "
                        "random-init weights",
            "batch_per_gpu": B, "line_px": [64, HWR["W"]],
            "l2": "activations + gradients of one step (~0.5 GB at B=8) exceed the 126 MB L2; no explicit flush"}


def cpu_lines_per_s(sample_B, reps):
    """The reference's CPU path for this step: torch fp32 modules/autograd (oracle port) + F.ctc_loss + Adam."""
    from oracle import hwr as ohwr, synth
    from handwriting_line_generation_b200 import CNNOnlyHWR
    torch.manual_seed(0)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in CNNOnlyHWR(HWR["C"], norm='batch').state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    img = torch.from_numpy(synth.hwr_case(sample_B, HWR["W"], 1))
    T = HWR["W"] // 4 - 6
    tg = torch.randint(1, HWR["C"], (sample_B, HWR["S"]), dtype=torch.int32)
    il, tl = torch.full((sample_B,), T, dtype=torch.int32), torch.full((sample_B,), HWR["S"], dtype=torch.int32)
    times = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        lp = ohwr.hwr_forward(sd, img, True, {})
        torch.nn.functional.ctc_loss(lp, tg, il, tl).backward()
        opt.step()
        if i:
            times.append(time.perf_counter() - t0)
    return sample_B / statistics.median(times), times


def main(args, rank, world, local_rank, load_peaks, ClockSampler):
    B = HWR["B"]
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        steps = min(args.steps, 30)
        t0 = time.time()
        lps, times = cpu_lines_per_s(4, steps)
        sample = (f"{len(times)} timed train steps on a 4-line half of the batch, torch fp32 on "
                  f"{torch.get_num_threads()} host threads, {time.time() - t0:.1f}s")
        print(json.dumps({"impl": "reference", "metric": "train-step lines/sec", "value": lps, "unit": "lines/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(B),
                          "cpu_baseline": {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(),
                                           "kind": "port", "sample": sample},
                          "e2e": {"value": lps, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return

    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import conv as hconv, dp
    import bench_inputs as synth
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = pkg.CNNOnlyHWR(HWR["C"], norm='batch').to(dev).train()
    # flat fused optimizer; the backward adds gradients straight into its buffer, the all-reduce buckets are slices of it
    opt = pkg.FlatAdam(model.parameters(), lr=1e-4, betas=(0.9, 0.999))
    model._grad_sink = opt
    reducer = dp.GradReducer(model.parameters(), bucket_bytes=16 << 20, flat=opt) if world > 1 else None
    if reducer is not None:
        model._grad_ready_cb = reducer.mark_ready
    T = HWR["W"] // 4 - 6
    n_sets = 4
    host = []
    for i in range(n_sets):
        img = torch.from_numpy(synth.hwr_case(B, HWR["W"], 100 * rank + i)).pin_memory()
        tg = torch.from_numpy(np.random.RandomState(i).randint(1, HWR["C"], (B, HWR["S"])).astype(np.int32)).pin_memory()
        host.append((img, tg))
    devsets = [(a.to(dev), b.to(dev)) for a, b in host]
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), HWR["S"], dtype=torch.int32, device=dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def train(img, tg):
        loss = pkg.CTCLoss(model(img), tg, il, tl)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    from handwriting_line_generation_b200 import graphs
    for i in range(2):
        train(*devsets[i])
    torch.cuda.synchronize()
    n0 = pkg._lib.launch_count()
    train(*devsets[0])
    launches_per_step = pkg._lib.launch_count() - n0
    graphed = None
    try:   # the whole step (forward, CTC, backward, all-reduce, optimizer) as one CUDA graph, replayed
        graphed = graphs.GraphedStep(train, list(devsets[0]), modules=[model], warmup=3)
    except Exception:
        graphed = None
        torch.cuda.synchronize()

    def step_device(i):
        if graphed is not None:
            return graphed(*devsets[i % n_sets])
        return train(*devsets[i % n_sets])

    def step_e2e(i):
        a, b = host[i % n_sets]
        if graphed is not None:
            graphed.static_in[0].copy_(a, non_blocking=True)
            graphed.static_in[1].copy_(b, non_blocking=True)
            graphed.graph.replay()
            loss = graphed.static_out
        else:
            loss = train(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True))
        loss_host.copy_(loss.detach(), non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return dp.max_over_ranks(e0.elapsed_time(e1), dev)

    for i in range(max(3, args.warmup)):
        step_device(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    n0 = pkg._lib.launch_count()
    t0 = time.time()
    ms = timed(step_device, args.steps)
    t1 = time.time()
    launches = (pkg._lib.launch_count() - n0) if graphed is None else launches_per_step * args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    prof = []
    hconv.PROFILE = prof
    psteps = min(args.steps, 5)
    barrier()
    for i in range(psteps):
        train(*devsets[i % n_sets])      # eager: per-launch events need host-side launches
    barrier()
    hconv.PROFILE = None
    conv_ms = sum(r[0].elapsed_time(r[1]) for r in prof) / psteps
    peaks = load_peaks()
    def leave():
        if world > 1:      # see bench_gan_train.leave(): collectives captured in a live graph block the teardown
            torch.cuda.synchronize()
            dist.barrier()
            import sys
            sys.stdout.flush()
            os._exit(0)

    if rank != 0:
        return leave()
    lines = B * world
    ms_step = ms / args.steps
    conv_flops = 3 * GF_FWD_PER_LINE * 1e9 * B
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    cpu = None
    if world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        tb = time.time()
        lps, times = cpu_lines_per_s(4, 20)
        cpu = {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{len(times)} timed train steps on a 4-line half of the batch, torch fp32, {time.time() - tb:.1f}s of CPU work"}
    print(json.dumps({
        "metric": "train-step lines/sec", "value": lines / (ms_step * 1e-3), "unit": "lines/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(B),
        "e2e": {"value": lines / (ms_e2e / args.steps * 1e-3), "unit": "lines/s",
                "h2d_bytes_per_step": int(B * 64 * HWR["W"] * 4 + B * HWR["S"] * 4), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_fprop_kernel + conv_wgrad_kernel (fprop, dgrad, wgrad)",
                     "achieved": achieved, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sust"],
                     "traffic": None, "peak_source": f"{peaks['src']} bf16 sustained",
                     "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / ms_step,
                     "algorithmic_gflop_per_step": conv_flops / 1e9},
        "cpu_baseline": cpu}), flush=True)
    leave()


# ----------------------------------------------------------------------------------------------------------
# Short device-timed measurements of the two training paths, embedded as "extra_workloads" in the default
# bench line (N=1) so that one bench run documents inference AND training throughput.
# ----------------------------------------------------------------------------------------------------------
def quick_train_numbers(dev, steps=10, gen_lesson=True):
    import handwriting_line_generation_b200 as pkg
    import bench_inputs as synth
    out = {}

    from handwriting_line_generation_b200 import graphs

    def time_steps(fn, mods):
        """fn() = one eager step with static inputs; timed as a replayed CUDA graph."""
        n0 = pkg._lib.launch_count()
        fn()
        launches = pkg._lib.launch_count() - n0
        g = graphs.GraphedStep(fn, [], modules=mods, warmup=2)
        for _ in range(2):
            g()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            g()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, launches

    # (1) configs[0]: recognizer + CTC train step, batch 8 and 32
    for B in (8, 32):
        torch.manual_seed(0)
        model = pkg.CNNOnlyHWR(HWR["C"], norm='batch').to(dev).train()
        opt = pkg.FlatAdam(model.parameters(), lr=1e-4, betas=(0.9, 0.999))
        model._grad_sink = opt
        T = HWR["W"] // 4 - 6
        img = torch.from_numpy(synth.hwr_case(B, HWR["W"], 1)).to(dev)
        tg = torch.randint(1, HWR["C"], (B, HWR["S"]), dtype=torch.int32, device=dev)
        il = torch.full((B,), T, dtype=torch.int32, device=dev)
        tl = torch.full((B,), HWR["S"], dtype=torch.int32, device=dev)

        def step():
            pkg.CTCLoss(model(img), tg, il, tl).backward()
            opt.step()

        ms, launches = time_steps(step, [model])
        out[f"hwr_ctc_train_step_B{B}"] = {"ms_per_step": ms, "lines_per_s": B / ms * 1e3, "hwg_launches_per_step": launches,
                                          "what": "CNNOnlyHWR fwd + CTC + bwd (dgrad+wgrad) + flat fused Adam, 64x1024 lines; one replayed CUDA graph per step"}
        del model, opt
    if not gen_lesson:
        return out
    # (2) the 'gen' lesson's recognition branch: generator -> frozen recognizer -> CTC -> backward into the
    #     generator -> Adam on the generator (trainer/hw_with_style_trainer.py:760-764), batch 16, T_s=256
    B, Ts = 16, 256
    torch.manual_seed(0)
    gen = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False).to(dev).train()
    hwr = pkg.CNNOnlyHWR(80, norm='batch').to(dev).train()
    for p in hwr.parameters():
        p.requires_grad_(False)
    opt = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=True)
    content, style = synth.gen_case(Ts, B, 80, 128, 3)
    c, s = torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev)
    T = Ts - 6
    tg = torch.randint(1, 80, (B, 40), dtype=torch.int32, device=dev)
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), 40, dtype=torch.int32, device=dev)

    def gstep():
        pkg.CTCLoss(hwr(gen(c, s)), tg, il, tl).backward()
        opt.step()
        opt.zero_grad(set_to_none=False)

    ms, launches = time_steps(gstep, [gen, hwr])
    out["gen_hwr_ctc_train_step_B16"] = {"ms_per_step": ms, "lines_per_s": B / ms * 1e3, "hwg_launches_per_step": launches,
                                         "what": "SpacedGenerator fwd+bwd, frozen CNNOnlyHWR fwd + input-gradient bwd, CTC, "
                                                 "Adam on the generator; 64x1024 lines (the recognition branch of the GAN 'gen' lesson); one replayed CUDA graph per step"}
    return out
