"""bench.py default workload: the HWWithStyle GAN "gen" lesson restricted to the SURVEY.md §8(a) rows — the
one data-parallel hot path BASELINE.json:north_star names.

One step (trainer/hw_with_style_trainer.py:514-530,752-764 `run_gen` with the genRecog loss, then :381-391):
  generator forward  (SpacedGenerator, model/pure_gen.py:42-50)          synthetic spaced text + style -> 64x1024 lines
  recognizer forward (CNNOnlyHWR, model/cnn_only_hwr.py:96-107)          frozen weights, BatchNorm in train mode
  CTC loss           (CTCLoss, model/loss.py:28-30)                      forward + backward
  backward           recognizer input-gradient chain (dgrad), generator dgrad + wgrad + norm/noise backward
  gradient all-reduce over NCCL (world > 1), launched from grad-ready hooks on a side stream (dp.GradReducer)
  clip_grad_value_(2) + Adam on the generator (lr 2e-4, betas (0.5, 0.999): configs/cf_IAM*.json:35-46, trainer :381)
The frozen discriminator's adversarial branch (HWG_BENCH_NO_DISC=1 drops it) runs inside the step; the
perceptual-encoder branch (Encoder2) is opt-in (HWG_BENCH_PERCEPTUAL=1) until that module has a green GPU parity run.

Batch 16 lines per GPU (weak scaling; 8 GPUs = BASELINE configs[3]'s global batch 128), T_s = 256 -> 64x1024 px,
IAM charset (80 classes), 40-character targets.
"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

GAN = dict(B=16, Ts=256, S=40, C=80, style=128, dim=256)
HWR_GF_FWD_PER_LINE = 24.661      # SURVEY.md §8a (a10), forward conv GFLOP per 64x1024 line
HWR_GF_STEM_PER_LINE = 0.075      # conv0 (fused stem kernel, not a tensor-core launch)
DISC_GF_FWD_PER_LINE = 11.295     # SURVEY.md §8d / Appendix D, DiscriminatorAP forward conv GFLOP per 64x1024 line
W_CTC, W_GEN = 1e-4, 1.0          # loss_weights genRecog / generator of the IAM GAN config (config json :58-61)
W_PERC = 0.5                      # loss_weights perceptual (config json :56)


def use_disc():
    return not os.environ.get("HWG_BENCH_NO_DISC")


def use_perceptual():
    """Opt-in (HWG_BENCH_PERCEPTUAL=1) until the Encoder2 drop-in has a green GPU parity run (encoder2.py: status): the
    perceptual branch of BASELINE configs[2] — L1 between Encoder2 features of a (synthetic) real line batch and of the
    generated lines (trainer :724-748), backward to the generated image."""
    return bool(os.environ.get("HWG_BENCH_PERCEPTUAL"))


def use_balanced():
    """Opt-in (HWG_BENCH_BALANCED=1, one GPU) until FlatAdam.stash()/balance() and Encoder2 have green GPU parity runs: the
    optimizer step as the reference's curriculum takes it with `balance_loss` — a no-step 'gen' lesson whose recognition
    and adversarial losses are back-propagated separately and stashed (trainer :312-338), then a lesson whose own gradient
    (here: the perceptual loss of the 'auto' lesson) the stashed sets are balanced into per parameter tensor
    (:340-377, `balance_var_x` of the config), then clip + Adam."""
    return bool(os.environ.get("HWG_BENCH_BALANCED"))


BALANCE_VAR_X = [0.6, 0.5]        # config :100 `balance_var_x` [0.6, 0.5, 0.4, 0.75]: the entries of the two sets stashed here
                                  # (recognition set first, adversarial set second, as the trainer stashes them)
DEFAULT_SYNC_BN = "peer"


def config(B, world, executed, sync_bn="off"):
    disc = ("frozen discriminator_ap fwd (train mode: spectral-norm power iteration, Dropout2d) + input-gradient bwd "
            "for the adversarial loss -mean(D(fake)), " if use_disc() else "")
    if use_balanced():
        return {"workload": "HWWithStyle GAN balanced optimizer step (BASELINE configs[2] shapes), the reference curriculum's pair of "
                            "lessons restricted to the built components: no-step 'gen' lesson = pure_gen generator fwd, frozen "
                            "discriminator_ap fwd+bwd and frozen cnn_only_hwr fwd+bwd + CTC, the adversarial and the recognition "
                            "loss back-propagated SEPARATELY through the generator and stashed; second lesson = generator fwd, "
                            "frozen Encoder2(32) perceptual loss against synthetic real lines, backward; per-tensor gradient "
                            "balancing of the two stashed sets into it (trainer :340-377), clip + Adam.  Two generator "
                            "forwards and three generator backwards per step; the style extractor / spacer lessons (SURVEY 8 "
                            "f3, f4) are not built",
                "batch_per_gpu": B, "global_batch": B * world, "line_px": [64, 4 * GAN["Ts"]], "classes": GAN["C"],
                "target_chars": GAN["S"], "parallelism": f"dp{world}", "execution": executed}
    return {"workload": "HWWithStyle GAN 'gen' lesson train step (BASELINE configs[2]/[3] shapes): "
                        "pure_gen generator fwd+bwd, frozen cnn_only_hwr fwd + input-gradient bwd (train-mode BatchNorm), "
                        "CTC loss fwd+bwd, " + disc + "ONE backward over the weighted sum of the two losses (the reference's per-loss gradient "
                        "balancing, trainer :300-377 / SURVEY 8 f2, which runs a backward per loss, is not built), "
                        "gradient all-reduce (N>1), clip + Adam on the generator; "
                        + ("perceptual branch (frozen Encoder2(32) on [synthetic real lines ; generated lines], L1 between the "
                           "halves of both feature tensors, input-gradient bwd over the generated half) included"
                           if use_perceptual() else
                           "the perceptual (Encoder2) branch of SURVEY 8 f1 is not included" if use_disc() else
                           "discriminator/perceptual branches (SURVEY 8 f1) not included"),
            "batch_per_gpu": B, "global_batch": B * world, "line_px": [64, 4 * GAN["Ts"]], "classes": GAN["C"],
            "target_chars": GAN["S"], "parallelism": f"dp{world}",
            "l2": "no explicit flush: the bf16 activations + gradients one step streams (~1.5 GB at B=16) exceed the "
                  "126 MB L2; weights stay cached, as in production",
            "noise": "NoiseInjection N(0,1) drawn in-kernel, re-seeded every step by a device-side counter",
            "batchnorm": {"peer": "recognizer BatchNorm statistics over the global batch: summed over the ranks inside "
                                  "the coefficient kernels through NVLink peer memory (14 in-kernel exchanges per step)",
                          "nccl": "recognizer BatchNorm statistics over the global batch: 14 small NCCL all-reduces per step",
                          "off": "recognizer BatchNorm statistics per rank"}[sync_bn],
            "execution": executed}


def gen_layers(T, n_in=208, dim=256):
    from bench import gen_conv_layers
    return gen_conv_layers(T, n_in, dim)


def cpu_lines_per_s(sample_B, reps):
    """The reference's CPU path for this step (oracle port: torch fp32 autograd through oracle/gen.py and
    oracle/hwr.py + F.ctc_loss + Adam on the generator), all host threads."""
    from oracle import disc as odisc, gen as ogen, hwr as ohwr, synth
    from handwriting_line_generation_b200 import CNNOnlyHWR, DiscriminatorAP, SpacedGenerator   # parameter containers only (CPU)
    torch.manual_seed(0)
    gmod = SpacedGenerator(GAN["C"], GAN["style"], GAN["dim"], n_style_trans=6, emb_dropout=False, append_style=True,
                           small=False)
    trainable = {n for n, _ in gmod.named_parameters()}
    gsd = {k: v.clone().requires_grad_(k in trainable) for k, v in gmod.state_dict().items()}
    hsd = {k: v.clone() for k, v in CNNOnlyHWR(GAN["C"], norm='batch').state_dict().items()}
    dsd = {k: v.clone() for k, v in DiscriminatorAP(64, use_low=True, use_med=True).state_dict().items()} if use_disc() else None
    params = [v for v in gsd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=2e-4, betas=(0.5, 0.999))
    content, style = synth.gen_case(GAN["Ts"], sample_B, GAN["C"], GAN["style"], 3)
    c, s = torch.from_numpy(content), torch.from_numpy(style)
    shapes = synth.gen_noise_shapes(GAN["Ts"], sample_B, GAN["dim"])
    T = GAN["Ts"] - 6
    tg = torch.randint(1, GAN["C"], (sample_B, GAN["S"]), dtype=torch.int32)
    il, tl = torch.full((sample_B,), T, dtype=torch.int32), torch.full((sample_B,), GAN["S"], dtype=torch.int32)
    times = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        noise = [torch.randn(sh) for sh in shapes]
        img = ogen.generator_forward(gsd, c, s, noise)
        lp = ohwr.hwr_forward(hsd, img, True, {})
        loss = W_CTC * torch.nn.functional.ctc_loss(lp, tg, il, tl)
        if dsd is not None:
            masks = {site: (torch.rand(sample_B, cm * 64) >= p).float() for site, p, cm in synth.DISC_SITES}
            upd = {}
            loss = loss + W_GEN * odisc.gen_loss(odisc.disc_forward(dsd, img, masks, training=True, update=upd))
            dsd.update(upd)                      # the spectral-norm vectors advance on every forward
        loss.backward()
        opt.step()
        if i:
            times.append(time.perf_counter() - t0)
    return sample_B / statistics.median(times), times


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    steps = min(args.steps, 12)
    t0 = time.time()
    sample_B = 4
    lps, times = cpu_lines_per_s(sample_B, steps)
    sample = (f"{len(times)} timed train steps on a {sample_B}-line quarter of the batch, torch fp32 on "
              f"{torch.get_num_threads()} host threads, {time.time() - t0:.1f}s")
    print(json.dumps({"impl": "reference", "metric": "GAN train-step lines/sec", "value": lps, "unit": "lines/s",
                      "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": config(GAN["B"], args.gpus, "oracle port on the host cores"),
                      "cpu_baseline": {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(),
                                       "kind": "port", "sample": sample},
                      "e2e": {"value": lps, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
          flush=True)


def main(args, rank, world, local_rank, load_peaks, ClockSampler):
    if args.impl == "reference":
        return run_reference(args, rank)
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import conv as hconv, dp, graphs
    from oracle import synth   # input builders only (numpy)

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl ours (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, Ts, S, C = GAN["B"], GAN["Ts"], GAN["S"], GAN["C"]
    T = Ts - 6
    torch.manual_seed(0)       # identical initial weights on every rank
    gen = pkg.SpacedGenerator(C, GAN["style"], GAN["dim"], n_style_trans=6, emb_dropout=False, append_style=True,
                              small=False).to(dev).train()
    hwr = pkg.CNNOnlyHWR(C, norm='batch').to(dev).train()
    for p in hwr.parameters():
        p.requires_grad_(False)            # hwr_frozen: no optimizer touches it; its wgrad is skipped
    disc = None
    if use_disc():
        disc = pkg.DiscriminatorAP(64, use_low=True, use_med=True).to(dev).train()   # IAM GAN config: dim 64, "use low"
        for p in disc.parameters():
            p.requires_grad_(False)        # 'gen' lesson: the discriminator only scores; its optimizer is not stepped
    enc, real = None, None
    if use_balanced():
        assert world == 1 and disc is not None, "HWG_BENCH_BALANCED: one GPU, with the discriminator branch"
        pkg.set_retain_graph(True)         # the generator is back-propagated through twice on one graph
    if use_perceptual() or use_balanced():
        enc = pkg.Encoder2(32).to(dev).train()      # the trainer never calls .eval() on it (:136-158): Dropout2d active
        real = torch.from_numpy(synth.hwr_case(B, 4 * Ts, 9000 + rank)).to(dev)
    # train-mode BatchNorm over the GLOBAL batch, as in the single-process reference: "peer" = in-kernel exchange over
    # NVLink peer memory (dp.PeerExchange), "nccl" = one NCCL all-reduce per layer and direction, "off" = per-rank
    sync_bn = os.environ.get("HWG_BENCH_SYNC_BN", DEFAULT_SYNC_BN) if world > 1 else "off"
    if sync_bn == "peer":
        try:
            hwr.sync_bn_group = dp.PeerExchange(dist.group.WORLD)
        except Exception as e:   # noqa: BLE001 - peer mapping refused on this box: same semantics through NCCL
            sys.stderr.write(f"[bench] PeerExchange unavailable ({e!r}); SyncBN through NCCL\n")
            sync_bn = "nccl"
        # every rank must take the same route
        route = torch.tensor([1.0 if sync_bn == "peer" else 0.0], device=dev)
        dist.all_reduce(route, op=dist.ReduceOp.MIN)
        if route.item() == 0:
            sync_bn = "nccl"
    if sync_bn == "nccl":
        hwr.sync_bn_group = dist.group.WORLD
    elif sync_bn not in ("off", "peer"):
        raise ValueError(f"HWG_BENCH_SYNC_BN={sync_bn!r}: expected peer, nccl or off")
    # flat fused optimizer: parameters / gradients / moments of the generator as slices of flat buffers; the
    # backward kernels add their gradients straight into the gradient buffer (gen._grad_sink), the all-reduce
    # buckets are slices of it, clip_grad_value_(2) + Adam + zero_grad is one launch
    opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
    gen._grad_sink = opt
    reducer = dp.GradReducer(gen.parameters(), flat=opt) if world > 1 else None
    if reducer is not None:
        gen._grad_ready_cb = reducer.mark_ready
    n_sets = 4
    host = []
    for i in range(n_sets):
        content, style = synth.gen_case(Ts, B, C, GAN["style"], 1000 * rank + i)
        tg = np.random.RandomState(7000 + 1000 * rank + i).randint(1, C, (B, S)).astype(np.int32)
        host.append(tuple(torch.from_numpy(a).pin_memory() for a in (content, style, tg)))
    devsets = [tuple(a.to(dev) for a in h) for h in host]
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    torch.manual_seed(1234 + rank)   # per-rank noise streams

    # the two critics of the generated image are independent until their image gradients meet: the discriminator
    # branch runs on a side stream next to recognizer + CTC (autograd replays each branch's backward on the stream of
    # its forward), so the small-grid launches of both — 1-D head, CTC chain, the discriminator's low-resolution
    # layers — share the GPU; captured as two parallel branches of the step's graph
    overlap = disc is not None and not os.environ.get("HWG_BENCH_NO_OVERLAP")
    side = torch.cuda.Stream() if overlap else None
    branch = {"parallel": overlap}         # the per-kernel roofline pass below times the launches one stream at a time

    def adversarial(img):                  # generator's adversarial loss, trainer/hw_with_style_trainer.py:810-821
        preds = disc(img)
        return -(W_GEN / len(preds)) * sum(p.mean() for p in preds)

    def train(c, s, tg):
        img = gen(c, s)
        par = branch["parallel"]
        if par:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                adv = adversarial(img)
        loss = W_CTC * pkg.CTCLoss(hwr(img), tg, il, tl)
        if par:
            main.wait_stream(side)
            loss = loss + adv
        elif disc is not None:
            loss = loss + adversarial(img)
        if enc is not None:
            loss = loss + W_PERC * enc.perceptual_loss(real, img)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()             # clip + Adam + gradient zeroing, one launch
        return loss

    def train_balanced(c, s, tg):
        img = gen(c, s)                                       # lesson 1: 'gen', no-step
        adv = adversarial(img)
        recog = W_CTC * pkg.CTCLoss(hwr(img), tg, il, tl)
        recog.backward(retain_graph=True)                     # trainer :312-323: the 'Recog' losses first, stashed
        opt.stash()
        adv.backward()                                        # :326-338: the rest of a no-step lesson, stashed
        opt.stash()
        perc = W_PERC * enc.perceptual_loss(real, gen(c, s))  # lesson 2: the 'auto' lesson's perceptual loss (:724-748)
        perc.backward()
        opt.balance(BALANCE_VAR_X)                            # :340-377
        opt.step()                                            # :381-391
        return adv.detach() + recog.detach() + perc.detach()

    if use_balanced():
        train = train_balanced                                # noqa: F811 - the opt-in step replaces the default one

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # eager steps: allocator warm-up, optimizer state, and the launch count of one step
    for i in range(2):
        train(*devsets[i])
    torch.cuda.synchronize()
    n0 = pkg._lib.launch_count()
    train(*devsets[0])
    launches_per_step = pkg._lib.launch_count() - n0
    graphed, executed = None, "eager launches"
    try:
        if os.environ.get("HWG_BENCH_NO_GRAPH"):
            raise RuntimeError("disabled by HWG_BENCH_NO_GRAPH")
        graphed = graphs.GraphedStep(train, list(devsets[0]), modules=[gen, hwr], warmup=3)
        executed = ("one CUDA graph per step (fixed shapes; forward, CTC, backward, "
                    + ("discriminator branch on a parallel stream, " if overlap else "")
                    + ("NCCL all-reduce on a side stream, " if world > 1 else "") + "Adam), replayed")
    except Exception as e:        # a capture failure must not lose the measurement: run the same step eagerly
        graphed = None
        executed = f"eager launches (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
        torch.cuda.synchronize()

    def step_device(i):
        if graphed is not None:
            return graphed(*devsets[i % n_sets])
        return train(*devsets[i % n_sets])

    def step_e2e(i):
        hc, hs, ht = host[i % n_sets]
        if graphed is not None:
            for dst, src in zip(graphed.static_in, (hc, hs, ht)):
                dst.copy_(src, non_blocking=True)       # H2D from pinned memory into the graph's input buffers
            graphed.graph.replay()
            loss = graphed.static_out
        else:
            loss = train(hc.to(dev, non_blocking=True), hs.to(dev, non_blocking=True), ht.to(dev, non_blocking=True))
        loss_host.copy_(loss.detach(), non_blocking=True)        # D2H of the step's loss

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return dp.max_over_ranks(e0.elapsed_time(e1), dev)

    W = max(3, args.warmup)
    for i in range(W):
        step_device(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t0 = time.time()
    ms = timed(step_device, args.steps)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    loss_value = float(loss_host.detach())
    peer_fault = None
    try:                                   # an in-kernel exchange that timed out sets a flag instead of hanging
        if hasattr(hwr.sync_bn_group, "fault"):
            peer_fault = int(hwr.sync_bn_group.fault.item())
    except Exception:                      # noqa: BLE001 - reporting only
        peer_fault = None

    # ---- rooflines: CUDA events around every convolution launch of a few eager steps (same kernels, same stream)
    prof = []
    branch["parallel"] = False
    gen.parallel_wgrad = False             # ... and with the generator's wgrad launches back on the main stream
    hconv.PROFILE = prof
    psteps = min(args.steps, 4)
    barrier()
    for i in range(psteps):
        train(*devsets[i % n_sets])
    barrier()
    hconv.PROFILE = None
    kern = {}
    for e0, e1, fl, kind, by in prof:
        k = kern.setdefault(kind, {"ms": 0.0, "launches": 0.0, "issued_flop": 0.0, "bytes": 0.0})
        k["ms"] += e0.elapsed_time(e1) / psteps
        k["launches"] += 1 / psteps
        k["issued_flop"] += fl / psteps
        k["bytes"] += by / psteps
    def leave():
        """A process group whose collectives were captured in a live CUDA graph does not tear down cleanly
        (destroy_process_group blocked until the launcher's timeout on the first 2-GPU run): synchronise, flush, exit."""
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            import sys
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        return leave()
    peaks = load_peaks()
    layers, _ = gen_layers(Ts)
    ms_step = ms / args.steps
    # algorithmic work per step by kernel: a layer's dgrad runs on the kernel that serves its fprop
    alg = {
        "conv_fprop_kernel": {"gflop": B * (2 * sum(l[1] for l in layers if l[3] == "conv_fprop_kernel") / 1e9
                                            + 2 * (HWR_GF_FWD_PER_LINE - HWR_GF_STEM_PER_LINE)
                                            + (2 * DISC_GF_FWD_PER_LINE if disc is not None else 0.0)),
                              "what": "fprop + dgrad of generator b0-b2, recognizer conv1-6 + 1-D head"
                                      + (" and every discriminator convolution" if disc is not None else "") + " (tcgen05)"},
        "conv_small_kernel": {"gflop": B * 2 * sum(l[1] for l in layers if l[3] == "conv_small_kernel") / 1e9,
                              "mb": B * 2 * sum(l[2] for l in layers if l[3] == "conv_small_kernel") / 1e6,
                              "what": "fprop + dgrad of generator b3-b4 (16-64 channels, HBM-bound)"},
        "conv_wgrad_kernel": {"gflop": B * sum(l[1] for l in layers) / 1e9,
                              "what": "wgrad of all generator convolutions (tcgen05 + staged-tile kernels)"},
    }

    # DRAM traffic per step by kernel from the committed ncu pass over one steady-state step (same command, B=16)
    traffic = {}
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_gan_train_r01.json")
    if os.path.exists(tp) and B == 16:
        names = {"conv_fprop_kernel": ("conv_fprop_kernel",), "conv_small_kernel": ("conv_small_kernel",),
                 "conv_wgrad_kernel": ("conv_wgrad_kernel", "wgrad_small_kernel")}
        for kind, prefixes in names.items():
            tot = sum((v["dram_read_mb"] + v["dram_write_mb"]) * 1e6 for n, v in json.load(open(tp))["kernels"].items()
                      if n.startswith(prefixes))
            traffic[kind] = int(tot)

    def roof(kind):
        k = kern[kind]
        a = alg.get(kind, {})
        if kind == "conv_small_kernel":
            ach = a["mb"] * 1e6 / (k["ms"] * 1e-3) / 1e9
            r = {"bound": "hbm", "kernel": kind, "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                 "frac": ach / peaks["hbm"], "peak_source": f"{peaks['src']} HBM copy bandwidth",
                 "algorithmic_mb_per_step": a["mb"]}
        else:
            gf = a.get("gflop", k["issued_flop"] / 1e9)
            ach = gf * 1e9 / (k["ms"] * 1e-3) / 1e12
            r = {"bound": "tensor", "kernel": kind, "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                 "frac": ach / peaks["tf_sust"],
                 "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside the step)",
                 "algorithmic_gflop_per_step": gf, "issued_gflop_per_step": k["issued_flop"] / 1e9}
        r.update({"covers": a.get("what"), "traffic": traffic.get(kind), "traffic_unit": "bytes per step (all launches "
                  "of this kernel; ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic_gan_train_r01.json)",
                  "launches_per_step": round(k["launches"]),
                  "kernel_ms_per_step": k["ms"], "share_of_step": k["ms"] / ms_step})
        return r

    top = max(kern, key=lambda kk: kern[kk]["ms"])
    lines = B * world
    cpu = None
    if world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        tb = time.time()
        lps, times = cpu_lines_per_s(4, 6)
        cpu = {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{len(times)} timed train steps on a 4-line quarter of the batch (T_s={Ts}), torch fp32, "
                         f"{time.time() - tb:.1f}s of CPU work"}
    h2d = int(Ts * B * C * 4 + B * GAN["style"] * 4 + B * S * 4)
    line = {
        "metric": "GAN train-step lines/sec", "value": lines / (ms_step * 1e-3), "unit": "lines/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(B, world, executed, sync_bn),
        "e2e": {"value": lines / (ms_e2e / args.steps * 1e-3), "unit": "lines/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks,
        "roofline": roof(top), "roofline_other_kernels": [roof(kk) for kk in kern if kk != top],
        "cpu_baseline": cpu, "final_loss": loss_value,
    }
    if peer_fault is not None:
        line["config"]["peer_exchange_timeouts"] = peer_fault      # 0 = every in-kernel exchange completed
    if world == 1 and not os.environ.get("HWG_BENCH_NO_EXTRAS"):
        try:   # the other configs, measured briefly in the same run (bench.py --workload gen_infer / hwr_train)
            import bench_hwr_train
            del gen, hwr, disc, opt, graphed
            torch.cuda.empty_cache()
            line["extra_workloads"] = bench_hwr_train.quick_train_numbers(dev, gen_lesson=False)
            line["extra_workloads"].update(quick_gen_infer(dev))
            line["extra_workloads"].update(quick_disc_lesson(dev))
        except Exception as e:   # never lose the headline over the extras
            line["extra_workloads"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    leave()


def quick_gen_infer(dev, steps=20):
    """BASELINE configs[1] (generator inference, batch 32), device-timed graph replays."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import graphs
    from oracle import synth
    B, Ts = 32, GAN["Ts"]
    torch.manual_seed(0)
    gen = pkg.SpacedGenerator(GAN["C"], GAN["style"], GAN["dim"], n_style_trans=6, emb_dropout=False,
                              append_style=True, small=False).to(dev).eval()
    content, style = synth.gen_case(Ts, B, GAN["C"], GAN["style"], 5)
    c, s = torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev)

    def fwd(c, s):
        with torch.no_grad():
            return gen(c, s)

    g = graphs.GraphedStep(fwd, [c, s], modules=[gen], warmup=3)
    for _ in range(3):
        g(c, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g(c, s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"gen_infer_B32": {"ms_per_step": ms, "lines_per_s": B / ms * 1e3,
                              "what": "BASELINE configs[1]: SpacedGenerator inference, 32 lines of 64x1024 px, one "
                                      "replayed CUDA graph per step (full line: bench.py --workload gen_infer)"}}


def quick_disc_lesson(dev, steps=20):
    """The 'disc' lesson of the GAN cycle (trainer/hw_with_style_trainer.py:785-804): generator forward without
    gradients, DiscriminatorAP on real || fake rows (2 x 16 lines), hinge loss, backward to all discriminator
    weights (tcgen05 wgrad, spectral-norm backward), clip + Adam on the discriminator; one replayed CUDA graph."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import graphs
    from oracle import synth
    B, Ts = GAN["B"], GAN["Ts"]
    torch.manual_seed(0)
    gen = pkg.SpacedGenerator(GAN["C"], GAN["style"], GAN["dim"], n_style_trans=6, emb_dropout=False,
                              append_style=True, small=False).to(dev).eval()
    disc = pkg.DiscriminatorAP(64, use_low=True, use_med=True).to(dev).train()
    opt = pkg.FlatAdam([p for p in disc.parameters() if p.requires_grad], lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
    content, style = synth.gen_case(Ts, B, GAN["C"], GAN["style"], 5)
    c, s = torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev)
    real = torch.from_numpy(synth.hwr_case(B, 4 * Ts, 11)).to(dev)

    def step(c, s, real):
        with torch.no_grad():
            fake = gen(c, s)
        preds = disc(torch.cat((real, fake), 0))
        loss = sum(torch.relu(1.0 - p[:B]).mean() + torch.relu(1.0 + p[B:]).mean() for p in preds) / len(preds)
        loss.backward()
        opt.step()
        return loss

    g = graphs.GraphedStep(step, [c, s, real], modules=[gen], warmup=3)
    for _ in range(3):
        g(c, s, real)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g(c, s, real)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"disc_lesson_B16": {"ms_per_step": ms, "lines_per_s": B / ms * 1e3, "loss": float(g.static_out),
                                "what": "GAN 'disc' lesson: generator forward (no grad) + DiscriminatorAP fwd+bwd on "
                                        "16 real + 16 generated lines, hinge loss, Adam on the discriminator; lines/s "
                                        "counts the 16 lines of the batch; one replayed CUDA graph"}}
