"""bench.py default workload: ONE balanced optimizer step of the HWWithStyle GAN curriculum (BASELINE configs[2] / [3]) on
the SURVEY.md §8 rows a1-a12, f1, f2 — the step the reference takes with `balance_loss` (trainer/hw_with_style_trainer.py:
300-391) over the pair of curriculum lessons ["no-step 'gen'", 'auto'] (configs/cf_IAMslant_*sMG.json:85-96), restricted to
the components this repo builds:

  lesson 1, 'gen' (no-step)       image = SpacedGenerator(spaced text, style)                    (trainer :577, pure_gen.py:42-50)
                                  genRecog  = 1e-4 * CTCLoss(CNNOnlyHWR(image), text)            (:760-764; frozen HWR, train-mode BN)
                                  generator = -mean DiscriminatorAP(image)                       (:810-821; frozen D, train mode)
                                  genRecog.backward(retain_graph=True) -> [all-reduce] -> stash  (:312-323)
                                  generator.backward()                 -> [all-reduce] -> stash  (:326-338)
  lesson 2, 'auto' (its perceptual part)
                                  recon = SpacedGenerator(...);  0.5 * L1(Encoder2([real ; recon]) halves)   (:724-748)
                                  backward -> [all-reduce]
  optimizer step                  per-tensor gradient balancing of the two stashed sets (:340-377, balance_var_x),
                                  clip_grad_value_(2) (:381), Adam lr 2e-4 betas (0.5, 0.999) (:383-391)

Two generator forwards, three generator backwards, recognizer + CTC, discriminator and perceptual-encoder forward/backward
per step.  Not in the step (outside the built path): the style extractor and spacer of the 'auto' lesson —
the second lesson re-uses the step's style vectors and spaced text; `HWG_BENCH_STEP=gen_only` selects round 1's reduced
step (one backward over the summed 'gen' losses).

Strong scaling (BASELINE configs[3]): GLOBAL batch 128 lines of 64x1024 px at every N -> 128 / 64 / 32 / 16 lines per GPU;
`HWG_BENCH_B=<lines per GPU>` fixes the per-GPU batch instead (weak scaling, stated in the line).
"""
import json
import os
import statistics
import sys
import time
import traceback

import numpy as np
import torch

GAN = dict(GLOBAL_B=128, Ts=256, S=40, C=80, style=128, dim=256)
HWR_GF_FWD_PER_LINE = 24.661      # SURVEY.md §8a (a10), forward conv GFLOP per 64x1024 line
HWR_GF_STEM_PER_LINE = 0.075      # conv0 (fused stem kernel, not a tensor-core launch)
DISC_GF_FWD_PER_LINE = 11.295     # SURVEY.md §8d / Appendix D, DiscriminatorAP forward conv GFLOP per 64x1024 line
DISC_GF_STEM_PER_LINE = 2 * 58 * 1024 * 64 * 49 / 1e9   # its 7x7 in_conv: forward on hwg_stem_conv (mma.sync), not a tcgen05 launch
ENC_GF_FWD_PER_LINE = 1.493       # SURVEY.md Appendix D, Encoder2(32) forward conv GFLOP per 64x1024 line
W_CTC, W_GEN, W_PERC = 1e-4, 1.0, 0.5   # loss_weights genRecog / generator / perceptual of the IAM GAN config (:54-62)
BALANCE_VAR_X = [0.6, 0.5]        # config :100 `balance_var_x` [0.6, 0.5, 0.4, 0.75]: the entries of the two sets stashed here
DEFAULT_SYNC_BN = "peer"
TRAFFIC_JSON = "traffic_gan_train_r02.json"


def step_kind():
    k = os.environ.get("HWG_BENCH_STEP", "balanced")
    if k not in ("balanced", "gen_only"):
        raise ValueError(f"HWG_BENCH_STEP={k!r}: expected balanced or gen_only")
    return k


def per_gpu_batch(world):
    if os.environ.get("HWG_BENCH_B"):
        return int(os.environ["HWG_BENCH_B"]), "weak"
    assert GAN["GLOBAL_B"] % world == 0
    return GAN["GLOBAL_B"] // world, "strong"


def config(B, world, executed, sync_bn="off"):
    if step_kind() == "balanced":
        what = ("HWWithStyle GAN balanced optimizer step (BASELINE configs[2]/[3]: generator + discriminator_ap + HWR CTC + "
                "autoencoder perceptual loss), as the reference curriculum takes it with balance_loss over the lesson pair "
                "[no-step 'gen', 'auto']: pure_gen generator fwd; frozen cnn_only_hwr fwd + CTC and frozen discriminator_ap fwd "
                "(train-mode BatchNorm / spectral norm / Dropout2d); genRecog and generator losses back-propagated SEPARATELY "
                "through the generator (retain_graph) and stashed; second generator fwd + frozen Encoder2(32) perceptual L1 "
                "against real lines, backward; per-tensor gradient balancing of the two stashed sets (trainer :340-377), "
                "clip_grad_value_(2), Adam.  2 generator forwards + 3 generator backwards per step.  The 'auto' lesson's style "
                "extractor / spacer (SURVEY 8 f3/f4) are outside the built path: the second lesson re-uses the step's styles")
    else:
        what = ("HWWithStyle GAN 'gen' lesson made self-contained (round-1 step, HWG_BENCH_STEP=gen_only): generator fwd+bwd, "
                "frozen cnn_only_hwr + CTC, frozen discriminator_ap, ONE backward over 1e-4*CTC - mean D(fake), clip + Adam")
    return {"workload": what, "batch_per_gpu": B, "global_batch": B * world, "line_px": [64, 4 * GAN["Ts"]],
            "classes": GAN["C"], "target_chars": GAN["S"], "parallelism": f"dp{world}",
            "lines_counted": "the B lines of the step's batch once (both lessons run on them)",
            "l2": "no explicit flush: the bf16 activations + gradients one step streams (> 1 GB per 16 lines) exceed the "
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


# ----------------------------------------------------------------------------------------------------------------------
# baselines: the stock-PyTorch port of the same step (oracle/gan_step.py) on the host cores / on the GPU with cuDNN
# ----------------------------------------------------------------------------------------------------------------------
def cpu_lines_per_s(sample_B, reps):
    """The reference's CPU path for this step (stock torch fp32 autograd: oracle/gan_step.py), all host threads."""
    from oracle.gan_step import PortStep
    st = PortStep("cpu", sample_B, GAN["Ts"], GAN["C"], GAN["S"], GAN["style"], GAN["dim"], step=step_kind())
    times = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        st()
        if i:
            times.append(time.perf_counter() - t0)
    return sample_B / statistics.median(times), times


def gpu_baseline(dev, B, steps):
    """SURVEY §8d 'the real bar': the same step with stock PyTorch on the same B200 — cuDNN convolutions (TF32 allowed, the
    torch default the reference would run with), ATen norm / pool / CTC kernels, torch.optim.Adam — eager, and replayed as
    one torch.cuda.CUDAGraph where the stock ops capture."""
    from oracle.gan_step import PortStep
    out = {"kind": "torch-cudnn", "batch": B, "unit": "lines/s",
           "what": "oracle/gan_step.py (stock torch autograd port of the same step, fp32 tensors, cudnn.allow_tf32 default) on "
                   "cuda:0, CUDA-event timed"}

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    torch.backends.cudnn.benchmark = True
    try:
        st = PortStep(dev, B, GAN["Ts"], GAN["C"], GAN["S"], GAN["style"], GAN["dim"], step=step_kind())
        ms = timed(st, steps)
        out.update(value=B / ms * 1e3, ms_per_step=ms, eager={"ms_per_step": ms, "lines_per_s": B / ms * 1e3})
        del st
    except Exception as e:            # noqa: BLE001 - a baseline failure must not lose the headline
        out["eager"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    torch.cuda.empty_cache()
    try:
        st = PortStep(dev, B, GAN["Ts"], GAN["C"], GAN["S"], GAN["style"], GAN["dim"], capturable=True, step=step_kind())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                st()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st()
        ms = timed(g.replay, steps)
        out["cuda_graph"] = {"ms_per_step": ms, "lines_per_s": B / ms * 1e3}
        if "value" not in out or B / ms * 1e3 > out["value"]:
            out.update(value=B / ms * 1e3, ms_per_step=ms)
        del g, st
    except Exception as e:            # noqa: BLE001
        out["cuda_graph"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        try:
            torch.cuda.synchronize()
        except Exception:             # noqa: BLE001
            pass
    torch.backends.cudnn.benchmark = False
    torch.cuda.empty_cache()
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    steps = min(args.steps, 16)               # bounded: ~11 s of CPU work at 0.65 s per 4-line step
    t0 = time.time()
    sample_B = 4
    B, scaling = per_gpu_batch(max(1, args.gpus))
    lps, times = cpu_lines_per_s(sample_B, steps)
    sample = (f"{len(times)} timed optimizer steps on a {sample_B}-line sample of the batch (64x{4 * GAN['Ts']} px), stock "
              f"torch fp32 on {torch.get_num_threads()} host threads, {time.time() - t0:.1f}s")
    print(json.dumps({"impl": "reference", "metric": "GAN train-step lines/sec", "value": lps, "unit": "lines/s",
                      "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": scaling,
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": config(B, max(1, args.gpus), "stock-torch port of the step on the host cores (no GPU)"),
                      "cpu_baseline": {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(),
                                       "kind": "port", "sample": sample},
                      "e2e": {"value": lps, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
          flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# the product step
# ----------------------------------------------------------------------------------------------------------------------
class GanStep:
    """Models, flat optimizer, reducer and the step function (shared with tools/step_runner.py and the tests)."""

    def __init__(self, dev, B, world=1, rank=0, dist=None, kind=None, sync_bn="off", overlap=True, Ts=None):
        import handwriting_line_generation_b200 as pkg
        from handwriting_line_generation_b200 import dp
        self.pkg, self.dev, self.B, self.world, self.kind = pkg, dev, B, world, kind or step_kind()
        C = GAN["C"]
        torch.manual_seed(0)       # identical initial weights on every rank
        self.gen = pkg.SpacedGenerator(C, GAN["style"], GAN["dim"], n_style_trans=6, emb_dropout=False, append_style=True,
                                       small=False).to(dev).train()
        self.hwr = pkg.CNNOnlyHWR(C, norm='batch').to(dev).train()
        for p in self.hwr.parameters():
            p.requires_grad_(False)            # hwr_frozen: no optimizer touches it; its wgrad is skipped
        self.disc = pkg.DiscriminatorAP(64, use_low=True, use_med=True).to(dev).train()   # IAM GAN config: dim 64, "use low"
        for p in self.disc.parameters():
            p.requires_grad_(False)            # generator-side lessons: the discriminator only scores
        self.enc = None
        if self.kind == "balanced":
            pkg.set_retain_graph(True)         # the generator is back-propagated through twice on one graph (:303-325)
            self.enc = pkg.Encoder2(32).to(dev).train()   # the trainer never calls .eval() on it (:136-158): Dropout2d active
            for p in self.enc.parameters():
                p.requires_grad_(False)
        # train-mode BatchNorm over the GLOBAL batch, as in the single-process reference: "peer" = in-kernel exchange over
        # NVLink peer memory (dp.PeerExchange), "nccl" = one NCCL all-reduce per layer and direction, "off" = per-rank
        self.sync_bn = sync_bn if world > 1 else "off"
        if self.sync_bn == "peer":
            try:
                self.hwr.sync_bn_group = dp.PeerExchange(dist.group.WORLD)
            except Exception as e:   # noqa: BLE001 - peer mapping refused on this box: same semantics through NCCL
                sys.stderr.write(f"[bench] PeerExchange unavailable ({e!r}); SyncBN through NCCL\n")
                self.sync_bn = "nccl"
            route = torch.tensor([1.0 if self.sync_bn == "peer" else 0.0], device=dev)     # every rank takes the same route
            dist.all_reduce(route, op=dist.ReduceOp.MIN)
            if route.item() == 0:
                self.sync_bn = "nccl"
        if self.sync_bn == "nccl":
            self.hwr.sync_bn_group = dist.group.WORLD
        elif self.sync_bn not in ("off", "peer"):
            raise ValueError(f"HWG_BENCH_SYNC_BN={self.sync_bn!r}: expected peer, nccl or off")
        # flat fused optimizer: parameters / gradients / moments of the generator as slices of flat buffers; the backward
        # kernels add their gradients straight into the gradient buffer (gen._grad_sink), the all-reduce buckets are slices
        # of it, stash / balance / clip + Adam + zero are flat launches
        self.opt = pkg.FlatAdam(self.gen.parameters(), lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
        self.gen._grad_sink = self.opt
        # gen_only step: bucketed all-reduce from grad-ready hooks (dp.GradReducer); balanced step: one all-reduce per gradient
        # set on a communication stream (_allreduce)
        self.reducer = dp.GradReducer(self.gen.parameters(), flat=self.opt) if (world > 1 and self.kind != "balanced") else None
        if self.reducer is not None:
            self.gen._grad_ready_cb = self.reducer.mark_ready
        self.comm = torch.cuda.Stream(device=dev) if world > 1 else None
        T = (Ts or GAN["Ts"]) - 6
        self.il = torch.full((B,), T, dtype=torch.int32, device=dev)
        self.tl = torch.full((B,), GAN["S"], dtype=torch.int32, device=dev)
        # the two critics of the generated image are independent: the discriminator's forward runs on a side stream next
        # to recognizer + CTC (autograd replays each branch's backward on the stream of its forward)
        self.parallel = bool(overlap) and not os.environ.get("HWG_BENCH_NO_OVERLAP")
        self.s2 = self.side = torch.cuda.Stream(device=dev)
        self.s3 = torch.cuda.Stream(device=dev)

    def adversarial(self, img):                # generator's adversarial loss, trainer :810-821
        preds = self.disc(img)
        return -(W_GEN / len(preds)) * sum(p.mean() for p in preds)

    def _reduce(self):
        if self.reducer is not None:
            self.reducer.finish()              # gradient all-reduce (average over the ranks) of what this backward produced

    def _critics(self, img, tg):
        pkg = self.pkg
        if self.parallel:
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                adv = self.adversarial(img)
        recog = W_CTC * pkg.CTCLoss(self.hwr(img), tg, self.il, self.tl)
        if self.parallel:
            main.wait_stream(self.side)
        else:
            adv = self.adversarial(img)
        return recog, adv

    def train_gen_only(self, c, s, tg, real=None):
        recog, adv = self._critics(self.gen(c, s), tg)
        loss = recog + adv
        loss.backward()
        self._reduce()
        self.opt.step()
        return loss

    def _allreduce(self, buf):
        """Gradient all-reduce (average over the ranks) of one whole gradient set on the communication stream, ordered after
        the stream that produced it; it overlaps whatever the step issues next (the balancing is nonlinear, so every set
        is reduced before it)."""
        if self.world == 1:
            return
        import torch.distributed as dist
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            buf.mul_(1.0 / self.world)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def train_balanced(self, c, s, tg, real):
        """Three independent chains on three streams; each backward pass leaves its gradient set in its own buffer
        (FlatAdam.sink: the trainer's backward -> clone into saved_grad -> zero, :312-338, without the copies):
          main  generator fwd -> recognizer fwd -> CTC -> backward (recognizer, generator)                      -> stash slot 0
          s2    (after the generator fwd) discriminator fwd -> backward (discriminator; generator on main)      -> stash slot 1
          s3    (after the generator fwd) generator fwd #2 -> Encoder2 perceptual loss -> backward (both)       -> main buffer
        then all-reduces (N > 1), balancing, clip + Adam on the main stream."""
        opt, gen = self.opt, self.gen
        main = torch.cuda.current_stream()
        par = self.parallel
        img = gen(c, s)                                            # lesson 1: 'gen', no-step (trainer :577)
        if par:
            self.s2.wait_stream(main)
            self.s3.wait_stream(main)                              # (the weight re-pack of this step ran inside gen fwd #1)
            with torch.cuda.stream(self.s2):
                adv = self.adversarial(img)                        # :810-821
            with torch.cuda.stream(self.s3):                       # lesson 2: 'auto''s perceptual loss (:724-748)
                perc = W_PERC * self.enc.perceptual_loss(real, gen(c, s))
        recog = W_CTC * self.pkg.CTCLoss(self.hwr(img), tg, self.il, self.tl)      # :760-764
        if not par:
            adv = self.adversarial(img)
        gen._grad_sink = opt.sink(0)
        recog.backward(retain_graph=True)                          # :312-323: the 'Recog' losses first -> first stashed set
        self._allreduce(opt.sink(0).buf)
        gen._grad_sink = opt.sink(1)
        adv.backward()                                             # :326-338: the rest of a no-step lesson -> second set
        self._allreduce(opt.sink(1).buf)
        gen._grad_sink = opt
        if par:
            with torch.cuda.stream(self.s3):
                perc.backward()
                self._allreduce(opt.flat_g)
            main.wait_stream(self.s2)
            main.wait_stream(self.s3)
        else:
            perc = W_PERC * self.enc.perceptual_loss(real, gen(c, s))
            perc.backward()
            self._allreduce(opt.flat_g)
        if self.world > 1:
            main.wait_stream(self.comm)
        opt.balance(BALANCE_VAR_X)                                 # :340-377
        opt.step()                                                 # :381-391: clip + Adam + gradient zeroing, one launch
        return recog.detach() + adv.detach() + perc.detach()

    def train(self, *a):
        return self.train_balanced(*a) if self.kind == "balanced" else self.train_gen_only(*a)

    def conv_gflop(self, layers):
        """Algorithmic conv GFLOP per step by kernel (reference layer counts): a layer's dgrad runs on the kernel
        that serves its fprop.  Encoder2's launches are added from their launch tags (enc_gflop)."""
        B = self.B
        nf, nb = (2, 3) if self.kind == "balanced" else (1, 1)
        g_tc = sum(l[1] for l in layers if l[3] == "conv_fprop_kernel") / 1e9
        g_small = sum(l[1] for l in layers if l[3] == "conv_small_kernel") / 1e9
        mb_small = sum(l[2] for l in layers if l[3] == "conv_small_kernel") / 1e6
        return {
            "conv_fprop_kernel": {"gflop": B * ((nf + nb) * g_tc + 2 * (HWR_GF_FWD_PER_LINE - HWR_GF_STEM_PER_LINE)
                                                + 2 * DISC_GF_FWD_PER_LINE - DISC_GF_STEM_PER_LINE),
                                  "what": f"fprop x{nf} + dgrad x{nb} of generator b0-b2, fprop + dgrad of recognizer conv1-6 + 1-D "
                                          "head and of every discriminator convolution (the 7x7 stem: dgrad only, its forward "
                                          "runs on stem_conv_kernel)"
                                          + (", Encoder2 layers with >= 64 channels" if self.enc is not None else "") + " (tcgen05)"},
            "conv_small_kernel": {"gflop": B * (nf + nb) * g_small, "mb": B * (nf + nb) * mb_small,
                                  "what": f"fprop x{nf} + dgrad x{nb} of generator b3-b4 (16-64 channels, HBM-bound)"
                                          + (", Encoder2's 16-32 channel layers" if self.enc is not None else "")},
            "conv_wgrad_kernel": {"gflop": B * nb * sum(l[1] for l in layers) / 1e9,
                                  "what": f"wgrad x{nb} of all generator convolutions (tcgen05 + staged-tile kernels)"},
        }


def main(args, rank, world, local_rank, load_peaks, ClockSampler):
    if args.impl == "reference":
        return run_reference(args, rank)
    from handwriting_line_generation_b200 import conv as hconv, dp, graphs
    import bench_inputs as synth   # input builders (numpy)

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl ours (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, scaling = per_gpu_batch(world)
    Ts, S, C = GAN["Ts"], GAN["S"], GAN["C"]
    st = GanStep(dev, B, world, rank, dist, sync_bn=os.environ.get("HWG_BENCH_SYNC_BN", DEFAULT_SYNC_BN))
    pkg, gen, hwr, opt = st.pkg, st.gen, st.hwr, st.opt
    train = st.train
    n_sets = 4
    host = []
    for i in range(n_sets):
        content, style = synth.gen_case(Ts, B, C, GAN["style"], 1000 * rank + i)
        tg = np.random.RandomState(7000 + 1000 * rank + i).randint(1, C, (B, S)).astype(np.int32)
        arrs = [content, style, tg]
        if st.kind == "balanced":
            arrs.append(synth.hwr_case(B, 4 * Ts, 9000 + 1000 * rank + i))      # the 'auto' lesson's real lines
        host.append(tuple(torch.from_numpy(a).pin_memory() for a in arrs))
    devsets = [tuple(a.to(dev) for a in h) for h in host]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    torch.manual_seed(1234 + rank)   # per-rank noise streams

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

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return dp.max_over_ranks(e0.elapsed_time(e1), dev)

    # eager number (no CUDA graph): what a trainer driving the modules call by call sees, host launch overhead included
    esteps = min(args.steps, 10)
    ms_eager = timed(lambda i: train(*devsets[i % n_sets]), esteps) / esteps

    graphed, executed = None, "eager launches"
    if os.environ.get("HWG_BENCH_NO_GRAPH"):
        executed = "eager launches (HWG_BENCH_NO_GRAPH)"
    else:
        try:
            graphed = graphs.GraphedStep(train, list(devsets[0]), modules=[gen, hwr], warmup=3)
            executed = ("one CUDA graph per step (fixed shapes; every forward, CTC, the three backward passes, "
                        + ("as three parallel branches (recognizer / discriminator / second lesson), " if st.parallel else "")
                        + ("one NCCL all-reduce per gradient set on a communication stream, " if world > 1 else "")
                        + "balancing, Adam), replayed")
        except Exception:             # noqa: BLE001 - a failed capture poisons the process: say why, then run eagerly afresh
            traceback.print_exc()
            sys.stderr.write("[bench] CUDA-graph capture failed (above); restarting with HWG_BENCH_NO_GRAPH=1\n")
            sys.stderr.flush()
            if world == 1:
                os.environ["HWG_BENCH_NO_GRAPH"] = "1"
                os.execv(sys.executable, [sys.executable] + sys.argv)
            raise

    def step_device(i):
        if graphed is not None:
            return graphed(*devsets[i % n_sets])
        return train(*devsets[i % n_sets])

    def step_e2e(i):
        h = host[i % n_sets]
        if graphed is not None:
            for dst, src in zip(graphed.static_in, h):
                dst.copy_(src, non_blocking=True)       # H2D from pinned memory into the graph's input buffers
            graphed.graph.replay()
            loss = graphed.static_out
        else:
            loss = train(*[a.to(dev, non_blocking=True) for a in h])
        loss_host.copy_(loss.detach(), non_blocking=True)        # D2H of the step's loss

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

    # ---- rank-sync proof: after all those optimizer steps every rank must hold bit-identical parameters
    in_sync = None
    if world > 1:
        bits = opt.flat_p.view(torch.int32).to(torch.int64)
        chk = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
        allchk = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        in_sync = all(torch.equal(allchk[0], c) for c in allchk[1:])

    # ---- rooflines: CUDA events around every convolution launch of a few eager steps (same kernels, same stream)
    prof = []
    st.parallel = False                    # one stream at a time for the per-kernel pass
    gen.parallel_wgrad = False             # ... and with the generator's wgrad launches back on the main stream
    enc_profile = {}
    if st.enc is not None:
        eprof = []
        hconv.PROFILE = eprof
        fake = torch.rand_like(devsets[0][3]).requires_grad_()
        st.enc.perceptual_loss(devsets[0][3], fake).backward()
        hconv.PROFILE = None
        for _, _, fl, kind, by, (_n, _ho, _wo, cin, cout, ntaps) in eprof:
            if ntaps == 5 and 16 in (cin, cout):
                fl *= 25.0 / 80.0
            v = enc_profile.setdefault(kind, {"gflop": 0.0, "mb": 0.0})
            v["gflop"] += fl / 1e9
            v["mb"] += by / 1e6
    hconv.PROFILE = prof
    psteps = min(args.steps, 4)
    barrier()
    for i in range(psteps):
        train(*devsets[i % n_sets])
    barrier()
    hconv.PROFILE = None
    kern = {}
    if os.environ.get("HWG_BENCH_DUMP_CONV") and rank == 0:      # development aid: per-launch table of the profiled steps
        rows = [dict(kind=kind, geom=list(geo[0]) if geo else None, ms=e0.elapsed_time(e1), gflop=fl / 1e9, mb=by / 1e6)
                for e0, e1, fl, kind, by, *geo in prof]
        json.dump({"batch": B, "steps": psteps, "launches": rows}, open(os.environ["HWG_BENCH_DUMP_CONV"], "w"))
    for e0, e1, fl, kind, by, *_ in prof:
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
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        return leave()
    peaks = load_peaks()
    layers, _ = gen_layers(Ts)
    ms_step = ms / args.steps
    alg = st.conv_gflop(layers)
    if st.enc is not None:
        # Encoder2 (SURVEY Appendix D, 1.493 GF/line forward): forward on [real ; recon] = 2B lines, input-gradient backward over
        # the recon half only.  Which kernel serves a layer is decided per launch, so its share is taken from the launches of
        # one perceptual-loss pass run alone: algorithmic FLOPs = the launch's MACs with the stem's 5 taps x 16 shifted channels
        # counted as the reference's 5x5 taps (25 of the 80 issued), activation bytes = input read once + output written once
        for kind, v in enc_profile.items():
            alg[kind]["gflop"] += v["gflop"]
            if "mb" in alg[kind]:
                alg[kind]["mb"] += v["mb"]

    # DRAM traffic per step by kernel from the committed ncu pass over one steady-state step (same command, same batch)
    traffic = {}
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", TRAFFIC_JSON)
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("batch") == B and tj.get("step", "balanced") == st.kind:
            names = {"conv_fprop_kernel": ("conv_fprop_kernel",), "conv_small_kernel": ("conv_small_kernel",),
                     "conv_wgrad_kernel": ("conv_wgrad_kernel", "wgrad_small_kernel")}
            for kind, prefixes in names.items():
                tot = sum((v["dram_read_mb"] + v["dram_write_mb"]) * 1e6 for n, v in tj["kernels"].items()
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
                  f"of this kernel; ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/{TRAFFIC_JSON})",
                  "launches_per_step": round(k["launches"]),
                  "kernel_ms_per_step": k["ms"], "share_of_step": k["ms"] / ms_step})
        return r

    top = max(kern, key=lambda kk: kern[kk]["ms"])
    lines = B * world
    h2d = int(sum(a.numel() * a.element_size() for a in host[0]))
    line = {
        "metric": "GAN train-step lines/sec", "value": lines / (ms_step * 1e-3), "unit": "lines/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(B, world, executed, st.sync_bn),
        "e2e": {"value": lines / (ms_e2e / args.steps * 1e-3), "unit": "lines/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks,
        "roofline": roof(top), "roofline_other_kernels": [roof(kk) for kk in kern if kk != top],
        "eager_ms_per_step": ms_eager, "eager_lines_per_s": lines / (ms_eager * 1e-3),
        "final_loss": loss_value,
    }
    if in_sync is not None:
        line["ranks_in_sync"] = bool(in_sync)       # bit-identical flat parameter buffers on every rank after the run
    if peer_fault is not None:
        line["config"]["peer_exchange_timeouts"] = peer_fault      # 0 = every in-kernel exchange completed
    if world == 1:
        # free the product's memory, then the two baselines of the same step
        pkg_ = pkg
        del graphed, st, gen, hwr, opt, train, devsets
        torch.cuda.empty_cache()
        if not os.environ.get("HWG_BENCH_NO_GPU_BASELINE"):
            line["gpu_baseline"] = gpu_baseline(dev, B, min(args.steps, 10))
        line["cpu_baseline"] = None
        if not os.environ.get("HWG_BENCH_NO_CPU_BASELINE"):
            torch.set_num_threads(os.cpu_count() or 1)
            tb = time.time()
            lps, times = cpu_lines_per_s(4, 14)      # ~10 s of CPU work (0.65 s per 4-line step on the box's 16 threads)
            line["cpu_baseline"] = {"value": lps, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{len(times)} timed optimizer steps on a 4-line sample of the batch (T_s={Ts}), "
                                              f"stock torch fp32, {time.time() - tb:.1f}s of CPU work"}
        if not os.environ.get("HWG_BENCH_NO_EXTRAS"):
            try:   # the other configs, measured briefly in the same run (bench.py --workload gen_infer / hwr_train)
                import bench_hwr_train
                line["extra_workloads"] = bench_hwr_train.quick_train_numbers(dev, gen_lesson=False)
                line["extra_workloads"].update(quick_gen_infer(dev))
                line["extra_workloads"].update(quick_disc_lesson(dev))
                line["extra_workloads"].update(quick_step_b16(dev))
            except Exception as e:   # never lose the headline over the extras
                line["extra_workloads"] = {"error": repr(e)}
            try:   # the curriculum's 7-lesson cycle with every module on the drop-ins, driven eagerly (bench_cycle.py)
                import bench_cycle
                torch.cuda.empty_cache()
                line["extra_workloads"].update(bench_cycle.measure(dev, 16, cycles=4, warmup=2))
            except Exception as e:   # noqa: BLE001
                line["extra_workloads"]["cycle_B16"] = {"error": repr(e)[:300]}
            finally:
                pkg_.set_retain_graph(False)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    leave()


def _time_graph(g, ins, steps):
    for _ in range(3):
        g(*ins)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g(*ins)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def quick_step_b16(dev, steps=20):
    """The same balanced step at 16 lines on one GPU — the per-GPU share of the 8-GPU strong-scaling point."""
    from handwriting_line_generation_b200 import graphs
    import bench_inputs as synth
    B, Ts = 16, GAN["Ts"]
    st = GanStep(dev, B)
    content, style = synth.gen_case(Ts, B, GAN["C"], GAN["style"], 5)
    ins = [torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev),
           torch.from_numpy(np.random.RandomState(3).randint(1, GAN["C"], (B, GAN["S"])).astype(np.int32)).to(dev)]
    if st.kind == "balanced":
        ins.append(torch.from_numpy(synth.hwr_case(B, 4 * Ts, 11)).to(dev))
    for _ in range(2):
        st.train(*ins)
    g = graphs.GraphedStep(st.train, ins, modules=[st.gen, st.hwr], warmup=3)
    ms = _time_graph(g, ins, steps)
    return {"gan_step_B16": {"ms_per_step": ms, "lines_per_s": B / ms * 1e3,
                             "what": "the headline step at 16 lines on one GPU (per-GPU batch of the 8-GPU point), one replayed "
                                     "CUDA graph"}}


def quick_gen_infer(dev, steps=20):
    """BASELINE configs[1] (generator inference, batch 32), device-timed graph replays."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import graphs
    import bench_inputs as synth
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
    ms = _time_graph(g, [c, s], steps)
    return {"gen_infer_B32": {"ms_per_step": ms, "lines_per_s": B / ms * 1e3,
                              "what": "BASELINE configs[1]: SpacedGenerator inference, 32 lines of 64x1024 px, one "
                                      "replayed CUDA graph per step (full line: bench.py --workload gen_infer)"}}


def quick_disc_lesson(dev, steps=20):
    """The 'disc' lesson of the GAN cycle (trainer/hw_with_style_trainer.py:785-804): generator forward without
    gradients, DiscriminatorAP on real || fake rows (2 x 16 lines), hinge loss, backward to all discriminator
    weights (tcgen05 wgrad, spectral-norm backward), clip + Adam on the discriminator; one replayed CUDA graph."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import graphs
    import bench_inputs as synth
    B, Ts = 16, GAN["Ts"]
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
    ms = _time_graph(g, [c, s, real], steps)
    return {"disc_lesson_B16": {"ms_per_step": ms, "lines_per_s": B / ms * 1e3, "loss": float(g.static_out.detach()),
                                "what": "GAN 'disc' lesson: generator forward (no grad) + DiscriminatorAP fwd+bwd on "
                                        "16 real + 16 generated lines, hinge loss, Adam on the discriminator; lines/s "
                                        "counts the 16 lines of the batch; one replayed CUDA graph"}}
