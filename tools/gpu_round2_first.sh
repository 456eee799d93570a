#!/bin/bash
# First GPU call of the next round (one GPU; typically 6-8 min, every step under its own timeout):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'  Everything below was written after round 1's GPU budget was spent.
# (1) the parked GPU tests (trainer-level gen lesson, flat gradient balancing, Encoder2 / perceptual loss, DTW alignment);
# (2) tools/halo_probe.cu: do shifted UMMA views of one swizzled halo tile read the right pixels, and with which descriptor;
# (3) the halo-mode main loop of conv_fprop_kernel (HWG_CONV_HALO=1|2, HWG_CONV_HALO_BO per the probe): numerics through the
#     existing conv / discriminator parity tests, then timing against the default path on the tall-activation layers;
# (4) the opt-in bench steps (perceptual branch; balanced two-lesson step);
# (5) the weight-stationary limit experiment on the 64->64 3x3 layers, and the step with each override.
mkdir -p gpurun_out
exec > >(tee gpurun_out/round2_first.log) 2>&1      # the whole transcript comes back with gpurun_out/
timeout 300 python -m pytest tools/pending_test_trainer_gen_gpu.py tools/pending_test_balance_gpu.py tools/pending_test_enc_gpu.py tools/pending_test_dtw_gpu.py -q -p no:cacheprovider 2>&1 | tail -8
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/halo_probe tools/halo_probe.cu && timeout 60 gpurun_out/halo_probe | tee gpurun_out/halo_probe.txt
BO=""
grep -q "with the base_offset field" gpurun_out/halo_probe.txt && BO="HWG_CONV_HALO_BO=1"
if grep -q "halo view usable: yes" gpurun_out/halo_probe.txt; then
  for mode in 1 2; do
    echo "== numerics, HWG_CONV_HALO=$mode $BO"
    env HWG_CONV_HALO=$mode $BO timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_conv_bwd_gpu.py tests/test_disc_gpu.py tests/test_hwr_train_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
  done
  for mode in 0 1 2; do
    echo "== timing, HWG_CONV_HALO=$mode $BO"
    env HWG_CONV_HALO=$mode $BO timeout 200 python tools/conv_bench.py t_disc_convs1_0 t_disc_convs1_3 t_disc_convs2_0 t_disc_convs3_0 t_disc_convs3_4 t_gen_b2c2
    env HWG_CONV_HALO=$mode $BO timeout 200 python tools/step_runner.py gen_train --B 16 --steps 20 --graph
  done
fi
# (4) the opt-in bench steps: perceptual branch inside the 'gen' step, and the balanced two-lesson optimizer step
echo "== bench, HWG_BENCH_PERCEPTUAL=1"
HWG_BENCH_PERCEPTUAL=1 HWG_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_perceptual.json 2> gpurun_out/bench_perceptual.err; tail -c 600 gpurun_out/bench_perceptual.json; tail -3 gpurun_out/bench_perceptual.err
echo "== bench, HWG_BENCH_BALANCED=1"
HWG_BENCH_BALANCED=1 HWG_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_balanced.json 2> gpurun_out/bench_balanced.err; tail -c 600 gpurun_out/bench_balanced.json; tail -3 gpurun_out/bench_balanced.err
for kb in 40 80; do
  echo "== HWG_CONV_WSTAT_KB=$kb"
  HWG_CONV_WSTAT_KB=$kb HWG_CONV_TILE_W=32 timeout 200 python tools/conv_bench.py t_disc_convs1_0 t_disc_convs1_3 t_disc_convs2_0 t_disc_convs3_0 t_disc_convs3_4 t_gen_b2c2
  HWG_CONV_WSTAT_KB=$kb timeout 200 python tools/step_runner.py gen_train --B 16 --steps 20 --graph
done
