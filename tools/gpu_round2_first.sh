#!/bin/bash
# First GPU call of the next round (one GPU, ~2 min): (1) the parked GPU tests (trainer-level gen lesson, flat gradient balancing, Encoder2 / perceptual loss), (2) the weight-stationary
# limit experiment on the 64->64 3x3 layers and the discriminator conv shapes, (3) the step with the override.
mkdir -p gpurun_out
timeout 200 python -m pytest tools/pending_test_trainer_gen_gpu.py tools/pending_test_balance_gpu.py tools/pending_test_enc_gpu.py -q -p no:cacheprovider 2>&1 | tail -6
for kb in 40 80; do
  echo "== HWG_CONV_WSTAT_KB=$kb"
  HWG_CONV_WSTAT_KB=$kb HWG_CONV_TILE_W=32 timeout 200 python tools/conv_bench.py t_disc_convs1_0 t_disc_convs1_3 t_disc_convs2_0 t_disc_convs3_0 t_disc_convs3_4 t_gen_b2c2
  HWG_CONV_WSTAT_KB=$kb timeout 200 python tools/step_runner.py gen_train --B 16 --steps 20 --graph
done
