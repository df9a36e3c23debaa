#!/bin/bash
# Training-step evidence on one B200 (run through gpurun): the --train bench line, the ncu launch list of the same command
# restricted to this repo's kernels, and `ncu --set full` captures of the wgrad kernel on two layers.
# usage: bash tools/gpu_train_profile.sh <tag>
tag=${1:-r2x}
timeout 600 python bench.py --train --steps 10 --warmup 3 > gpurun_out/train_$tag.json 2> gpurun_out/train_$tag.err
tail -c 400 gpurun_out/train_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:conv_tc|wgrad|bn_|pack_train|dilate2|spp|simota|head_|sgd_ema' -s 1700 -c 850 --csv \
  --log-file gpurun_out/train_launches_$tag.csv python tools/gpu_train_steps.py 3 > gpurun_out/ncu_train_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 6 -c 1 -o gpurun_out/prof_wgrad128_80_$tag \
  python tools/gpu_prof_wgrad.py 8 128 128 80 3 1 > gpurun_out/ncu_w1_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 6 -c 1 -o gpurun_out/prof_wgrad64_160_$tag \
  python tools/gpu_prof_wgrad.py 64 128 128 80 3 1 > gpurun_out/ncu_w2_$tag.log 2>&1
ls -la gpurun_out/*_$tag*
