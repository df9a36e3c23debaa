#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): default bench line + per-op table, the ncu launch list of the
# same command, and `ncu --set full` captures of the longest conv launch, the stem and the head prediction GEMM.
# usage: bash tools/gpu_round_profile.sh <tag>      (outputs under gpurun_out/, summarised into profiles/ afterwards)
tag=${1:-r1x}
timeout 400 python bench.py --profile-ops > gpurun_out/bench_$tag.json 2> gpurun_out/ops_$tag.txt
tail -c 300 gpurun_out/bench_$tag.json
# -k: this repo's kernels only (bench.py first calibrates BN statistics with torch ops, hundreds of cudnn/ATen launches)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:conv_tc|bneck|stem_tc|spp|dwconv|filter_kernel|sort_nms|conv_simt|focus' -c 400 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_conv3x3_$tag \
  python tools/gpu_prof_conv.py 64 128:256:3:1:80 > gpurun_out/ncu_c_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 1 -c 1 -o gpurun_out/prof_stem_$tag \
  python tools/gpu_prof_stem.py > gpurun_out/ncu_s_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o gpurun_out/prof_head_$tag \
  python tools/gpu_prof_head.py 64 80 > gpurun_out/ncu_h_$tag.log 2>&1
ls -la gpurun_out/*_$tag*
# cluster NMS (config 5 dense scene) and SimOTA: live comparison + one full capture each (profiles/r2_nms_cluster.*, r2_simota.md)
timeout 150 python tools/gpu_nms_cluster.py gpurun_out/nms_cluster_$tag.json > gpurun_out/nms_cluster_$tag.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:sort_nms -s 2 -c 1 -f -o gpurun_out/prof_nms_$tag \
  python tools/gpu_nms_one.py > gpurun_out/ncu_n_$tag.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:simota_assign -c 1 -f -o gpurun_out/prof_simota_$tag \
  python -m pytest tests/test_gpu_simota.py -q -m gpu -k "test_simota_assign_vs_reference and s640" > gpurun_out/ncu_sim_$tag.log 2>&1
for k in nms simota; do ncu -i gpurun_out/prof_${k}_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${k}_${tag}_raw.csv 2>/dev/null; done
