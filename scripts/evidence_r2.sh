#!/bin/bash
# Round-2 evidence on a B200 box (run through gpurun): bench line, ncu launch list of the same bench command (shorter step
# counts), full captures of the 2D chain kernel and of the fused / stencil image kernels.  Outputs under gpurun_out/.
TAG=${1:-r2}
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-strong --skip-set --skip-gpu-reference > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gmm2d_lean --launch-skip 2 -c 2 -o gpurun_out/prof_gmm_$TAG python bench.py --steps 2 --warmup 3 --skip-image --skip-cpu --skip-strong > gpurun_out/ncu_gmm_$TAG.log 2>&1; echo "gmm rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:deblur_ata|conv3x3_ts_kernel|conv3x3_kernel|pre_inpaint" --launch-skip 30 -c 8 -o gpurun_out/prof_img_$TAG python bench.py --only-image --skip-drunet --skip-gpu-reference > gpurun_out/ncu_img_$TAG.log 2>&1; echo "img rc=$?"
