#!/bin/bash
# Round-2 evidence, second batch: full ncu captures of the A^T A deblur kernel and of the hidden-layer pair kernel at the bench's
# image workload, and the launch list of the final bench command.  Outputs under gpurun_out/.
TAG=${1:-r2p}
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:deblur_ata" --launch-skip 4 -c 2 -o gpurun_out/prof_ata_$TAG python bench.py --only-image --skip-drunet --skip-gpu-reference > gpurun_out/ncu_ata_$TAG.log 2>&1; echo "ata rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:conv3x3_ts2" --launch-skip 60 -c 2 -o gpurun_out/prof_ts2_$TAG python bench.py --only-image --skip-drunet --skip-gpu-reference > gpurun_out/ncu_ts2_$TAG.log 2>&1; echo "ts2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-strong --skip-set --skip-gpu-reference > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
