#!/bin/bash
# Round-end evidence on a B200 box (run through gpurun): GPU tests, smoke, the bench line, the ncu launch list of the same
# bench command (shorter step counts) and one full capture of the dominant image kernel.  Outputs under gpurun_out/.
TAG=${1:-r1}
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_ts2 --launch-skip 60 -c 2 -o gpurun_out/prof_ts2_$TAG python bench.py --only-image --skip-drunet > gpurun_out/ncu_ts2_$TAG.log 2>&1; echo "ts2 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
for k in ("image", "image_deblur", "image_drunet"):
    i = d[k]
    print(k, i["value"], i.get("roofline", {}).get("achieved"), i.get("whole_iteration_tensor_tflops"), i.get("e2e", {}).get("value"))
print(d["image"]["single_chain_iterations_per_sec"], d["clocks"])
PY
