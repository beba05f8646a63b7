python -m pytest tests/test_drunet_gpu.py tests/test_image_gpu.py -q -m gpu -x 2>&1 | tail -1
python scripts/ab_lib.py scripts/_ab/libD_bias_hidden_only.so scripts/_ab/libE_nores.so 3 -- scripts/iter_probe.py
for L in D_bias_hidden_only E_nores; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv3x3 -s 40 -c 60 --csv --log-file gpurun_out/h3_$L.csv python scripts/ab_lib.py --run scripts/_ab/lib$L.so scripts/iter_probe.py 32 256 256 4 > /dev/null 2>&1
done
