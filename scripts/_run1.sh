python -m pytest tests -x -q -m gpu > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3d_pytest.log
python scripts/drunet_layers.py > gpurun_out/r3d_iter.txt 2>&1
DIAG_N=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3d_drunet_launches.csv python scripts/drunet_layers.py > gpurun_out/r3d_ncu.log 2>&1; echo "rc=$?"
python scripts/drunet_layers.py >> gpurun_out/r3d_iter.txt 2>&1
cat gpurun_out/r3d_iter.txt
