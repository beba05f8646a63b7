#!/bin/bash
# Every documented A/B switch still yields a correct library: the parity tests of the path it touches, once per switch.
run() {  # run VAR=value "k expression" files...
  local setting=$1 kexpr=$2; shift 2
  echo "== $setting"
  if [ -n "$kexpr" ]; then env "$setting" python -m pytest -q -m gpu -x "$@" -k "$kexpr" 2>&1 | tail -1
  else env "$setting" python -m pytest -q -m gpu -x "$@" 2>&1 | tail -1; fi
}
run PSGLA_CONV_PAIR=0 "conv_layer or dncnn_forward or psgla_replay or fused_next" tests/test_image_gpu.py
run PSGLA_CONV_SS=1 "conv_layer or dncnn_forward or psgla_replay" tests/test_image_gpu.py
run PSGLA_CONV_FUSE2=0 "dncnn_forward or psgla_replay or reference_fixture or fused_layer" tests/test_image_gpu.py
run PSGLA_CONV_ISSUE=0 "conv_layer or dncnn_forward or psgla_replay" tests/test_image_gpu.py
run PSGLA_CHAIN=1 "dncnn_forward or psgla_replay or reference_fixture" tests/test_image_gpu.py
run PSGLA_FUSE_PRE=0 "psgla_replay or batched or statistics_only or image_set" tests/test_image_gpu.py
run PSGLA_CONV_ALTERNATE=0 "dncnn_forward or psgla_replay" tests/test_image_gpu.py
run PSGLA_BLUR_4PASS=1 "deblur or pnpula" tests/test_image_gpu.py tests/test_torch_stream_gpu.py
run PSGLA_CG_PAIR=0 "" tests/test_drunet_gpu.py
run PSGLA_CG_REUSE=0 "" tests/test_drunet_gpu.py
run PSGLA_CG_MODE=1 "" tests/test_drunet_gpu.py
run PSGLA_GMM_GEOM=4,1,128,0,0 "" tests/test_gmm2d_gpu.py tests/test_gmm2d_metric_gpu.py
run PSGLA_GMM_GEOM=1,4,32,0,1 "" tests/test_gmm2d_gpu.py
run PSGLA_GMM_STRUCT=0 "" tests/test_gmm2d_gpu.py
