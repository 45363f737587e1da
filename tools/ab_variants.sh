#!/bin/bash
# A/B of experimental builds of libsarpro_gpu (sarpro_b200.build --variant=...): parity of the tensor-core pass B against the
# oracle, then the C3 step time and device timeline. Usage: tools/ab_variants.sh <tag> <variant> [<variant> ...] ("default" = the product .so)
tag=$1; shift
for v in "$@"; do
  lib=""; [ "$v" != default ] && lib="$PWD/sarpro_b200/libsarpro_gpu_$v.so"
  echo "=== variant $v" >> gpurun_out/${tag}.log
  SARPRO_GPU_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_pass_b or golden_wide or pipeline_synrgb" >> gpurun_out/${tag}.log 2>&1
  SARPRO_GPU_LIB=$lib TRACE_RANK=0 timeout 300 python tools/batch_trace.py >> gpurun_out/${tag}.log 2>&1
done
tail -n 120 gpurun_out/${tag}.log
