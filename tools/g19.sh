set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "production_kernels or golden or pipeline_synrgb or tensor_core or single_resized or multiband or streamed or batch" > gpurun_out/r02u_tests.log 2>&1; tail -5 gpurun_out/r02u_tests.log
timeout 300 python tools/fallback_bench.py > gpurun_out/r02u_fallback.json 2> gpurun_out/r02u_fallback.err; cat gpurun_out/r02u_fallback.json; tail -3 gpurun_out/r02u_fallback.err
