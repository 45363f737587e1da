set -x
R=r02i
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "read_ or jpeg or encoded or streamed or batch or polops" > gpurun_out/${R}_tests.log 2>&1; tail -15 gpurun_out/${R}_tests.log
timeout 200 python bench.py --config read > gpurun_out/${R}_bench_read.json 2> gpurun_out/${R}_bench_read.err; cat gpurun_out/${R}_bench_read.json; tail -5 gpurun_out/${R}_bench_read.err
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; cat gpurun_out/${R}_bench.json; tail -3 gpurun_out/${R}_bench.err
SARPRO_STREAM_UPLOAD=0 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${R}_bench_nostream.json 2> gpurun_out/${R}_bench_nostream.err; cat gpurun_out/${R}_bench_nostream.json
