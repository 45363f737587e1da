set -x
R=r02o
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "read_ or polops" > gpurun_out/${R}_tests.log 2>&1; tail -3 gpurun_out/${R}_tests.log
timeout 200 python bench.py --config read > gpurun_out/${R}_bench_read.json 2> gpurun_out/${R}_bench_read.err; cat gpurun_out/${R}_bench_read.json | cut -c1-2000 | grep -o '"roofline.*'; tail -3 gpurun_out/${R}_bench_read.err
ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_f32_quantize|k_f32_hist4096|k_f32_scan' -c 3 -o gpurun_out/${R}_c4_full -f python tools/prof_c4.py > gpurun_out/${R}_c4_ncu.log 2>&1; tail -3 gpurun_out/${R}_c4_ncu.log
