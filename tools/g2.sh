set -x
R=r02g
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/${R}_tests.log 2>&1; tail -6 gpurun_out/${R}_tests.log
timeout 200 python bench.py --config c4 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; cat gpurun_out/${R}_bench_c4.json; tail -5 gpurun_out/${R}_bench_c4.err
ITERS=2 timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum --clock-control none -k regex:'k_f32' --csv --log-file gpurun_out/${R}_c4_launches.csv python tools/prof_c4.py > gpurun_out/${R}_c4_ncu.log 2>&1
cat gpurun_out/${R}_c4_launches.csv | tail -8
timeout 120 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; cat gpurun_out/${R}_bench.json
