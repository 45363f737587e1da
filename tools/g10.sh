set -x
R=r02n
N=${N:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
FULL=1 timeout 600 $T --master-port 29601 tools/shard_check.py > gpurun_out/${R}_shard_check_g$N.log 2>&1; echo rc=$?; grep -c "True" gpurun_out/${R}_shard_check_g$N.log; grep "False\|Error\|error" gpurun_out/${R}_shard_check_g$N.log | head -12
timeout 200 $T --master-port 29602 tools/shard_trace.py > gpurun_out/${R}_trace_g$N.log 2>&1; grep "^trace\|^iter 5" gpurun_out/${R}_trace_g$N.log | tail -30
timeout 300 $T --master-port 29603 bench.py --gpus $N > gpurun_out/${R}_bench_g$N.json 2> gpurun_out/${R}_bench_g$N.err; cat gpurun_out/${R}_bench_g$N.json | cut -c1-330; tail -2 gpurun_out/${R}_bench_g$N.err | cut -c1-300
timeout 300 $T --master-port 29604 bench.py --gpus $N --config c4 > gpurun_out/${R}_bench_c4_g$N.json 2> gpurun_out/${R}_bench_c4_g$N.err; cat gpurun_out/${R}_bench_c4_g$N.json | cut -c1-330; tail -2 gpurun_out/${R}_bench_c4_g$N.err | cut -c1-300
timeout 400 $T --master-port 29605 bench.py --gpus $N --config c5 --steps 5 > gpurun_out/${R}_bench_c5_g$N.json 2> gpurun_out/${R}_bench_c5_g$N.err; cat gpurun_out/${R}_bench_c5_g$N.json | cut -c1-330; tail -2 gpurun_out/${R}_bench_c5_g$N.err | cut -c1-300
