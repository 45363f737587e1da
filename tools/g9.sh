set -x
R=${R:-r02m}
N=${N:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 tools/shard_check.py > gpurun_out/${R}_shard_check_g$N.log 2>&1; echo rc=$?; grep -c "True" gpurun_out/${R}_shard_check_g$N.log; grep "False\|Error\|error\|wide" gpurun_out/${R}_shard_check_g$N.log | head -12
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29592 tools/shard_trace.py > gpurun_out/${R}_trace_g$N.log 2>&1; grep "^trace\|^iter 5" gpurun_out/${R}_trace_g$N.log | tail -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus $N > gpurun_out/${R}_bench_g$N.json 2> gpurun_out/${R}_bench_g$N.err; cat gpurun_out/${R}_bench_g$N.json | cut -c1-330; tail -2 gpurun_out/${R}_bench_g$N.err | cut -c1-300
