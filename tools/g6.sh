set -x
R=r02k
N=${N:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/shard_check.py > gpurun_out/${R}_shard_check_g$N.log 2>&1; echo rc=$?; grep -c "True" gpurun_out/${R}_shard_check_g$N.log; grep "False\|Error\|error" gpurun_out/${R}_shard_check_g$N.log | head
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 tools/shard_trace.py > gpurun_out/${R}_trace_g$N.log 2>&1; grep "^trace\|^iter" gpurun_out/${R}_trace_g$N.log | tail -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N > gpurun_out/${R}_bench_g$N.json 2> gpurun_out/${R}_bench_g$N.err; cat gpurun_out/${R}_bench_g$N.json | cut -c1-600; tail -3 gpurun_out/${R}_bench_g$N.err
for nt in 16 8; do SARPRO_STRIP_NT=$nt timeout 100 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | cut -c1-330; done
