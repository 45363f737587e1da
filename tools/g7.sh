set -x
R=${R:-r02l}
N=${N:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tools/shard_trace.py > gpurun_out/${R}_trace_g$N.log 2>&1; grep "^trace\|^iter 5" gpurun_out/${R}_trace_g$N.log | tail -34
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N > gpurun_out/${R}_bench_g$N.json 2> gpurun_out/${R}_bench_g$N.err; cat gpurun_out/${R}_bench_g$N.json | cut -c1-400; tail -2 gpurun_out/${R}_bench_g$N.err | cut -c1-300
