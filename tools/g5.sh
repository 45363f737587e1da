set -x
R=r02j
N=${N:-4}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "jpeg or encoded or streamed or polops" > gpurun_out/${R}_tests.log 2>&1; tail -4 gpurun_out/${R}_tests.log
timeout 200 python bench.py --config c4 --no-cpu-baseline > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; cat gpurun_out/${R}_bench_c4.json | cut -c1-1800; tail -3 gpurun_out/${R}_bench_c4.err
timeout 200 python bench.py --config read > gpurun_out/${R}_bench_read.json 2> gpurun_out/${R}_bench_read.err; cat gpurun_out/${R}_bench_read.json | cut -c1-2200; tail -3 gpurun_out/${R}_bench_read.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/shard_trace.py > gpurun_out/${R}_trace_g$N.log 2>&1; grep -v "^\*\|OMP_NUM\|^W1\|^$" gpurun_out/${R}_trace_g$N.log | tail -60
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N > gpurun_out/${R}_bench_g$N.json 2> gpurun_out/${R}_bench_g$N.err; cat gpurun_out/${R}_bench_g$N.json | cut -c1-2600; tail -3 gpurun_out/${R}_bench_g$N.err
