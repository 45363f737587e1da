set -x
R=r02h
N=${N:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/shard_check.py > gpurun_out/${R}_shard_check_g$N.log 2>&1; echo rc=$?; grep -c "True" gpurun_out/${R}_shard_check_g$N.log; grep -v "True$\|True (" gpurun_out/${R}_shard_check_g$N.log | tail -15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N > gpurun_out/${R}_bench_g$N.json 2> gpurun_out/${R}_bench_g$N.err; cat gpurun_out/${R}_bench_g$N.json; tail -3 gpurun_out/${R}_bench_g$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --config c4 > gpurun_out/${R}_bench_c4_g$N.json 2> gpurun_out/${R}_bench_c4_g$N.err; cat gpurun_out/${R}_bench_c4_g$N.json; tail -3 gpurun_out/${R}_bench_c4_g$N.err
