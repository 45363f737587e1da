set -x
N=${N:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/shard_check.py > gpurun_out/r02v_shard_check_g$N.log 2>&1; echo rc=$?; grep -c "True" gpurun_out/r02v_shard_check_g$N.log; grep "False\|Error\|error\|8197.*clahe rows\|8197.*robust" gpurun_out/r02v_shard_check_g$N.log | head -8
