N=${N:-4}
run() { echo "=== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/shard_trace.py 2>&1 | grep "stage 6\|stage 2\|^trace total"; }
run SARPRO_X=1
run SARPRO_NCCL_NO_CONFIG=1
run SARPRO_NCCL_NO_CONFIG=1 SARPRO_SPARE_SMS=16
