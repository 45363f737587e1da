N=${N:-2}
run() { echo "=== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 tools/shard_trace.py 2>&1 | grep "^trace  4\|^trace  9\|^trace 10\|^trace 12\|^trace 15\|^trace total"; }
run SARPRO_SPARE_SMS=16
run SARPRO_SPARE_SMS=4 NCCL_MAX_NCHANNELS=4
run SARPRO_SPARE_SMS=8 NCCL_MAX_NCHANNELS=8
run SARPRO_SPARE_SMS=16 NCCL_MAX_NCHANNELS=2
