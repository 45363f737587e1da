# Round-end measurement set (one gpurun call): bench, reference arm, ncu launch list, one full capture of pass B.
# Every step is bounded so that the call ends inside the GPU budget that is left.
R=${R:-r01q}
set -x
timeout 70 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
if [ -z "$QUICK" ]; then
timeout 40 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
fi
ITERS=2 timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|sarpro' --csv --log-file gpurun_out/${R}_launches.csv python tools/prof_one.py clahe > gpurun_out/${R}_ncu.log 2>&1
if [ -z "$QUICK" ]; then
ITERS=1 timeout 55 ncu --set full --clock-control none --import-source on -k regex:'k_hmma' -c 2 -o gpurun_out/${R}_full -f python tools/prof_one.py clahe >> gpurun_out/${R}_ncu.log 2>&1
fi
cat gpurun_out/${R}_bench.json
