set -x
python bench.py > gpurun_out/r01p_bench.json 2> gpurun_out/r01p_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01p_bench_reference.json 2>> gpurun_out/r01p_bench.err
ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|sarpro' --csv --log-file gpurun_out/r01p_launches.csv python tools/prof_one.py clahe > gpurun_out/r01p_ncu.log 2>&1
ITERS=1 ncu --set full --clock-control none --import-source on -k regex:'k_hmma|k_dn_hist' -o gpurun_out/r01p_full -f python tools/prof_one.py clahe >> gpurun_out/r01p_ncu.log 2>&1
cat gpurun_out/r01p_bench.json
