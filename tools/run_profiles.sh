# Round measurement set (one gpurun call, one GPU): every BASELINE config of bench.py, the reference arm, ncu launch list of the
# C3 step, one full capture of pass B and of the downsample-on-read kernel, sanitizer logs. Every step is bounded.
R=${R:-r02}
set -x
timeout 150 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout 120 python bench.py --config c2 > gpurun_out/${R}_bench_c2.json 2>> gpurun_out/${R}_bench.err
timeout 60 python bench.py --config c1 > gpurun_out/${R}_bench_c1.json 2>> gpurun_out/${R}_bench.err
timeout 200 python bench.py --config c4 > gpurun_out/${R}_bench_c4.json 2>> gpurun_out/${R}_bench.err
timeout 300 python bench.py --config c5 --steps 5 > gpurun_out/${R}_bench_c5.json 2>> gpurun_out/${R}_bench.err
timeout 200 python bench.py --config read > gpurun_out/${R}_bench_read.json 2>> gpurun_out/${R}_bench.err
timeout 90 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
ITERS=3 timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|sarpro' -s 26 --csv --log-file gpurun_out/${R}_launches.csv python tools/prof_plan.py > gpurun_out/${R}_ncu.log 2>&1
if [ -z "$QUICK" ]; then
ITERS=2 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_hmma' -s 2 -c 2 -o gpurun_out/${R}_full -f python tools/prof_plan.py >> gpurun_out/${R}_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_read_average' -c 1 -o gpurun_out/${R}_read_full -f python tools/prof_read.py >> gpurun_out/${R}_ncu.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/${R}_sanitizer_memcheck.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/${R}_sanitizer_racecheck.log 2>&1
fi
cat gpurun_out/${R}_bench.json
tail -3 gpurun_out/${R}_sanitizer_*.log
