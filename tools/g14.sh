set -x
R=r02q
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/${R}_tests.log 2>&1; tail -6 gpurun_out/${R}_tests.log
timeout 100 python tools/trace_c3.py > gpurun_out/${R}_trace_c3.log 2>&1; grep "^trace\|^stage" gpurun_out/${R}_trace_c3.log | tail -40
timeout 200 python bench.py --config read > gpurun_out/${R}_bench_read.json 2> gpurun_out/${R}_bench_read.err; cat gpurun_out/${R}_bench_read.json | grep -o '"roofline.*"peak_source"'
