#!/bin/bash
# gpurun job: GPU test suite, smoke(), the default bench line and the reference arm.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_gpu.log
tail -5 $O/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2_smoke.log
timeout 600 python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err; echo "bench rc=$?"
cat $O/r2_bench_default.json; tail -5 $O/r2_bench_default.err
