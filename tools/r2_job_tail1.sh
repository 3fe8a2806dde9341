#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -x -q > $O/r2_pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/r2_pytest_bwd.log
A="--steps 200 --warmup 20 --no-sweep --no-module --skip-cpu --skip-e2e"
run() { python bench.py $A $2 > $O/r2_emu.json 2>> $O/r2_emu.err
  python - "$1 $2" <<PY
import json, sys
try:
    d = json.loads([l for l in open("$O/r2_emu.json").read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "|", round(d["ms_per_step"] * 1e3, 2), "us", {k: (v["avg_us"], v["launches"]) for k, v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
run "" ""
for c in 2 3 4 6 8; do B200VQ_DP_TAIL_CTAS_PER_SM=$c run "ctas/sm=$c" "--emulate-dp"; done
run "" "--emulate-dp --no-graph"
run "" "--emulate-dp --dp-mode inline"
tail -n 5 $O/r2_emu.err
