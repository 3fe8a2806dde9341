#!/bin/bash
set -u
O=gpurun_out
: > $O/r2_emu.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest_gpu.log
A="--steps 200 --warmup 20 --no-sweep --no-module --skip-cpu --skip-e2e"
run() { python bench.py $A "$@" > $O/r2_emu.json 2>> $O/r2_emu.err; python - "$*" <<PY
import json, sys
try:
    d = json.loads([l for l in open("$O/r2_emu.json").read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "|", round(d["ms_per_step"] * 1e3, 2), "us", {k: (v["avg_us"], v["launches"]) for k, v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
run --graph
run
for m in split inline deferred side; do run --emulate-dp --dp-mode $m; run --emulate-dp --dp-mode $m --no-graph; done
B="python bench.py --steps 20 --warmup 5 --sweep-all --skip-e2e --skip-cpu"
B200VQ_SPLIT_ROWS=0 $B > $O/r2_sweep_fused.json 2> $O/r2_sweep.err
B200VQ_SPLIT_ROWS=1 $B --no-module > $O/r2_sweep_split.json 2>> $O/r2_sweep.err
python - <<PY
import json
r = {}
for f in ["fused", "split"]:
    try:
        d = json.loads([l for l in open("$O/r2_sweep_%s.json" % f).read().splitlines() if l.startswith("{")][-1])
        r[f] = {(p["K"], p["D"]): p for p in d["sweep"]["points"]}
        if d.get("module"): print("module", d["module"])
    except Exception as e:
        print(f, "ERR", e)
for key in sorted(r.get("fused", {})):
    print(key, " ".join("%s fwd %.0f bwd %.0f (%s %.2f) frac %.3f |" % (f, r[f][key]["fwd_us"], r[f][key]["bwd_us"], r[f][key]["backward_path"], r[f][key]["bwd_frac_hbm"], r[f][key]["frac"]) for f in r if key in r[f]))
PY
tail -n 3 $O/r2_emu.err $O/r2_sweep.err
