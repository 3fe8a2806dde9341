#!/bin/bash
set -u
O=gpurun_out
B="python bench.py --steps 20 --warmup 5"
timeout 200 $B > $O/r2_bench_rir256_final.json 2> $O/r2_final.err; echo "bench rc=$?"
timeout 150 $B --sweep-all --skip-e2e --no-module --skip-cpu > $O/r2_bench_rir256_sweep_all_final.json 2>> $O/r2_final.err; echo "sweep rc=$?"
python - <<PY
import json
def rd(f):
    return json.loads([l for l in open("$O/" + f).read().splitlines() if l.startswith("{")][-1])
try:
    d = rd("r2_bench_rir256_final.json")
    print("final", round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us frac", d["roofline"]["frac"], "step", d["roofline"]["step"], "traffic", d["roofline"]["traffic"])
    print("  e2e", round(d["e2e"]["value"] / 1e6, 1), "lean", round(d["e2e_lean"]["value"] / 1e6, 1), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "module", d["module"], "tf32", d["peaks"]["tf32"])
    for p in d["sweep"]["points"]: print("  corner", p["K"], p["D"], p["fwd_us"], p["bwd_us"], p["frac_fwd"], p["bwd_frac_hbm"], p["frac"], p["backward_path"])
    print("  clocks", d["clocks"])
except Exception as e:
    print("final ERR", e)
try:
    d = rd("r2_bench_rir256_sweep_all_final.json")
    for p in d["sweep"]["points"]: print("  sweep", p["K"], p["D"], p["fwd_us"], p["bwd_us"], p["frac_fwd"], p["bwd_frac_hbm"], p["frac"], p["backward_path"])
    print("  tf32", d["peaks"]["tf32"])
except Exception as e:
    print("sweep ERR", e)
PY
tail -n 5 $O/r2_final.err
