#!/bin/bash
# A/B of the run-time variants on the bench workload (short runs, one JSON line each);
# AB_FLAGS: extra bench.py flags (e.g. "--full-memvars --ntheta 1792")
mkdir -p gpurun_out
run() { # name, env...
    local name=$1; shift
    env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --repeats 3 --no-cpu-baseline --no-e2e --no-check ${AB_FLAGS:-} \
        > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    print(n, "ms/step", round(d["ms_per_step"],3), "min", round(d["repeats"]["ms_per_step_min"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["clocks"]["reasons"])
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/ab_{n}.err").read()[-800:])
PY
}
for v in "$@"; do
    case $v in
    default) run default AXB_NOP=1 ;;
    nolean) run nolean AXB_LEAN=0 ;;
    noahead) run noahead AXB_CORR_AHEAD=0 ;;
    classic) run classic AXB_LEAN=0 AXB_GRAPH=0 ;;
    nograph) run nograph AXB_GRAPH=0 ;;
    *) run "$v" AXB_LIBRARY="$PWD/axisem_b200/libaxisem_b200_$v.so" ;;
    esac
done
