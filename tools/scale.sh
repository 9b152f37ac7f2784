#!/bin/bash
# Strong-scaling bench lines on one box: tools/scale.sh TAG "N..." [bench flags]   (gpurun --gpus 8)
TAG=$1; shift; NS=$1; shift
mkdir -p gpurun_out
for N in $NS; do
    if [ "$N" = 1 ]; then
        timeout 600 python bench.py --gpus 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
    else
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N \
            bench.py --gpus $N "$@" > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
    fi
    python - "$TAG" "$N" <<'PY'
import json,sys
t,n=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/{t}_n{n}.json").read().strip().splitlines()[-1])
    print(t, "N", n, "ms/step", round(d["ms_per_step"],4), "value %.2f G" % (d["value"]/1e9), "e2e %.2f G" % (d["e2e"]["value"]/1e9 if d.get("e2e") else 0),
          "check", d.get("check") and d["check"]["rel_l2"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print(t, n, "FAILED", e); print(open(f"gpurun_out/{t}_n{n}.err").read()[-1500:])
PY
done
