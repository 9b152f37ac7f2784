mkdir -p gpurun_out
B="python bench.py --ntheta 1792 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for st in 2 3; do echo "== TE16 stages $st"; AXB_SOLID_STAGES=$st timeout 200 $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; done
for st in 2 3 4 6; do echo "== TE8 stages $st"; AXB_LIBRARY=$PWD/axisem_b200/libaxisem_b200_te8.so AXB_SOLID_STAGES=$st timeout 200 $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solid_tile -s 4 -c 1 -o gpurun_out/prof_solid_tile $B > gpurun_out/ncu_solid.log 2>&1
tail -3 gpurun_out/ncu_solid.log
