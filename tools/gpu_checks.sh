#!/bin/bash
# The GPU-side checks of a round, as run on a B200 box from the repository root (all outputs under
# gpurun_out/, which is scratch; copy what is to be kept into profiles/).
#   tools/gpu_checks.sh tests      parity tests + smoke            (~1 min of box time)
#   tools/gpu_checks.sh bench      1-GPU bench line                (~1 min)
#   tools/gpu_checks.sh launches   ncu launch list of the bench    (~1 min)
#   tools/gpu_checks.sh full       ncu --set full of S_A and S_B   (~1 min, half-size mesh)
#   tools/gpu_checks.sh scale N    torchrun bench on N GPUs        (charged N x)
set -euo pipefail
mkdir -p gpurun_out
case "${1:-tests}" in
tests)
    python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -5
    python -c "import __graft_entry__ as g; g.smoke()" ;;
bench)
    python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    cat gpurun_out/bench_n1.json ;;
launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1
    tail -3 gpurun_out/launches.csv ;;
full)
    ncu --set full --clock-control none --import-source on -k regex:"k_solid_tile|k_solid_corrector" -s 6 -c 2 \
        -o gpurun_out/full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ntheta 1792 \
        > gpurun_out/ncu_full.log 2>&1
    python profiles/ncu_summary.py gpurun_out/full.ncu-rep ;;
scale)
    N="${2:?number of GPUs}"
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 2950"$N" \
        bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n"$N".json 2> gpurun_out/bench_n"$N".err
    cat gpurun_out/bench_n"$N".json ;;
*) echo "unknown check $1"; exit 2 ;;
esac
