#!/bin/bash
# Measurement hygiene of a round on one B200 (from the repository root; outputs in gpurun_out/):
# for every source order / attenuation variant: DRAM traffic of the dominant kernel (ncu, two
# metrics, one pass) -> profiles/dram_traffic.json, then the bench line that quotes it; the ncu
# launch list of the default bench; ncu --set full of the kernels not covered by r02c.
#   tools/hygiene.sh TAG [NTHETA]        (NTHETA: 1792 = half the bench mesh, still >> L2)
set -uo pipefail
TAG="${1:?tag}"; NT="${2:-1792}"; NR=1116
mkdir -p gpurun_out
cfg() { case "$1" in
    monopole_anel) echo "--source explosion" ;;
    quadpole_anel) echo "--source mtp" ;;
    dipole_elastic) echo "--no-anel" ;;
    dipole_anel_full) echo "--full-memvars" ;;
    dipole_anel) echo "" ;;
esac; }
for key in dipole_anel monopole_anel quadpole_anel dipole_elastic dipole_anel_full; do
    flags="$(cfg $key) --ntheta $NT --nr $NR --no-check --no-cpu-baseline"
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --print-units base \
        --clock-control none -k regex:"k_solid_tile|k_anel_full" -s 3 -c 2 --csv --log-file gpurun_out/${TAG}_${key}_dram.csv \
        python bench.py $flags --steps 2 --warmup 3 --repeats 1 --no-e2e > gpurun_out/${TAG}_${key}_ncu.log 2>&1
    python tools/traffic_from_ncu.py gpurun_out/${TAG}_${key}_dram.csv $key $NT $NR "lean Newmark, product build" \
        || echo "traffic $key failed"
    python bench.py $flags --steps 100 --warmup 10 --repeats 3 > gpurun_out/${TAG}_bench_${key}.json 2> gpurun_out/${TAG}_bench_${key}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${key}.json"))
r = d["roofline"]
print("$key", "ms/step %.4f" % d["ms_per_step"], "value %.2f G" % (d["value"] / 1e9), "S_A frac %.3f" % r["frac"],
      "step frac %.3f" % r["step_frac_of_peak"], "traffic/alg %.3f" % (r["traffic"] / r["algorithmic_bytes_per_launch"]) if r.get("traffic") else "")
PY
done
# launch list of the default bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e --no-check > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.csv
# full captures: k_anel_full on the half mesh; the halo pack / wait kernels on the 2-slice loop-back test
ncu --set full --clock-control none --import-source on -k regex:"k_anel_full" -s 3 -c 1 -o gpurun_out/${TAG}_anel_full -f \
    python bench.py --full-memvars --ntheta $NT --nr $NR --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e --no-check \
    > gpurun_out/${TAG}_anel_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_halo_pack" -s 10 -c 2 -o gpurun_out/${TAG}_halo_pack -f \
    python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_slices" -x > gpurun_out/${TAG}_halo_pack.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep
