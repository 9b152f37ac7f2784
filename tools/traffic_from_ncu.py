#!/usr/bin/env python
"""ncu launch list (--csv --print-units base, metrics dram__bytes_read.sum / dram__bytes_write.sum
/ gpu__time_duration.sum) -> DRAM bytes per launch of the dominant solid kernel, per solid element,
stored under `key` in profiles/dram_traffic.json (what bench.py's roofline.traffic is scaled from).
With --full-memvars a "launch" is the pair k_solid_tile + k_anel_full (bench.py times them as one).

    tools/traffic_from_ncu.py CSV KEY NTHETA NR [NOTE]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path, key, ntheta, nr = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    from axisem_b200.host import prem_mesh_spec
    spec = prem_mesh_spec(ntheta=ntheta, nr_target=nr)
    nel_s = int(ntheta * (~spec.fluid_ir).sum())
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head = rows[0]
    ik, im, iv, ii = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("ID")
    per = {}
    for r in rows[1:]:
        per.setdefault((r[ii], r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
    kinds = {"k_solid_tile": [], "k_anel_full": []}
    for (_, name), m in per.items():
        for k in kinds:
            if k in name:
                kinds[k].append(m)
    want = ["k_solid_tile"] + (["k_anel_full"] if key.endswith("_full") else [])
    rd = wr = ns = 0.0
    for k in want:
        ms = kinds[k]
        assert ms, f"no {k} launch in {path}"
        rd += sum(m["dram__bytes_read.sum"] for m in ms) / len(ms)
        wr += sum(m["dram__bytes_write.sum"] for m in ms) / len(ms)
        ns += sum(m["gpu__time_duration.sum"] for m in ms) / len(ms)
    out = os.path.join(ROOT, "profiles", "dram_traffic.json")
    tr = json.load(open(out))
    tr[key] = {"bytes_per_solid_element": (rd + wr) / nel_s,
               "source": f"{os.path.basename(path)} (ncu dram__bytes_read.sum + dram__bytes_write.sum of {' + '.join(want)}, "
                         f"{nel_s:,} solid elements: {rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written per launch, "
                         f"{ns / 1e6:.3f} ms under ncu){'; ' + note if note else ''}"}
    json.dump(tr, open(out, "w"), indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(tr, open(os.path.join(ROOT, "gpurun_out", "dram_traffic.json"), "w"), indent=1)
    print(key, tr[key])


if __name__ == "__main__":
    main()
