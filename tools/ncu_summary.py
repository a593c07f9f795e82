"""Turn an `ncu --set full` report into the committed evidence under profiles/:
  * profiles/<tag>_details.csv        the raw metric page of the captured launch
  * profiles/r02_ncu_traffic.json     per kernel: DRAM read + write bytes of ONE launch, duration, issue-slot utilisation —
                                      what bench.py reports as roofline.traffic (a measurement, never a literal)

    python tools/ncu_summary.py gpurun_out/r02b_prof_crop_tile.ncu-rep --key crop_tile_kernel --tag r02b_prof_crop_tile \
        --n-hyp 252 --alg-bytes 154828800 --what "crop producer, 252 hypotheses, refiner mode"
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--key", required=True, help="entry name in r02_ncu_traffic.json (kernel-name prefix)")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--n-hyp", type=int, default=None)
    ap.add_argument("--alg-bytes", type=float, default=None)
    ap.add_argument("--what", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(head, units, vals)}
    with open(os.path.join(ROOT, "profiles", a.tag + "_details.csv"), "w") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit", "value"])
        for h in head:
            w.writerow([h, m[h][0], m[h][1]])

    def num(name):
        u, v = m[name]
        x = float(v.replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1, "second": 1, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1)
        return x * scale

    entry = {"kernel": m["Kernel Name"][1], "what": a.what, "source": f"profiles/{a.tag}_details.csv (ncu --set full --clock-control none)",
             "dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"), "dram_read_bytes": num("dram__bytes_read.sum"),
             "dram_write_bytes": num("dram__bytes_write.sum"), "duration_s": num("gpu__time_duration.sum"),
             "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
             "warp_instructions": num("smsp__inst_executed.sum"), "registers_per_thread": int(float(m["launch__registers_per_thread"][1])),
             "algorithmic_bytes": a.alg_bytes, "n_hyp": a.n_hyp}
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[a.key] = entry
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
