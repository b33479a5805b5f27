"""profiles/r02_traffic.json from the raw page of an `ncu --set full` capture: dram__bytes_read.sum + dram__bytes_write.sum per
launch of each kernel, with the workload (samples per iteration) and the commit it was taken on -- what bench.py reports as
`roofline.traffic` (only when its own workload matches).
    python scripts/ncu_traffic.py <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <bench json of the same build> [more raw.csv ...]"""
import csv
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raws, bench = [a for a in sys.argv[1:] if a.endswith(".csv")], [a for a in sys.argv[1:] if a.endswith(".json")][0]
    line = [l for l in open(bench).read().splitlines() if l.startswith("{")][-1]
    b = json.loads(line)
    kernels = {}
    for raw in raws:
        rr = list(csv.reader(open(raw)))
        h, u = rr[0], rr[1]
        ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        for r in rr[2:]:
            name = re.sub(r"^void ", "", r[ki]).split("(")[0].split("<")[0]      # base name: bench.py looks kernels up without template arguments
            val = float(r[ri].replace(",", "")) * UNIT[u[ri]] + float(r[wi].replace(",", "")) * UNIT[u[wi]]
            kernels.setdefault(name, []).append(val)
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"commit": commit, "workload": b["config"]["workload"], "samples_per_iter": b["config"]["samples_per_iter_per_gpu"],
           "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches)",
           "kernels": {k: sum(v) / len(v) for k, v in kernels.items()}}
    json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
