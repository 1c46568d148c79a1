#!/usr/bin/env python
"""profiles/r2_traffic.json from the `ncu --set full` captures: DRAM read / write bytes of ONE launch of each sweep kernel at
the bench's launch shape (64 pairs of 1024 x 1024 per launch), which bench.py scales to its pairs-per-launch and
reports as roofline.traffic.  Usage: python tools/make_traffic.py key=path.ncu-rep[:pairs] ..."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def metrics(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    tu = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    i = hdr.index("gpu__time_duration.sum")
    return {"kernel": r[hdr.index("Kernel Name")][:120], "dram_read_bytes": get("dram__bytes_read.sum"),
            "dram_write_bytes": get("dram__bytes_write.sum"),
            "duration_us": float(r[i].replace(",", "")) * tu.get(units[i], 1.0)}


def main():
    out = {}
    for spec in sys.argv[1:]:
        key, _, rest = spec.partition("=")
        path, _, rest2 = rest.partition(":")
        pairs, _, size = rest2.partition(":")
        m = metrics(path)
        m["pairs"] = int(pairs or 64)
        m["frame"] = int(size or 1024)
        m["source"] = "profiles/r2_ncu_%s.txt (ncu --set full --clock-control none, ONE launch of %d pairs of %dx%d)" % (
            key, m["pairs"], m["frame"], m["frame"])
        out[key] = m
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
