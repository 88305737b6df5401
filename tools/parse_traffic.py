#!/usr/bin/env python
"""profiles/price_traffic.json from the `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum` captures of tools/r02_profile.sh (one entry per slab shape; bench.py reads it for
`roofline.traffic`).  Usage: python tools/parse_traffic.py profiles/r02_price_traffic_*.csv"""
import csv
import json
import os
import re
import sys

out = []
for path in sys.argv[1:]:
    m = re.search(r"traffic_(\d+)_(\d+)\.csv$", path)
    D, S_loc = int(m.group(1)), int(m.group(2))
    vals, kernel = {}, None
    for row in csv.reader(open(path)):
        if len(row) > 14 and row[0].isdigit():
            kernel = row[4]
            vals.setdefault(row[12], []).append(float(row[14].replace(",", "")))
    rd = sum(vals["dram__bytes_read.sum"]) / len(vals["dram__bytes_read.sum"])
    wr = sum(vals["dram__bytes_write.sum"]) / len(vals["dram__bytes_write.sum"])
    ns = sum(vals["gpu__time_duration.sum"]) / len(vals["gpu__time_duration.sum"])
    out.append({"kernel": kernel, "S_loc": S_loc, "D": D, "algorithmic_bytes": 8 * S_loc * D, "dram_bytes_read": rd,
                "dram_bytes_write": wr, "traffic_bytes": rd + wr, "traffic_over_algorithmic": (rd + wr) / (8.0 * S_loc * D),
                "ncu_time_ns": ns, "launches_averaged": len(vals["gpu__time_duration.sum"]),
                "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                          "--clock-control none (tools/r02_profile.sh), profiles/" + os.path.basename(path)})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "price_traffic.json"), "w"),
          indent=1)
for e in out:
    print(e["S_loc"], e["D"], round(e["traffic_over_algorithmic"], 4), e["ncu_time_ns"])
