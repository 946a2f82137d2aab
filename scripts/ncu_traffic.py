"""profiles/ncu_traffic.json from an ncu_summary.py CSV: DRAM bytes per launch of the four step kernels
(bench.py reports the dominant kernel's as roofline.traffic)."""
import csv
import json
import sys

src, out, M, P = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
kern = {}
for key, pat in (("k1_zeta_step", "k1_zeta_step"), ("k2_fft_forward", "k2_"), ("k3_ysolve", "k3_ysolve"), ("k4_fft_inverse", "k4_")):
    for r in data:
        if pat in r[0]:
            rd = float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
            wr = float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
            kern[key] = {"dram_bytes_per_launch": rd + wr, "us_under_ncu": float(r[ix["gpu__time_duration.sum"]]),
                         "kernel": r[0].split("(")[0]}
            break
json.dump({"source": f"{src} (ncu --set full --clock-control none, cold caches, {M}x{P}, one launch per kernel)",
           "grid": [M, P], "kernels": kern}, open(out, "w"), indent=1)
print(json.dumps(kern, indent=1))
