#!/usr/bin/env python
"""DRAM traffic per launch of the VAD kernel (vad_lane_kernel or vad_energy_zcr_kernel) from an `ncu --set full` capture ->
profiles/r2_vad_traffic.json (bench.py reads it for roofline.traffic instead of a hard-coded ratio).

    python tools/ncu_traffic.py gpurun_out/r2_vad.ncu-rep [pairs in the captured launch]

Without the pair count it is inferred from the launch's DRAM reads (230.4 MB of PCM per 2 h pair);
in the round-2 capture (`-s 4 -c 1` on `bench.py --pairs 16`) the captured launch is the VAD of the
bench's host-buffer (e2e) call, 4 pairs.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BYTES_VAD = 2 * 16000 * 7200 + 4 * 100 * 7200


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def main(path, pairs):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name = hdr.index("Kernel Name")
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    dur = hdr.index("gpu__time_duration.sum")
    best = None
    for row in rows[2:]:
        if "vad_lane_kernel" not in row[name] and "vad_energy_zcr_kernel" not in row[name]:
            continue
        traffic = to_bytes(row[rd], units[rd]) + to_bytes(row[wr], units[wr])
        if pairs <= 0:
            pairs = max(1, int(round(to_bytes(row[rd], units[rd]) / 230.4e6)))
        if best is None or traffic > best["dram_bytes_per_launch"]:
            best = {"kernel": row[name].split("(")[0], "pairs": pairs,
                    "dram_bytes_read": to_bytes(row[rd], units[rd]), "dram_bytes_write": to_bytes(row[wr], units[wr]),
                    "dram_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": BYTES_VAD * pairs,
                    "duration_under_ncu": "%s %s" % (row[dur], units[dur]),
                    "source": "profiles/%s" % os.path.basename(path).replace(".ncu-rep", "_summary.txt")}
    if best is None:
        raise SystemExit("no VAD kernel launch in %s" % path)
    best["ratio"] = best["dram_bytes_per_launch"] / best["algorithmic_bytes_per_launch"]
    out = os.path.join(ROOT, "profiles", "r2_vad_traffic.json")
    with open(out, "w") as fh:
        json.dump(best, fh, indent=1)
    print(json.dumps(best))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
