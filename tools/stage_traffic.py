#!/usr/bin/env python
"""Measured DRAM traffic per pipeline stage and image from an `ncu --set full --page raw --csv` export of one
extraction (tools/profile_run.py --images N): dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's
launches, divided by N. bench.py reads the result for `roofline.traffic`.

    python tools/stage_traffic.py gpurun_out/r1z_full_raw.csv 64 > profiles/r1z_stage_traffic.json
"""
import csv
import json
import re
import sys

STAGE_OF = [("k_level0", "level0"), ("k_contrast", "contrast"), ("k_prep", "prep"), ("k_fed", "fed"), ("k_detector", "detector"),
            ("k_rowcount", "compact"), ("k_rowscan", "compact"), ("k_scatter", "compact"), ("k_dedup", "dedup"),
            ("k_class_ranges", "finalize"), ("k_filter_refine", "finalize"), ("k_keep_scan", "finalize"), ("k_orientation", "finalize"),
            ("k_descriptor", "descriptor")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, n_images):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stages = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).split("::")[-1]
        st = next((s for p, s in STAGE_OF if name.startswith(p)), None)
        if st is None:
            continue
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[idx[m]].replace(",", "")) * SCALE.get(units[idx[m]], 1.0)
        d = stages.setdefault(st, {"dram_bytes_per_image": 0.0, "launches_per_batch": 0})
        d["dram_bytes_per_image"] += b / n_images
        d["launches_per_batch"] += 1
    json.dump({"source": "%s: ncu --set full --clock-control none, tools/profile_run.py --images %d (dram__bytes_read.sum + "
                         "dram__bytes_write.sum summed over the stage's launches, per image)" % (path, n_images),
               "images_in_capture": n_images, "stages": stages}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
