"""profiles/traffic.json from an `ncu --set full` capture of the K3 kernels (tools/prof_ba.py): DRAM bytes read + written per launch,
stamped with a hash of the bundle-adjustment sources in csrc/ (the files the K3/K4 kernels are compiled from) so that bench.py refuses a
figure that does not belong to the sources it was built from.
    python tools/update_traffic.py gpurun_out/r02_ba.ncu-rep cfg3"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BA_SOURCES = ("ba.cu", "ba_math.cuh", "ba_row.cuh", "chol.cuh")   # the files the captured kernels are written in (common.cuh: context plumbing only)


def source_id():
    h = hashlib.sha1()
    d = os.path.join(ROOT, "sfm-toy-library_b200", "csrc")
    for f in BA_SOURCES:
        h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()
        per[name] = {"dram_bytes": float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]], "us_under_ncu": float(r[it])}
    k3 = sum(v["dram_bytes"] for k, v in per.items() if any(s in k for s in ("ba_point_kernel", "ba_pair_kernel", "ba_camera_kernel", "ba_combine_kernel")))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    cur[workload] = k3
    cur[workload + "_per_kernel"] = per
    cur["source_id"] = source_id()
    cur["capture"] = os.path.basename(rep)
    json.dump(cur, open(path, "w"), indent=1)
    print(json.dumps({workload: k3, "source_id": cur["source_id"]}))


if __name__ == "__main__":
    main()
