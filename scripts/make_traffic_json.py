"""profiles/<summary>.txt (profiles/ncu_summary.py output of one `ncu --set full` capture of a 512^3 stage) ->
profiles/r02_traffic.json: DRAM bytes (read + write) per launch of the three sweep kernels, the constants bench.py
reports as roofline.traffic.  usage: python scripts/make_traffic_json.py profiles/ncu_full_r02a_summary.txt"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    kernels, cur = [], None
    for line in open(path):
        m = re.match(r"kernel: void (\S+?)<([\d, ]+)>", line)
        if m:
            cur = {"name": m.group(1), "targs": [int(x) for x in m.group(2).split(",")], "read": 0.0, "write": 0.0}
            kernels.append(cur)
            continue
        m = re.match(r"\s+DRAM (read|write)\s+([\d.]+) (\w)byte", line)
        if m and cur is not None:
            cur[m.group(1)] = float(m.group(2)) * {"G": 1e9, "M": 1e6, "K": 1e3}[m.group(3)]
    out = {"file": os.path.relpath(path, ROOT)}
    for k in kernels:
        total = k["read"] + k["write"]
        if k["name"] == "sweep_march" and k["targs"][0] == 0:
            out.setdefault("sweep_x", total)
        elif k["name"] == "sweep_march" and k["targs"][0] == 1:
            out.setdefault("sweep_y", total)
        elif k["name"] == "sweep_rows":
            out.setdefault("sweep_z_epilogue", total)
            out.setdefault("sweep_z_epilogue_instantiation", "sweep_rows<%s>" % ", ".join(map(str, k["targs"])))
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(out)


if __name__ == "__main__":
    main(os.path.abspath(sys.argv[1]))
