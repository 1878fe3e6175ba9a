"""profiles/ncu_traffic.json from the `ncu --set full` captures of a visit: DRAM bytes moved by the observation kernels of
one lock-step step, per (config, batch size) — what bench.py reports as roofline.traffic.
usage: python tools/gpu/ncu_traffic.py TAG Test_03:1024 Test_02:8192 ...   (reads gpurun_out/ncu_<cfg>_<TAG>_raw.csv)
A split launch (k_observe<..., 1> index kernel + k_observe<..., 2> tree kernel) contributes one launch of each."""
import csv
import re
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MODE = {"0": "fused", "1": "index", "2": "trees"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    tag = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    for spec in sys.argv[2:]:
        cfg, n_envs = spec.split(":")
        src = "gpurun_out/ncu_%s_%s_raw.csv" % (cfg, tag)
        try:
            rows = list(csv.reader(open(os.path.join(ROOT, src))))
        except OSError:
            print("missing", src)
            continue
        rows = [r for r in rows if len(r) > 10]
        if len(rows) < 3:
            print("empty", src)
            continue
        hdr, units = rows[0], rows[1]
        col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
        per_kernel = {}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            if "k_observe" not in name:
                continue
            m = re.search(r"k_observe<([^>]*)>", name)
            targs = m.group(1).replace(" ", "").split(",") if m else []
            mode = MODE.get(targs[2] if len(targs) > 2 else "0", "fused")
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            per_kernel[mode] = {"read": rd, "write": wr, "ncu_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")),
                                "kernel": "k_observe<%s>" % ", ".join(targs)}   # last launch of a kind wins
        if not per_kernel:
            print("no k_observe launch in", src)
            continue
        total = sum(v["read"] + v["write"] for v in per_kernel.values())
        table["%s:%s" % (cfg, n_envs)] = {"dram_bytes": total, "kernels": per_kernel,
                                           "source": "profiles/%s_ncu_full_%s.txt (ncu --set full, one launch per kernel)" % (tag, cfg)}
        print(cfg, n_envs, "%.1f MB" % (total / 1e6), {k: round((v["read"] + v["write"]) / 1e6, 1) for k, v in per_kernel.items()})
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
