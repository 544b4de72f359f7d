"""Launch list (ncu --csv, metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) -> one entry of
profiles/r02_traffic.json.   usage: traffic_from_ncu.py launches.csv workload [traffic.json]

Per kernel name: launches, total ms, DRAM bytes; the entry keeps the photon-loop kernels' bytes per CALL (all launches of a
call summed: packet-per-lane kernel + packet-per-warp launches), which is what bench.py's roofline.traffic quotes."""
import csv, json, sys, re, collections


def parse(path):
    rows = collections.OrderedDict()
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    return list(rows.values())


def short(name):
    m = re.match(r"(?:void )?([A-Za-z_0-9:]+)", name)
    return (m.group(1) if m else name).split("::")[-1]


def main():
    path, workload = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else None
    rows = parse(path)
    per = collections.OrderedDict()
    for r in rows:
        s = per.setdefault(short(r["name"]), {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
        s["launches"] += 1
        s["ms"] += r.get("gpu__time_duration.sum", 0.0)
        s["dram_bytes"] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
    total_ms = sum(s["ms"] for s in per.values())
    loop = {k: v for k, v in per.items() if k in ("mc_photon_loop_kernel", "mc_warp_engine_kernel")}
    lane = per.get("mc_photon_loop_kernel", {"launches": 0})
    eng = per.get("mc_warp_engine_kernel", {"launches": 0})
    # one lane launch per large call; engine-only calls have no lane launch
    calls = lane["launches"] if lane["launches"] else max(1, eng["launches"])
    entry = {
        "source": path,
        "calls": calls,
        "kernels": {k: {"launches": v["launches"], "ms_under_ncu": round(v["ms"], 3), "dram_bytes": v["dram_bytes"]} for k, v in per.items()},
        "dram_bytes_per_launch": sum(v["dram_bytes"] for v in loop.values()) / calls,
        "share_of_gpu_time_in_photon_loop_kernels": sum(v["ms"] for v in loop.values()) / total_ms if total_ms else None,
    }
    print(json.dumps(entry, indent=1))
    if out:
        try:
            allw = json.load(open(out))
        except Exception:
            allw = {}
        allw[workload] = entry
        json.dump(allw, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
