"""Measures roofline.traffic on a GPU box and stamps it with the kernel sources' hash:

    python scripts/ncu_traffic.py [--config 1] [bench.py arguments...]      -> gpurun_out/ncu_traffic.json (merge into profiles/)

Runs `ncu --set full` over ONE launch of the workload's drain kernel inside the bench command (never a bench number),
sums dram__bytes_read.sum + dram__bytes_write.sum, and writes the entry bench.py reads (profiles/ncu_traffic.json):
bench.py refuses an entry whose src_hash differs from the sources it runs (roofline.traffic = null, with the reason)."""
import csv
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    extra = sys.argv[1:]
    sys.argv = ["bench.py"] + extra
    args = bench.parse_args()
    rep = os.path.join(ROOT, "gpurun_out", f"traffic_{args.handler}")
    os.makedirs(os.path.dirname(rep), exist_ok=True)
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "-k", "regex:drain3_kernel", "-s", "3", "-c", "1", "-o", rep, "-f",
           sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--e2e-steps", "1", "--steps", "2", "--warmup", "1",
           "--sustain-seconds", "0"] + extra
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    out = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, row = rows[0], rows[2]
    units = rows[1]

    def metric(name):
        i = hdr.index(name)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    dram = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
    dur_us = float(row[hdr.index("gpu__time_duration.sum")].replace(",", ""))
    path = os.path.join(ROOT, "gpurun_out", "ncu_traffic.json")
    table = {}
    for p in (os.path.join(ROOT, "profiles", "ncu_traffic.json"), path):
        if os.path.exists(p):
            table.update(json.load(open(p)))
    table[bench.workload_key(args, args.tasks)] = {
        "dram_bytes": dram, "kernel_us_under_ncu": dur_us, "src_hash": bench.kernel_source_hash(),
        "captured": time.strftime("%Y-%m-%d"), "kernel": row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "drain3_kernel"}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(path, table[bench.workload_key(args, args.tasks)])


if __name__ == "__main__":
    main()
