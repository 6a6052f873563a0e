"""Prints the few numbers of bench.py JSON lines: python scripts/bench_lines.py gpurun_out/s1_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        if d.get("impl") == "reference":                    # the CPU arm: no kernel, no roofline
            print("%-44s reference arm: value %.4g %s  cores %s" % (f.split("/")[-1], d["value"], d["unit"], d.get("cpu_baseline", {}).get("cores")))
            continue
        r = d["roofline"]
        print("%-44s value %.4g  kernel_ms %.4f  frac %.3f  e2e %.4g  wall/step %.4f" % (f.split("/")[-1], d["value"], r["kernel_ms"], r["frac"], d["e2e"]["value"], d.get("wall_ms_per_step", 0)))
    except Exception as e:
        print(f, "unreadable:", e)
