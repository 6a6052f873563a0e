"""Condenses an `ncu --set full` report into the few numbers DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--source N]  > profiles/<name>_summary.txt

Per kernel launch: duration, DRAM bytes, instruction counts, occupancy limits, issue utilisation,
stall reasons (> 0.3 warps per issue-active cycle); with --source N also the N hottest CUDA source
lines by executed warp-instructions (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield hdr, units, r


def source(rep, topn):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    hdr, data, f, kern = None, [], "", ""
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path": f = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": kern = r[1]; continue
        if r[0] == "Line No":
            hdr = r; ie = hdr.index("Instructions Executed"); ws = hdr.index("Warp Stall Sampling (All Samples)"); te = hdr.index("Thread Instructions Executed"); continue
        if hdr is None or r[0] == "":
            continue
        try:
            data.append((int(r[ie]), int(r[ws]), int(r[te]), f + ":" + r[0], r[1].strip()[:110], kern))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    smp = sum(d[1] for d in data) or 1
    print(f"\nhottest source lines (of {tot} warp-instructions, {smp} stall samples):")
    for d in sorted(data, key=lambda d: -d[0])[:topn]:
        print(f"  {100 * d[0] / tot:5.1f}% inst  {100 * d[1] / smp:5.1f}% smp  {d[2] / max(d[0], 1):4.1f} thr/inst  {d[3]:28s} {d[4]}")


def main():
    rep = sys.argv[1]
    for hdr, units, r in raw(rep):
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:64s} {r[i]:>16s} {units[i]}")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    if float(r[i]) > 0.3:
                        print(f"  stall {h.split('issue_stalled_')[1].split('_per_')[0]:24s} {float(r[i]):6.2f} warps / issue-active cycle")
                except ValueError:
                    pass
    if "--source" in sys.argv:
        source(rep, int(sys.argv[sys.argv.index("--source") + 1]))


if __name__ == "__main__":
    main()
