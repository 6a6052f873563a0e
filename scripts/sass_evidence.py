"""SASS evidence for profiles/: per kernel of libb9gpu.so the instruction count, the opcode histogram and the Blackwell /
Hopper-era mnemonics that matter (UBLKCP = cp.async.bulk, SYNCS = mbarrier, ELECT, ...), plus the full listing of the
identity drain kernel.   python scripts/sass_evidence.py [round tag]   (no GPU needed: cuobjdump reads the .so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "beta9_b200", "libb9gpu.so")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
MARK = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "UTC", "LDTM", "STTM", "HMMA", "HGMMA", "LDGSTS", "STG.E.128", "LDS.128", "ATOM", "RED", "LDL", "STL")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = re.findall(r"arch = (sm_\w+)", txt)
    out = [f"libb9gpu.so: arch {sorted(set(arch))}", ""]
    for part in txt.split("Function : ")[1:]:
        name = part.split("\n")[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        lines = [ln for ln in part.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln)]
        ops = [re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?", "", ln).split()[0].rstrip(";") for ln in lines]
        hist = collections.Counter(o.split(".")[0] for o in ops)
        marks = collections.Counter()
        for o in ops:
            for m in MARK:
                if o.startswith(m) or (m.count(".") and o.startswith(m)):
                    marks[m] += 1
        out.append(f"== {dem}")
        out.append(f"   {len(ops)} instructions; top opcodes: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
        out.append("   marks: " + (", ".join(f"{k} x{v}" for k, v in sorted(marks.items())) or "-"))
        if "drain3_kernelILi0" in name:
            with open(os.path.join(ROOT, "profiles", f"{TAG}_sass_drain3_identity.txt"), "w") as f:
                f.write(f"// cuobjdump -sass of {dem} ({len(ops)} instructions), libb9gpu.so built from this tree\n")
                f.write("\n".join(re.sub(r"^\s+(/\*[0-9a-f]+\*/)\s+", r"\1 ", re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln)) for ln in lines) + "\n")
    path = os.path.join(ROOT, "profiles", f"{TAG}_sass_evidence.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
