"""Opcode histogram (executed warp instructions) of one kernel from an ncu report's source page.
  python tests/tools/ncu_opcodes.py gpurun_out/<tag>_prof.ncu-rep <kernel regex> [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    ops, tot, launches = collections.Counter(), 0, 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            launches += 1
            if launches > 1:
                break
            continue
        if r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) <= hdr["Instructions Executed"]:
            continue
        src = r[hdr["Source"]].strip().split()
        if not src:
            continue
        op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
        op = op.rstrip(";")
        key = op.split(".")[0]
        if key in ("LDS", "STS", "LDG", "STG", "ATOMS", "RED", "MUFU", "ATOMG"):
            key = ".".join(op.split(".")[:2])
        n = int(r[hdr["Instructions Executed"]])
        ops[key] += n
        tot += n
    print(f"{kern}: {tot / 1e6:.1f} M warp instructions (first captured launch)")
    for k, v in ops.most_common(top):
        print(f"  {k:14s} {v / 1e6:8.1f} M {v / tot:6.1%}")


if __name__ == "__main__":
    main()
