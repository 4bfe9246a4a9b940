"""Static evidence of the built library (no GPU needed): per kernel, registers / spills / shared memory from ptxas -v
(csrc/build/*.log) and the SASS mnemonics that matter for this path (LOP3 = chi and the theta xors, SHF = the 64-bit
rotations as funnel shifts, LDG/STG widths, local-memory traffic).  `python profiles/sass_summary.py > profiles/<tag>_sass_summary.txt`"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "proof_protocol_decoder_b200", "csrc")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return {n: re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("ppd::", "") for n, d in zip(names, out)}


def ptxas():
    rows = {}
    for log in sorted(glob.glob(os.path.join(CSRC, "build", "*.log"))):
        cur = None
        for ln in open(log):
            m = re.search(r"Compiling entry function '(\w+)' for 'sm_100a'", ln)
            if m:
                cur = m.group(1)
                rows[cur] = {"file": os.path.basename(log)[:-4]}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and cur:
                rows[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", ln)
            if m and cur:
                rows[cur].update(regs=int(m.group(1)), smem=int(m.group(3) or 0))
    return rows


def sass():
    so = os.path.join(ROOT, "proof_protocol_decoder_b200", "libppd_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    ops, cur = {}, None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\w+)", ln)
        if m:
            cur = m.group(1)
            ops[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            ops[cur][m.group(1)] += 1
    return ops


def main():
    rows, ops = ptxas(), sass()
    names = demangle(sorted(rows))
    print("%-44s %-14s %5s %6s %6s %7s | %6s %6s %6s %7s %7s %6s %6s" % ("kernel", "file", "regs", "spill", "stack", "smem", "instr", "LOP3", "SHF", "LDG.128", "STG.128", "LDL", "STL"))
    for k in sorted(rows, key=lambda k: (rows[k]["file"], names[k])):
        r, c = rows[k], ops.get(k, collections.Counter())
        tot = sum(c.values())
        pick = lambda p: sum(v for o, v in c.items() if o.startswith(p))
        wide = lambda p: sum(v for o, v in c.items() if o.startswith(p) and ".128" in o)
        print("%-44s %-14s %5d %6d %6d %7d | %6d %6d %6d %7d %7d %6d %6d" % (names[k][:44], r["file"], r.get("regs", 0), r.get("spill_st", 0) + r.get("spill_ld", 0), r.get("stack", 0), r.get("smem", 0), tot, pick("LOP3"), pick("SHF"), wide("LDG"), wide("STG"), pick("LDL"), pick("STL")))


if __name__ == "__main__":
    main()
