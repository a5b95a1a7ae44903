#!/usr/bin/env python
"""Per-kernel SASS opcode histogram and ptxas resource summary of libtsr_b200.so (CPU only).

  python tools/sass_report.py > profiles/r02_sass_opcodes.txt

Counts the Blackwell-specific opcodes that prove which hardware path a kernel uses:
  UTCHMMA / UTCIMMA / UTCQMMA  tcgen05.mma kind::f16 / kind::i8 / kind::f8f6f4   (5th-gen tensor cores, TMEM accumulators)
  LDTM / STTM                  tcgen05.ld / st (TMEM <-> registers)
  UTCBAR                       tcgen05.commit (MMA completion -> mbarrier)
  UTCATOMSWS / UTCALLOC*       TMEM allocation
  UBLKCP                       cp.async.bulk (TMA 1-D bulk copy)      UTMALDG: tensor-map TMA loads (not used here)
  SYNCS                        mbarrier operations                     REDG / ATOMG: global reductions / atomics
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "template-speech-recognition_b200", "libtsr_b200.so")
LOG = os.path.join(ROOT, "template-speech-recognition_b200", "build", "ptxas.log")
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "ATOMG", "HMMA", "IMMA",
        "FFMA", "DADD", "DFMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k in ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "UBLKCP", "LDTM", "STTM", "SYNCS", "UTCATOMSWS") and op.startswith(k)):
                    kernels[cur][k] += 1
                    break
    res = {}
    if os.path.exists(LOG):
        cur = None
        for line in open(LOG):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                res.setdefault(cur, {}).update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                res.setdefault(cur, {}).update(regs=int(m.group(1)))
                sm = re.search(r"(\d+) bytes smem", line)
                res[cur]["smem"] = int(sm.group(1)) if sm else 0
    names = demangle(list(kernels))
    print("SASS opcode histogram of libtsr_b200.so (sm_100a), cuobjdump -sass; resources from ptxas -v")
    print("%-78s %6s %5s %5s %5s" % ("kernel", "instr", "regs", "spill", "smem") + "".join(" %7s" % k[:7] for k in KEYS[:12]))
    tot = collections.Counter()
    for k, c in kernels.items():
        r = res.get(k, {})
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("tsr::", "").replace("void ", "")
        print("%-78s %6d %5s %5s %5s" % (short[:78], c["_total"], r.get("regs", "-"), r.get("spill_st", 0) + r.get("spill_ld", 0), r.get("smem", "-"))
              + "".join(" %7d" % c[x] for x in KEYS[:12]))
        tot.update(c)
    print("%-78s %6d %5s %5s %5s" % ("TOTAL", tot["_total"], "", "", "") + "".join(" %7d" % tot[x] for x in KEYS[:12]))
    print("\ntcgen05 kernels (UTC*MMA > 0): " + ", ".join(sorted({re.sub(r"[(<].*", "", names[k]).replace("tsr::", "").replace("void ", "")
                                                                for k, c in kernels.items() if c["UTCHMMA"] + c["UTCIMMA"] + c["UTCQMMA"]})))
    print("legacy mma.sync (HMMA/IMMA) instructions: %d" % (tot["HMMA"] + tot["IMMA"]))


if __name__ == "__main__":
    main()
