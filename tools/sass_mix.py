#!/usr/bin/env python
"""SASS instruction mix of one kernel of libcolosseum_b200.so by issue pipe (dev helper, static counts).

    python tools/sass_mix.py ttt_rollout_kernelILi4 [--dump]

ALU pipe (2 warp-inst / clk / SM on B200, tools/int_peak_probe.cu): LOP3 SHF ISETP SEL IADD3 VIADD LEA PRMT VIMNMX PLOP3 MOV ...
FMA pipe: the IMAD family (IMAD, IMAD.SHL, IMAD.MOV, IMAD.IADD, IMAD.HI, IMAD.WIDE);  XU: POPC, FLO, BREV ...
"""
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "colosseumrl_b200", "libcolosseum_b200.so")
ALU = {"LOP3", "SHF", "ISETP", "SEL", "IADD3", "VIADD", "LEA", "PRMT", "VIMNMX", "PLOP3", "MOV", "IABS", "VABSDIFF", "SGXT", "BMSK", "IADD", "VIADDMNMX", "FSEL", "LOP", "P2R", "R2P"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD"}
XU = {"POPC", "FLO", "BREV", "MUFU", "I2F", "F2I"}


def main():
    pat = sys.argv[1]
    names = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    cur, body = None, {}
    for line in names.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", line)
            ins = re.sub(r"\s*/\*.*", "", ins).strip().rstrip(";").strip()
            body[cur].append(ins)
    for fn, ins in body.items():
        if pat not in fn:
            continue
        cnt = {"alu": 0, "fma": 0, "xu": 0, "other": 0}
        ops = {}
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") else t[0]
            base = op.split(".")[0]
            ops[base] = ops.get(base, 0) + 1
            k = "alu" if base in ALU else "fma" if base in FMA else "xu" if base in XU else "other"
            cnt[k] += 1
        print(fn, len(ins), cnt)
        print("  ", sorted(ops.items(), key=lambda x: -x[1])[:24])
        if "--dump" in sys.argv:
            print("\n".join(ins))


if __name__ == "__main__":
    main()
