#!/usr/bin/env python
"""Per-pipe SASS instruction histogram of a kernel's hot loop (cuobjdump -sass on the built
libuwtrack.so).  The hot loop is the innermost loop (smallest backward-branch span) that
contains the marker instruction (default: the byte gather LDG.E.U8 of the residual sweep).

    python tools/sass_mix.py --kernel estimate_flow_kernelILb0 --out profiles/r02_k4_sass_mix.md
    python tools/sass_mix.py --kernel gradient_kernelILb0ELb0 --marker IDP.4A --whole
"""
import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PIPES = [
    ("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX)"),
    ("xu (conversion / MUFU)", r"^(MUFU|F2F|F2I|I2F\.F64|I2F\b|F2FP|FRND)"),
    ("fma / fp32", r"^(FFMA|FMUL|FADD|FSETP|FSEL|FMNMX|FCHK|HFMA2|HADD2|HMUL2|I2FP|FSWZADD)"),
    ("int / alu", r"^(IMAD|IADD|VIADD|LOP3|SHF|LEA|ISETP|SEL|IABS|IMNMX|VIADDMNMX|VIMNMX|PLOP3|"
                  r"POPC|FLO|BREV|PRMT|SGXT|IDP|MOV|CS2R|S2R|BMSK|VABSDIFF|P2R|R2P|LOP\b)"),
    ("uniform datapath", r"^(U[A-Z0-9]+|R2UR|S2UR|LDCU|REDUX|VOTEU)"),
    ("lsu shared", r"^(LDS|STS|LDSM|ATOMS)"),
    ("lsu global / const", r"^(LDG|STG|LD\b|ST\b|LDC|ATOMG|ATOM|RED|LDL|STL|CCTL|MEMBAR|FENCE|"
                           r"ERRBAR|UBLKCP|UTMALDG|UTMASTG|UTMAPF|LDGSTS|LDGDEPBAR|DEPBAR|SYNCS)"),
    ("branch / control", r"^(BRA|BSSY|BSYNC|BREAK|CALL|RET|EXIT|WARPSYNC|BAR|NOP|YIELD|NANOSLEEP|"
                         r"BMOV|JMP|BRX|JMX|ELECT|ENDCOLLECTIVE|ACQBULK)"),
    ("warp shuffle / vote", r"^(SHFL|VOTE|MATCH|REDUX)"),
]


def sass_of(so, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        if kernel in name:
            ins = []
            for l in f.split("\n"):
                m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip()))
            return name, ins
    raise SystemExit("kernel %s not found in %s" % (kernel, so))


def hot_loop(ins, marker):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            s = addr[tgt]
            if any(marker in x for _, x in ins[s:i + 1]):
                if best is None or (i - s) < (best[1] - best[0]):
                    best = (s, i)
    return best


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--so", default=os.path.join(ROOT, "uw_slam_b200", "libuwtrack.so"))
    ap.add_argument("--kernel", required=True, help="substring of the mangled kernel name")
    ap.add_argument("--marker", default="LDG.E.U8")
    ap.add_argument("--whole", action="store_true", help="histogram of the whole kernel")
    ap.add_argument("--out")
    ap.add_argument("--json")
    ap.add_argument("--listing", action="store_true", help="append the SASS of the loop")
    args = ap.parse_args()
    name, ins = sass_of(args.so, args.kernel)
    if args.whole:
        s, e = 0, len(ins) - 1
    else:
        hl = hot_loop(ins, args.marker)
        if hl is None:
            raise SystemExit("no loop containing %s" % args.marker)
        s, e = hl
    body = ins[s:e + 1]
    ops = collections.Counter(opcode(t) for _, t in body)
    pipes = collections.Counter()
    for op, n in ops.items():
        for pname, rx in PIPES:
            if re.match(rx, op):
                pipes[pname] += n
                break
        else:
            pipes["other"] += n
    out = ["# SASS instruction mix: `%s`" % name, "",
           "%s: %d instructions (addresses 0x%x .. 0x%x), `cuobjdump -sass` of the built library."
           % ("whole kernel" if args.whole else "hot loop (innermost loop containing `%s`)"
              % args.marker, len(body), body[0][0], body[-1][0]), "",
           "| pipe | instructions | share |", "|---|---|---|"]
    for pname, n in sorted(pipes.items(), key=lambda x: -x[1]):
        out.append("| %s | %d | %.1f %% |" % (pname, n, 100.0 * n / len(body)))
    out += ["", "| opcode | count |", "|---|---|"]
    for op, n in sorted(ops.items(), key=lambda x: (-x[1], x[0])):
        out.append("| `%s` | %d |" % (op, n))
    if args.listing:
        out += ["", "```"] + ["%05x  %s" % (a, t) for a, t in body] + ["```"]
    text = "\n".join(out) + "\n"
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"kernel": name, "instructions": len(body), "pipes": dict(pipes),
                       "opcodes": dict(ops)}, f, indent=1)


if __name__ == "__main__":
    main()
