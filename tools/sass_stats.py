#!/usr/bin/env python
"""Static SASS statistics for a kernel of liboptmc.so (no GPU needed).

usage: tools/sass_stats.py <demangled-name-regex> [--loop N] [--dump]

Prints, for each matching kernel, the opcode histogram of every loop body
(backward branch target..branch) and the fp64-pipe share.  fp64-pipe opcodes:
DFMA DADD DMUL DSETP DMNMX (+ F2F.F64/I2F.F64 conversions listed separately).
"""
import collections
import os
import re
import subprocess
import sys

SO = os.environ.get("OPTMC_LIB", "optpricing_b200/liboptmc.so")
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")


def functions():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def reg_reads(ins):
    """32-bit vector-register operand reads of one SASS instruction (B200: the
    register file delivers 2 per cycle per sub-partition; measured with
    tools/pipe_probe3.cu).  64-bit sources of D* ops count twice; .reuse hits,
    uniform registers, constants and immediates are free."""
    body = ins.split(None, 1)
    if body[0].startswith("@"):
        body = body[1].split(None, 1)
    op = body[0]
    if len(body) < 2:
        return 0
    ops = [o.strip() for o in body[1].split(",")]
    n_dst = 2 if op.startswith(("ISETP", "DSETP", "FSETP", "LOP3.LUT P")) or " P" in (ops[0] + " ") and op.startswith("LOP3") else 1
    if op.startswith(("ST", "BRA", "BAR", "EXIT", "RED", "ATOM")):
        n_dst = 0
    srcs = ops[n_dst:]
    wide = op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    seen = set()
    for o in srcs:
        if ".reuse" in o:
            continue
        m = re.search(r"(?<![U])\bR(\d+)\b", o)
        if m and "UR" not in o.split("R" + m.group(1))[0][-1:]:
            seen.add(int(m.group(1)))
    return len(seen) * (2 if wide else 1)


def opcode(ins):
    parts = ins.split()
    if parts[0].startswith("@"):
        parts = parts[1:]
    return parts[0]


def main():
    pat = re.compile(sys.argv[1])
    dump = "--dump" in sys.argv
    fns = functions()
    names = list(fns)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for mangled, nice in zip(names, dem):
        if not pat.search(nice):
            continue
        ins = fns[mangled]
        print(f"== {nice}: {len(ins)} instructions")
        loops = []
        for a, i in ins:
            if "BRA" in i:
                m = re.search(r"0x([0-9a-f]+)", i)
                if m and int(m.group(1), 16) < a:
                    loops.append((int(m.group(1), 16), a))
        for lo, hi in loops:
            body = [(a, i) for a, i in ins if lo <= a <= hi]
            h = collections.Counter(opcode(i).split(".")[0] for _, i in body)
            n64 = sum(h[k] for k in FP64)
            cvt = sum(1 for _, i in body if re.match(r"(@\S+\s+)?(I2F|F2F|F2I)\S*F64", i))
            mufu = h["MUFU"]
            reads = sum(reg_reads(i) for _, i in body)
            print(f"  loop {lo:#x}..{hi:#x}: {len(body)} instr, fp64-pipe {n64} "
                  f"({100.0 * n64 / len(body):.0f}%), f64 cvt {cvt}, MUFU {mufu}, "
                  f"LDS {h['LDS']}, CALL {h['CALL']}, reg reads {reads} (= {reads / 2:.0f} cycles)")
            print("    " + ", ".join(f"{k}:{v}" for k, v in h.most_common(14)))
            if dump:
                for a, i in body:
                    print(f"      {a:04x}  {i}")


if __name__ == "__main__":
    main()
