#!/usr/bin/env python3
"""Turn a straight-line SASS listing (cuobjdump -sass) into readable SSA for the FP32 dataflow.

Used while writing the bit-exact projection kernel: it shows which multiplies nvcc contracted into
FFMAs in a given build, so the same rounding sequence can be written with explicit
__fmaf_rn/__fmul_rn/__fadd_rn intrinsics (and fmaf() in the CPU oracle).
usage: cuobjdump -sass x.o | tools/sass_ssa.py <function-substring> [start_hex end_hex]
"""
import re, sys

def main():
    want = sys.argv[1]
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
    text = sys.stdin.read().split("Function :")
    body = [b for b in text if want in b.split("\n")[0]][0]
    regs = {}
    n = [0]
    def val(tok):
        tok = tok.strip().replace(".reuse", "")
        neg = tok.startswith("-")
        if neg: tok = tok[1:]
        ab = tok.startswith("|")
        tok = tok.strip("|")
        if tok == "RZ": v = "0"
        elif re.fullmatch(r"U?R\d+", tok): v = regs.get(tok, tok + "?")
        elif tok.startswith("c["): v = tok
        else: v = tok
        if ab: v = f"abs({v})"
        return ("-" + v) if neg else v
    def new(dst, expr, addr):
        n[0] += 1
        name = f"t{n[0]}"
        print(f"{addr:04x}  {name} = {expr}")
        regs[dst] = name
    for line in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if not m: continue
        addr = int(m.group(1), 16)
        if addr < lo or addr > hi: continue
        ins = m.group(2).strip()
        pred = ""
        pm = re.match(r"(@!?U?P\d+)\s+(.*)", ins)
        if pm: pred, ins = pm.group(1) + " ", pm.group(2)
        op, _, rest = ins.partition(" ")
        args = [a.strip() for a in re.split(r",\s*(?![^\[]*\])", rest)] if rest else []
        base = op.split(".")[0]
        if base in ("FMUL", "FADD", "FFMA", "FMNMX", "DADD", "DMUL", "DFMA"):
            f = {"FMUL": "mul", "FADD": "add", "FFMA": "fma", "FMNMX": "mnmx", "DADD": "dadd", "DMUL": "dmul", "DFMA": "dfma"}[base]
            sfx = op[len(base):]
            new(args[0], f"{pred}{f}{sfx}({', '.join(val(a) for a in args[1:])})", addr)
        elif base == "MUFU":
            new(args[0], f"{pred}{op}({val(args[1])})", addr)
        elif base in ("LDG", "LD"):
            am = re.search(r"\[(R\d+)(?:\.64)?(?:\+(0x[0-9a-f]+))?\]", rest)
            b = regs.get(am.group(1), am.group(1)); off = am.group(2) or "0x0"
            new(args[0], f"{pred}ld[{b}+{off}]", addr)
        elif base in ("LDC", "LDCU"):
            cm = re.search(r"c\[0x0\]\[(0x[0-9a-f]+)\]", rest)
            if cm:
                regs[args[0]] = f"c[{cm.group(1)}]"
                if ".64" in op:
                    rm = re.fullmatch(r"(U?R)(\d+)", args[0])
                    regs[f"{rm.group(1)}{int(rm.group(2))+1}"] = f"c[{cm.group(1)}+4]"
        elif base in ("MOV", "UMOV"):
            regs[args[0]] = val(args[1])
        elif base in ("F2F", "F2I", "I2F", "I2FP", "FSETP", "FCHK", "STG", "ST", "BRA", "CALL", "EXIT", "BSSY", "BSYNC", "FSEL", "SEL"):
            print(f"{addr:04x}  {pred}{op} {', '.join(val(a) if re.fullmatch(r'-?\|?U?R\d+(\.reuse)?\|?', a) else a for a in args)}")
            if base in ("F2F", "F2I", "I2F", "I2FP", "FSEL", "SEL"):
                n[0] += 1; regs[args[0]] = f"t{n[0]}"; print(f"      -> t{n[0]}")
        else:
            # integer / address ops: result is opaque
            if args and re.fullmatch(r"U?R\d+", args[0]):
                regs[args[0]] = f"{args[0]}@{addr:04x}"
main()
