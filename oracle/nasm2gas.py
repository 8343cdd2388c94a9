"""oracle/nasm2gas.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NASM is not in this image, so the reference's 18-bit codec (getiq64.s: compress_rawdat_net / _disk, expand_rawdat)
could not be built for the oracle in round 1.  This script rewrites that file, line by line and without touching
an instruction, into GNU-assembler Intel syntax (`.intel_syntax noprefix`), which gcc assembles.  The output goes to
oracle/_ref/ (not committed: the reference's source stays where it is); oracle/Makefile runs this when
/root/reference is present.

What differs between the two dialects for the instructions this file uses:
  ; comment                -> dropped
  section / extern / global -> .section / .extern / .globl
  label without a colon    -> label:
  mov reg, symbol          -> mov reg, OFFSET symbol      (NASM: a bare symbol is its address)
  11000000B                -> 0b11000000
usage: python nasm2gas.py /root/reference/getiq64.s out.s"""
import re
import sys


def translate(lines):
    symbols = set()
    for ln in lines:
        m = re.match(r"\s*(extern|global)\s+(\w+)", ln)
        if m:
            symbols.add(m.group(2))
    out = [".intel_syntax noprefix"]
    for ln in lines:
        code = ln.split(";", 1)[0].rstrip()
        if not code.strip():
            continue
        t = code.strip()
        m = re.match(r"section\s+(\S+)$", t)
        if m:
            out.append('.section .note.GNU-stack,"",@progbits' if m.group(1) == ".note.GNU-stack" else m.group(1))
            continue
        m = re.match(r"extern\s+(\w+)$", t)
        if m:
            out.append(f".extern {m.group(1)}")
            continue
        m = re.match(r"global\s+(\w+)$", t)
        if m:
            out.append(f".globl {m.group(1)}")
            continue
        if re.match(r"^\w+:?$", t) and t.rstrip(":") not in ("ret",):
            out.append(t.rstrip(":") + ":")
            continue
        t = re.sub(r"\b([01]+)B\b", lambda g: "0b" + g.group(1), t)
        m = re.match(r"mov\s+(\w+)\s*,\s*(\w+)$", t)
        if m and m.group(2) in symbols:
            t = f"mov {m.group(1)}, OFFSET {m.group(2)}"
        out.append("  " + t)
    return out


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        res = translate(f.read().splitlines())
    with open(dst, "w") as f:
        f.write("\n".join(res) + "\n")
