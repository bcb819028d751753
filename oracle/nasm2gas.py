#!/usr/bin/env python
"""TEST INFRASTRUCTURE — build-time translator, never part of the product.

nasm is not in this image, so the reference's x86-64 field arithmetic (rust-rapidsnark/rapidsnark/src/asm/fr.asm,
fq.asm: MULX/ADCX/ADOX Montgomery code, selected by meson.build:105-113 on x86-64) cannot be assembled as is. This
script rewrites the NASM source text, read where it lies, into GNU-as Intel syntax on stdout so that `as` (binutils)
can assemble it: a faithful CPU baseline. Nothing is copied into the repository: the output goes to
oracle/_ref/ (git-ignored) at build time, exactly like the objects compiled from the reference's C++ sources.

Only the handful of NASM constructs those two files use are handled:
  ; comments, global/extern, DEFAULT REL (symbol memory operands become rip-relative), section .text/.data,
  %ifdef PIC / %else / %endif (the PIC branch is kept), dq/dd data with or without a leading label,
  size keywords (qword [..] -> qword ptr [..]), `mov qword reg, imm`, $hex literals, movsx r64, r32 (-> movsxd).
"""
import re
import sys

REGS = set("""rax rbx rcx rdx rsi rdi rbp rsp r8 r9 r10 r11 r12 r13 r14 r15 eax ebx ecx edx esi edi ebp esp
r8d r9d r10d r11d r12d r13d r14d r15d ax bx cx dx al bl cl dl rip""".split())


def mem_fix(m):
    inner = m.group(1).strip()
    first = re.match(r"[A-Za-z_][A-Za-z0-9_]*", inner)
    if first and first.group(0).lower() not in REGS:
        return "[rip + " + inner + "]"
    return "[" + inner + "]"


def convert(text):
    out = [".intel_syntax noprefix"]
    skip_else = False
    in_ifdef = False
    for raw in text.splitlines():
        line = raw.split(";", 1)[0].rstrip()
        s = line.strip()
        if not s:
            continue
        low = s.lower()
        if low.startswith("%ifdef"):
            in_ifdef, skip_else = True, False
            continue
        if low.startswith("%else"):
            skip_else = True
            continue
        if low.startswith("%endif"):
            in_ifdef, skip_else = False, False
            continue
        if in_ifdef and skip_else:
            continue
        if low.startswith("global "):
            out.append(".globl " + s.split()[1])
            continue
        if low.startswith("extern ") or low == "default rel":
            continue
        if low.startswith("section "):
            out.append(".text" if ".text" in low else ".data")
            continue
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)?\s*\b(dq|dd)\b\s+(.*)$", s)
        if m and (m.group(1) is None or m.group(1).lower() not in ("mov", "add")):
            label, kind, vals = m.groups()
            if label:
                out.append(label + ":")
            out.append(("    .quad " if kind == "dq" else "    .long ") + vals)
            continue
        s = re.sub(r"\bWRT \.\.plt\b", "", s, flags=re.I)
        s = re.sub(r"^(call\s+)([A-Za-z_][A-Za-z0-9_]*)\s*$", r"\1\2@PLT", s) if "Fr_fail" in s or "Fq_fail" in s else s
        s = re.sub(r"\$([0-9a-fA-F]+)\b", lambda mm: "0x" + mm.group(1), s)
        s = re.sub(r"\b(mov)\s+qword\s+(r[a-z0-9]+)\s*,", r"\1 \2,", s)
        s = re.sub(r"\b(qword|dword|word|byte)\s*\[", r"\1 ptr [", s)
        s = re.sub(r"\[([^\]]+)\]", mem_fix, s)
        s = re.sub(r"^movsx(\s+r[a-z0-9]+\s*,\s*(?:e[a-z]{2}|r\d+d)\s*)$", r"movsxd\1", s)
        out.append(("" if s.endswith(":") else "    ") + s)
    out.append('.section .note.GNU-stack,"",@progbits')
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    sys.stdout.write(convert(open(sys.argv[1]).read()))
