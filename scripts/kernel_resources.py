"""Registers, stack (spill / local arrays), static shared memory and spill-instruction counts of every kernel in the
shipped libkzp_b200.so, read from the binary itself (cuobjdump --dump-resource-usage, cuobjdump -sass): no GPU needed.

    python scripts/kernel_resources.py > profiles/r02_kernel_resources.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "keyless-zk-proofs_b200", "libkzp_b200.so")
STAMP = os.path.join(ROOT, "keyless-zk-proofs_b200", "build", "stamp.sha256")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = name.replace("kzp::", "")
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)            # drop the argument list
    name = name.replace("XyzzT<Fp<FqParams> >", "G1").replace("XyzzT<Fp2T<Fp<FqParams> > >", "G2")
    name = name.replace("Fp2T<Fp<FqParams> >", "Fq2").replace("Fp<FqParams>", "Fq").replace("Fp<FrParams>", "Fr")
    return name


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    rows = {}
    fn = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows[fn] = [int(x) for x in m.groups()]
            fn = None
    # spill traffic: STL / LDL instruction counts per kernel in the SASS
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    spills, total, wide, cur = {}, {}, {}, None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            spills[cur] = [0, 0]
            total[cur] = 0
            wide[cur] = 0
            continue
        if cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            total[cur] += 1
            if "IMAD.WIDE" in line:
                wide[cur] += 1
            if re.search(r"\bSTL(\.|\b)", line):
                spills[cur][0] += 1
            elif re.search(r"\bLDL(\.|\b)", line):
                spills[cur][1] += 1
    names = demangle(sorted(rows))
    import hashlib

    body = "\n".join(l for l in sass.split("\n") if l and not l.startswith("Fatbin") and "identifier" not in l)
    sass_md5 = hashlib.md5((body + "\n").encode()).hexdigest()
    print("# kernels of keyless-zk-proofs_b200/libkzp_b200.so (sm_100a), library stamp %s" % open(STAMP).read().strip())
    print("# md5 of the SASS of all kernels (cuobjdump -sass without the Fatbin / identifier lines): %s" % sass_md5)
    print("#   -- equal digests = identical device code; host-only changes move the stamp but not this")
    print("# regs = registers per thread; stack = bytes of per-thread stack (spills + local arrays); STL/LDL = local-memory")
    print("# store / load instructions in the kernel's SASS (static counts); smem = static shared memory per CTA (dynamic")
    print("# shared memory of the NTT / sort kernels is set at launch); instr = SASS instructions, wide = IMAD.WIDE among them")
    print("# (static mix: a loop body counts once however often it runs, and the out-of-line Fq2 routines of the G2 kernels")
    print("#  count once however often they are called)")
    print("%-78s %5s %6s %6s %5s %5s %7s %6s %5s" % ("kernel", "regs", "stack", "smem", "STL", "LDL", "instr", "wide", "share"))
    for mangled in sorted(rows, key=lambda k: short(names[k])):
        reg, stack, shared, local = rows[mangled]
        stl, ldl = spills.get(mangled, [0, 0])
        n_i, n_w = total.get(mangled, 0), wide.get(mangled, 0)
        print("%-78s %5d %6d %6d %5d %5d %7d %6d %4.0f%%" % (short(names[mangled])[:78], reg, stack, shared, stl, ldl, n_i, n_w,
                                                             100.0 * n_w / max(n_i, 1)))


if __name__ == "__main__":
    sys.exit(main())
