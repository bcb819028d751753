"""SASS evidence for the bulk-async NTT staging: python scripts/sass_excerpt.py > profiles/r02_sass_ntt_staging.txt
Dumps, for every NTT kernel in the built library, the copy-engine / mbarrier instructions with their addresses and
an instruction-class histogram (cuobjdump -sass on keyless-zk-proofs_b200/libkzp_b200.so)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "keyless-zk-proofs_b200", "libkzp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, body = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        body[fn] = []
    elif fn and "/*" in line and ";" in line:
        body[fn].append(line.rstrip())
print("cuobjdump -sass keyless-zk-proofs_b200/libkzp_b200.so (sm_100a), NTT kernels; stamp %s" %
      open(os.path.join(ROOT, "keyless-zk-proofs_b200", "build", "stamp.sha256")).read().strip()[:16])
KEY = re.compile(r"UTMALDG|UBLKCP|SYNCS|FENCE\.VIEW\.ASYNC|UTMAPF|LDGSTS")
for fn, lines in body.items():
    if "ntt" not in fn:
        continue
    demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(1)
            cls = ("IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0])
            ops[cls] += 1
    print("\n== %s\n   %d instructions; %s" % (demangled[:150], len(lines), ", ".join(
        "%s %d" % (k, ops[k]) for k in ("IMAD.WIDE", "IMAD", "IADD3", "LDG", "STG", "LDS", "STS", "LDL", "STL", "UTMALDG", "UBLKCP", "SYNCS", "BAR", "WARPSYNC") if ops[k])))
    for l in lines:
        if KEY.search(l):
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
