"""Extracts the limb-level known-answer vectors of the reference's raw field API from
rust-rapidsnark/rapidsnark/src/test_prover.cpp ('Field KATs', e.g. Fr_Rw_mul_unit_test :340-412,
Fq_Rw_mul_unit_test :13352) into tests/golden/field_kats.json. Run in the build container only
(needs /root/reference); the JSON is committed so the GPU box does not need the reference tree.

Only the cases whose compare_Result() call pairs a vector with its own expected value are kept (the reference
file compares some cases against the wrong expected array, test_prover.cpp:406-411 — those are vacuous there).
"""
import json
import re
import sys

SRC = "/root/reference/rust-rapidsnark/rapidsnark/src/test_prover.cpp"
OPS = {"mul": "mul", "Msquare": "square", "add": "add", "sub": "sub", "neg": "neg", "toMontgomery": "to_montgomery",
       "fromMontgomery": "from_montgomery"}


def main():
    text = open(SRC).read()
    out = []
    for m in re.finditer(r"void (F[rq])_Rw_(\w+?)_unit_test\(\)\s*\{(.*?)\n\}", text, re.S):
        field, opname, body = m.group(1), m.group(2), m.group(3)
        if opname not in OPS:
            continue
        arrays = {}
        for a in re.finditer(r"F[rq]RawElement\s+(\w+)\s*=\s*\{([^}]*)\}", body):
            vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", a.group(2))]
            if len(vals) == 4:
                arrays[a.group(1)] = sum(v << (64 * i) for i, v in enumerate(vals))
        for c in re.finditer(r"compare_Result\(\s*(\w+),\s*(\w+),\s*(\w+),\s*(?:(\w+),\s*)?(\d+),", body):
            exp, got, a_name, b_name, idx = c.groups()
            if got != exp + "_c":
                continue  # mismatched expected/actual pairing in the reference file: skip
            if not (a_name.endswith(idx) and exp.endswith(idx)):
                continue
            if exp not in arrays or a_name not in arrays:
                continue
            rec = {"field": field, "op": OPS[opname], "case": int(idx), "a": hex(arrays[a_name]),
                   "expected": hex(arrays[exp])}
            if b_name and b_name in arrays and b_name != a_name:
                rec["b"] = hex(arrays[b_name])
            out.append(rec)
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/field_kats.json", "w"), indent=0)
    print(len(out), "vectors")


if __name__ == "__main__":
    main()
