"""Regenerates the committed golden fixtures. Run in the build container (needs oracle/_ref, i.e. /root/reference):

    python tests/golden/make_golden.py

  toy/       the reference's own fixture triple (prover-service/resources/toy_circuit/{toy_1.zkey,toy.wtns,toy_vk.json})
             + expected.json = what the reference prover outputs for it with the blinding scalars fixed
  syn256/    a 300-constraint / 256-wire / domain-512 synthetic circuit from oracle/bn254.py (seed 11), its witness,
             and the reference prover's outputs for it
  field_kats.json   see extract_field_kats.py
expected.json holds: r, s (hex LE), proof JSON string, H coefficients (hex), the five MSM results (hex, 384 bytes),
a and b after the SpMV (hex), public inputs.
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bn254 as o  # noqa: E402
import refutil  # noqa: E402

R_FIXED = 0x1D0A8FF4C8E5744C06FD9959908F97ECDFE72D24FCDEF34E00D1C7F8BB929DBB % (o.R_MOD >> 2)
S_FIXED = 0x0123456789ABCDEFFEDCBA98765432100F1E2D3C4B5A69788796A5B4C3D2E1F0 % (o.R_MOD >> 2)


def expected(ref, zkey, wtns, domain, public):
    r, s = o.le32(R_FIXED), o.le32(S_FIXED)
    js, _ = ref.prove(zkey, wtns, r, s)
    ab, h, msm = ref.dump(zkey, wtns, domain, want_ab=True)
    return {"r": r.hex(), "s": s.hex(), "proof": js, "h": h.hex(), "msm": msm.hex(), "ab": ab.hex(), "public": public}


def main():
    ref = refutil.load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    toy_src = "/root/reference/prover-service/resources/toy_circuit"
    toy = os.path.join(HERE, "toy")
    os.makedirs(toy, exist_ok=True)
    for f in ("toy_1.zkey", "toy.wtns", "toy_vk.json"):
        if os.path.exists(os.path.join(toy_src, f)):
            shutil.copyfile(os.path.join(toy_src, f), os.path.join(toy, f))
            os.chmod(os.path.join(toy, f), 0o644)
    exp = expected(ref, os.path.join(toy, "toy_1.zkey"), os.path.join(toy, "toy.wtns"), 4, [2])
    json.dump(exp, open(os.path.join(toy, "expected.json"), "w"), indent=0)

    syn = os.path.join(HERE, "syn256")
    os.makedirs(syn, exist_ok=True)
    r1cs, w = o.synth_circuit(300, 256, seed=11)
    assert o.check_r1cs(r1cs, w)
    zk, trap = o.trapdoor_setup(r1cs, seed=11)
    o.write_zkey(os.path.join(syn, "syn256.zkey"), zk)
    o.write_wtns(os.path.join(syn, "syn256.wtns"), w)
    exp = expected(ref, os.path.join(syn, "syn256.zkey"), os.path.join(syn, "syn256.wtns"), zk.domain_size, [w[1]])
    # independent check of the fixture: the trapdoor formula predicts the same proof
    pa, pb, pc = o.trapdoor_expected_proof(zk, trap, w, R_FIXED, S_FIXED)
    assert o.proof_json(pa, pb, pc) == exp["proof"], "reference output disagrees with the trapdoor prediction"
    json.dump(exp, open(os.path.join(syn, "expected.json"), "w"), indent=0)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
