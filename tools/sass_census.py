"""SASS census of libeffimvs.so: per cubin, how often the mnemonics that prove the Blackwell paths occur (no GPU needed).

    python tools/sass_census.py > profiles/<tag>_sass_census.md
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
        ("FFMA2/FMUL2/FADD2", r"\bF(FMA|MUL|ADD)2\b"), ("LDG.256", r"LDG\.E\.ENL2\.256"), ("PREEXIT", r"\bPREEXIT"), ("ACQBULK", r"\bACQBULK"),
        ("HMMA", r"\bHMMA")]


def main():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "effi-mvs-plus_b200", "libeffimvs.so")], cwd=tmp, check=True,
                   capture_output=True)
    print("# SASS census of libeffimvs.so\n")
    print("`cuobjdump -xelf all effi-mvs-plus_b200/libeffimvs.so`, then per cubin `cuobjdump -sass` and a count per mnemonic (tools/sass_census.py).")
    print("tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor (TMA) -> UTMALDG, cp.async.bulk -> UBLKCP,")
    print("packed fp32 (fma / mul / add.rn.f32x2) -> FFMA2 / FMUL2 / FADD2, 256-bit global loads -> LDG.E.ENL2.256, griddepcontrol.launch_dependents /")
    print(".wait (programmatic dependent launch) -> PREEXIT / ACQBULK.  HMMA (legacy mma.sync / wmma) must be absent.\n")
    print("| cubin | instructions | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    tot = [0] * (len(COLS) + 1)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        lines = [ln for ln in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln)]
        row = [len(lines)] + [sum(1 for ln in lines if re.search(p, ln)) for _, p in COLS]
        tot = [a + b for a, b in zip(tot, row)]
        print("| {} | {} |".format(f.split(".")[0], " | ".join(str(v) for v in row)))
    print("| **total** | {} |".format(" | ".join(str(v) for v in tot)))


if __name__ == "__main__":
    sys.exit(main())
