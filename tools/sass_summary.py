"""SASS evidence per kernel of libtdx.so: counts of the Blackwell-native mnemonics (profiles/rNN_sass_summary.txt).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt      (no GPU needed: cuobjdump reads the built .so)
"""
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "terrain_diffusion_b200" / "libtdx.so"
PAT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDUX", "R2UR", "HMMA", "MUFU.TANH",
       "LDG.E.128", "STG.E.128", "BAR.SYNC", "UCGABAR", "ACQBULK", "MEMBAR"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    print(f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (sm_100a)")
    print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, TMA tiled load -> UTMALDG, bulk copy -> "
          "UBLKCP, mbarrier -> SYNCS; HMMA would be the legacy mma.sync path (must be 0)")
    print(f"{'kernel':48s} {'instr':>6s} " + " ".join(f"{p:>9s}" for p in PAT))
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        body = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
        counts = [sum(1 for l in body if re.search(r"\b" + re.escape(p) + r"\b", l) or (("." in p) and p in l)) for p in PAT]
        print(f"{short[-48:]:48s} {len(body):6d} " + " ".join(f"{c:9d}" for c in counts))


if __name__ == "__main__":
    main()
