"""Per-object counts of the SASS mnemonics that show which hardware paths the kernels use (tcgen05 MMA, TMEM loads,
TMA tensor and bulk copies, mbarriers, dp4a).  usage: python tools/sass_summary.py [--out profiles/r02_sass_summary.txt]"""
import argparse
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = {
    "UTCIMMA (tcgen05.mma kind::i8)": r"\bUTCIMMA",
    "UTCBAR (tcgen05.commit)": r"\bUTCBAR",
    "LDTM (tcgen05.ld)": r"\bLDTM",
    "UTMALDG (cp.async.bulk.tensor)": r"\bUTMALDG",
    "UBLKCP (cp.async.bulk)": r"\bUBLKCP",
    "UBLKPF (cp.async.bulk.prefetch)": r"\bUBLKPF",
    "SYNCS (mbarrier)": r"\bSYNCS",
    "IDP.4A (dp4a)": r"\bIDP\.4A",
    "IMMA (mma.sync m16n8k32 u8)": r"\bIMMA\.16832",
    "DFMA/DMUL/DADD (float64)": r"\bD(FMA|MUL|ADD)\b",
    "REDG/ATOMG (global atomics)": r"\b(REDG|ATOMG|RED\.E|ATOM\.E)",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    lines = ["# SASS mnemonic counts per object (cuobjdump -sass, sm_100a); built by csrc/Makefile", ""]
    objs = sorted(glob.glob(os.path.join(ROOT, "go-vectorsearch_b200", "build", "*.o")))
    hdr = f"{'object':14s}" + "".join(f"{k.split(' ')[0]:>15s}" for k in PAT)
    lines.append(hdr)
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        kernels = len(re.findall(r"^\s*Function : ", sass, flags=re.M))
        row = f"{os.path.basename(o):14s}" + "".join(f"{len(re.findall(p, sass)):15d}" for p in PAT.values())
        lines.append(row + f"   ({kernels} kernels)")
    lines += ["", "legend:"] + [f"  {k}" for k in PAT]
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
