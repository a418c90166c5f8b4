"""Histogram of the Blackwell-native SASS mnemonics per kernel of libkemr.so (`cuobjdump -sass`): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, mbarrier -> SYNCS, legacy tensor path -> HMMA,
plus register count per thread.  Runs without a GPU.

    python tools/sass_summary.py [libkemr.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "knowledge_enhanced_multimodal_retrieval_b200", "libkemr.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDGSTS",
        "ATOMS", "ATOMG", "SHFL", "DFMA", "F2F", "FFMA", "LDG", "STG", "LDS", "STS", "UCGABAR", "ELECT", "VOTE", "MEMBAR")


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:            # noqa: BLE001
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "UTMALDG", "LDTM") and m.group(2):
                cur[m.group(1) + m.group(2)] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, architectures {sorted(arch)}")
    print("# SASS mnemonic counts per kernel (static instruction counts, not executions)")
    for k, c in kernels.items():
        total = sum(v for kk, v in c.items() if "." not in kk)
        hot = {kk: c[kk] for kk in KEYS if c.get(kk)}
        detail = {kk: v for kk, v in c.items() if "." in kk}
        short = re.sub(r"\(.*", "", names[k]).replace("kemr::", "").replace("void ", "")
        print(f"{short:60s} regs {regs.get(k, '?'):>4s} instr {total:6d}  " + " ".join(f"{kk}={v}" for kk, v in hot.items())
              + ("  | " + " ".join(f"{kk}={v}" for kk, v in sorted(detail.items())) if detail else ""))


if __name__ == "__main__":
    main()
