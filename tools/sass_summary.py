"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05 MMA
(UTCHMMA / .2CTA), TMEM loads / stores (LDTM / STTM), TMA loads / stores (UTMALDG / UTMASTG), mbarrier / cluster ops and
the legacy tensor path (HMMA).   python tools/sass_summary.py > profiles/r02_sass_summary.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "UCGABAR", "HMMA",
             "MUFU.EX2", "F2FP.SATFINITE", "ACQBULK", "CCTL"]


def main():
    for variant in ("fp16", "bf16"):
        so = os.path.join(ROOT, "spegnet_b200", f"libspegnet_b200_{variant}.so")
        out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
        kernels = collections.OrderedDict()
        cur = None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = cur.replace("spg::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
                cur = re.sub(r"^void ", "", cur)
                cur = re.sub(r"\((?!.*>).*$", "", cur)  # drop the argument list, keep template arguments
                kernels[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    kernels[cur][mn] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
        print(f"== libspegnet_b200_{variant}.so: {len(kernels)} kernels (cuobjdump -sass, sm_100a)")
        tot = collections.Counter()
        for name, c in kernels.items():
            hits = ", ".join(f"{mn} {c[mn]}" for mn in MNEMONICS if c[mn])
            short = name if len(name) < 110 else name[:107] + "..."
            print(f"{short}: {c['_total']} instr" + (f" | {hits}" if hits else ""))
            tot.update(c)
        print("TOTAL: " + ", ".join(f"{mn} {tot[mn]}" for mn in MNEMONICS if tot[mn]))
        print()


if __name__ == "__main__":
    sys.exit(main())
