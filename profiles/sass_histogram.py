"""Per-kernel SASS opcode histogram of the shipped library: the mnemonics that prove a Blackwell-native
kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP/UBLKRED,
legacy mma.sync/wmma -> HMMA).  Runs on the CPU build box:   python profiles/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flash_cosine_sim_attention_b200", "libfcsa_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "UTMAPF", "LDTM", "STTM",
         "MUFU.EX2", "MUFU.RCP", "MUFU.RSQ", "MUFU.LG2", "REDG", "ATOMG", "SYNCS", "HMMA", "HGMMA", "LDGSTS", "FFMA2", "F2FP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   ({len(kernels)} kernels; counts of static instructions)")
    cols = [w for w in WATCH if any(c[w] for c in kernels.values())]
    print("kernel".ljust(78) + "".join(c.rjust(11) for c in ["instrs"] + cols))
    tot = collections.Counter()
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", demangle(name)).replace("fcsa::", "").replace("void ", "")
        print(short[:77].ljust(78) + "".join(str(c[k]).rjust(11) for k in ["_total"] + cols))
        tot.update(c)
    print("TOTAL".ljust(78) + "".join(str(tot[k]).rjust(11) for k in ["_total"] + cols))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core path found"


if __name__ == "__main__":
    sys.exit(main())
