"""Counts of the SASS mnemonics that identify each code path, per kernel of libopnet_b200.so (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, mma.sync -> HMMA).  Runs here (cuobjdump, no GPU)."""
import collections, os, re, subprocess, sys
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "objectpermanence_b200", "lib", "libopnet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "HMMA", "FFMA", "MUFU", "LDGSTS", "SYNCS", "REDG", "ATOMG"]
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        counts[name] = collections.Counter()
        continue
    if name:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in MNEMONICS:
                counts[name][op] += 1
print(f"{'kernel':70s} " + " ".join(f"{m:>8s}" for m in MNEMONICS))
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UBLKCP", "HMMA")     # "--all": every kernel, else the tensor-core / TMA ones
for k, c in counts.items():
    if sum(c.values()) and ("--all" in sys.argv or any(c[m] for m in KEY)):
        print(f"{k[:70]:70s} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
