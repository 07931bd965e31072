"""Count the Blackwell-specific and atomic SASS mnemonics per kernel of the built library (runs without a GPU):

    python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, FFMA2 = packed fp32 FMA, ATOM / ATOMG / RED = atomics."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "protograd.jl_b200", "libprotograd_b200.so")
WANT = {"UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UBLKCP", "UTCCP", "SYNCS", "FFMA2", "HMMA", "ATOM", "ATOMG", "ATOMS", "RED"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur is not None and m.group(1).split(".")[0] in WANT:
            cur[m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    body = []
    for name, c in zip(names, per.values()):
        if not c:
            continue
        total.update(c)
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        body.append(name[:120] + "\n    " + ", ".join(f"{o} x{n}" for o, n in sorted(c.items())))
    fp_atomics = [o for o in total if o.split(".")[0] in ("ATOM", "ATOMG", "ATOMS", "RED") and (".F32" in o or ".F64" in o or ".F16" in o)]
    print("# cuobjdump -sass protograd.jl_b200/libprotograd_b200.so: Blackwell-specific and atomic mnemonics per kernel (tools/sass_mnemonics.py)")
    print("# floating-point atomics in the whole library:", ", ".join(fp_atomics) if fp_atomics else "none (integer tickets / tile counters / timing only)")
    print("\nTOTAL: " + ", ".join(f"{o} x{n}" for o, n in sorted(total.items())) + "\n")
    print("\n".join(body))


if __name__ == "__main__":
    main()
