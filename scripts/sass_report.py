"""Per-kernel count of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md), from `cuobjdump -sass`.

    python scripts/sass_report.py video_vae_b200/libvvae.so > profiles/r02zzz_sass_mnemonics.md      (no GPU needed)"""
import collections
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2")


def main(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k in KEYS:
            if (re.search(r"(?<![A-Z])HMMA", line) if k == "HMMA" else k in line):
                counts[cur][k] += 1
                if k == "UTCHMMA" and ".2CTA" in line:
                    counts[cur]["UTCHMMA.2CTA"] += 1
    names = list(counts)
    dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = {n: re.sub(r"\(.*$", "", d).replace("void ", "").replace("vvae::", "").replace("(anonymous namespace)::", "")
             for n, d in zip(names, dm)}
    fam = collections.OrderedDict()
    for n in names:
        base = re.sub(r"<.*$", "", short[n])
        f = fam.setdefault(base, [0, collections.Counter()])
        f[0] += 1
        f[1].update(counts[n])
    tot = collections.Counter()
    for _, c in fam.values():
        tot.update(c)
    cols = KEYS[:1] + ("UTCHMMA.2CTA",) + KEYS[1:]
    print(f"# SASS mnemonics per kernel family in {lib} (sm_100a; instantiations summed)\n")
    print("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / "
          "UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = mma.sync (the one-warp temporal "
          "attention kernels only), MUFU.EX2 = exp2.\n")
    print(f"Totals over {len(names)} kernels: " + ", ".join(f"{k} {tot[k]}" for k in cols) + "\n")
    print("| kernel family | instantiations | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for base, (n, c) in fam.items():
        if not any(c[k] for k in cols if k not in ("SYNCS", "MUFU.EX2")):
            continue
        print(f"| `{base}` | {n} | " + " | ".join(str(c[k]) for k in cols) + " |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "video_vae_b200/libvvae.so")
