"""Resource table of every kernel in libvvae from `nvcc -Xptxas=-v` logs (registers, spills, static shared memory).

    for f in video_vae_b200/csrc/*.cu; do nvcc <flags of video_vae_b200/build.py> -Xptxas=-v -c $f -o /tmp/ptxas/$(basename $f).o \
        > /tmp/ptxas/$(basename $f).log 2>&1; done
    python scripts/ptxas_report.py /tmp/ptxas > profiles/r02zzz_ptxas_resources.md

No GPU needed: this is the check B200_PROFILING.md asks for before spending GPU time (spills, register ceilings)."""
import glob
import os
import re
import subprocess
import sys


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(d):
    rows = []
    for log in sorted(glob.glob(os.path.join(d, "*.log"))):
        src = os.path.basename(log)[:-4]
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                             r"(\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", text):
            name, stack, st, ld, regs, rest = m.groups()
            smem = re.search(r"(\d+) bytes smem", rest)
            bar = re.search(r"used (\d+) barriers", rest)
            rows.append((src, name, int(regs), int(stack), int(st), int(ld), int(smem.group(1)) if smem else 0,
                         int(bar.group(1)) if bar else 0))
    dm = demangle([r[1] for r in rows])
    print("# ptxas resource usage of every kernel in libvvae.so (sm_100a, nvcc 12.9, flags of video_vae_b200/build.py)\n")
    print(f"{len(rows)} kernels (template instantiations counted separately).  Spilling kernels: "
          f"{sum(1 for r in rows if r[4] or r[5])}.  Kernels with a stack frame: {sum(1 for r in rows if r[3])}.\n")
    print("| source | kernel | registers | stack B | spill st / ld B | static smem B | barriers |")
    print("|---|---|---|---|---|---|---|")
    for src, name, regs, stack, st, ld, smem, bar in rows:
        short = re.sub(r"\(.*$", "", dm.get(name, name)).replace("void ", "").replace("vvae::", "").replace("(anonymous namespace)::", "")
        print(f"| {src} | `{short}` | {regs} | {stack} | {st} / {ld} | {smem} | {bar} |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/ptxas")
