"""Aggregates ncu SASS-level stall samples of one kernel by CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <kernel substring> <cubin> [source.cu]
Needs the cubin built with -lineinfo (cuobjdump -xelf all lib.so)."""
import csv
import re
import subprocess
import sys


def main(rep, kern, cubin, src=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "# Samples" in r)
    i_addr, i_s, i_src = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
    i_ie = hdr.index("Instructions Executed")
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > i_s and r[i_addr].startswith("0x")]
    base = int(data[0][i_addr], 16)
    samples = {int(r[i_addr], 16) - base: (int(r[i_s]), int(r[i_ie]), r[i_src].strip()) for r in data}
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate the kernel's text section
    start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kern in l)
    line, off2line = (None, 0), {}
    for l in dis[start + 1:]:
        if l.startswith("\t.section"):
            break
        m = re.search(r'//## File "(.*)", line (\d+)', l)
        if m:
            line = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
        if m:
            off2line[int(m.group(1), 16)] = line
    agg = {}
    for off, (s, ie, txt) in samples.items():
        ln = off2line.get(off, (None, -1))
        a = agg.setdefault(ln, [0, 0, 0])
        a[0] += s
        a[1] += ie
        a[2] += 1
    tot = sum(a[0] for a in agg.values())
    cache = {}

    def text_of(f, ln):
        if f is None:
            return ""
        if f not in cache:
            try:
                cache[f] = open(f).read().splitlines()
            except OSError:
                cache[f] = []
        L = cache[f]
        return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""

    print(f"total samples {tot}, instructions {len(samples)}")
    import os
    for (f, ln), (s, ie, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{os.path.basename(f or '?'):16s}:{ln:5d}: {s:6d} samples {100 * s / max(tot, 1):5.1f}%  warp-instr {ie:9d}  sass {n:4d} | {text_of(f, ln)}")


if __name__ == "__main__":
    main(*sys.argv[1:])
