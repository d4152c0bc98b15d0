"""SASS excerpts for profiles/: a window of instructions around the characteristic mnemonic of each hot kernel
(cuobjdump -sass of the in-tree sm_100a library).  Usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_retinanet_b200", "lib", "librn_b200.so")
# (substring of the mangled name, mnemonic to centre on, lines before, lines after, what to look at)
WANT = [
    ("loss_kernelILi4ELb0ELb1ELb0ELb0E", "MUFU.EX2", 14, 40, "forward-only loss, mid path: 128-bit streaming loads (LDG.E.NA.128), VOTE.ALL, "
     "one MUFU.EX2 + FFMA Horner chain per element, no MUFU.RCP"),
    ("loss_kernelILi4ELb1ELb1ELb0ELb1E", "STG.E.NA.128", 30, 8, "fused loss+filter with gradients: MUFU.EX2 + MUFU.RCP per element, streaming "
     "128-bit gradient store"),
    ("score_filter_kernelILi4ELb1E", "LDG.E.NA.128", 4, 30, "score filter: 4 x 128-bit streaming loads in flight, FMNMX tree, one FSETP per vector"),
    ("lazy2_nms_kernel", "REDUX", 12, 12, "per-class NMS: fixed-point sweep with one warp-wide OR reduction (REDUX) per sweep"),
    ("lazy2_nms_kernel", "MATCH.ANY", 6, 10, "stable regrouping by class: MATCH.ANY per 32-rank chunk"),
    ("loss_finalize_kernel", "STG.E.64.STRONG.SYS", 10, 12, "multi-GPU exchange: (seq << 32 | value) words stored system-scope into the peers' "
     "slots, system-scope spin loads on the local slots"),
    ("match_kernelILb1E", "VOTE.ANY", 10, 14, "matcher: warp-cooperative culling (ballot over 32 GT boxes)"),
]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
    print("# SASS excerpts of pytorch_retinanet_b200/lib/librn_b200.so (sm_100a), cuobjdump -sass; see tools/sass_excerpt.py")
    for sub, mnem, before, after, note in WANT:
        name = next((f for f in funcs if sub in f), None)
        if name is None:
            print(f"\n## {sub}: not found")
            continue
        ins = funcs[name]
        idx = next((i for i, l in enumerate(ins) if mnem in l), None)
        print(f"\n## {name}\n# {note}")
        if idx is None:
            idx = next((i for i, l in enumerate(ins) if mnem.split('.')[0] in l), 0)
        for l in ins[max(0, idx - before): idx + after]:
            print(l)


if __name__ == "__main__":
    main()
