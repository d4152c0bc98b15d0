"""SASS evidence for profiles/: per kernel of librn_b200.so (sm_100a cubin) the instruction count, the histogram of the
mnemonics that characterise it (128-bit streaming loads/stores, MUFU, REDUX, VOTE, MATCH, atomics, barriers, system-scope
loads/stores of the peer exchange) and register / shared-memory usage from the build log.  Usage:
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_retinanet_b200", "lib", "librn_b200.so")
WANT = ["loss_kernel", "loss_levels_kernel", "loss_finalize_kernel", "exchange_kernel", "match_kernel", "score_filter_kernel",
        "score_filter_levels_kernel", "lazy2_nms_kernel", "lazy_nms_kernel", "nms_kernel", "image_topk_kernel",
        "anchor_grid_kernel", "pack_targets_kernel"]
KEYS = ["LDG.E.64.STRONG.SYS", "STG.E.64.STRONG.SYS", "LD.E.64.STRONG.SYS", "ST.E.64.STRONG.SYS",
        "LDG.E.128", "LDG.E.NA.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.NA.128", "STG.E", "MUFU.EX2", "MUFU.RCP", "MUFU.LG2", "FFMA", "FMUL", "FADD", "REDUX", "VOTE", "MATCH",
        "SHFL", "ATOMS", "ATOMG", "RED", "BAR.SYNC", "LDS", "STS", "DFMA", "DADD", "CS2R", "UTMALDG", "UTCMMA"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    print("# SASS summary of", os.path.relpath(LIB, ROOT), "(cuobjdump -sass; arch line of the cubin below)")
    m = re.search(r"arch = (sm_\w+)", out)
    print("# arch:", m.group(1) if m else "?")
    cur, body = None, collections.OrderedDict()
    for line in out.splitlines():
        f = re.match(r"\s*Function : (\S+)", line)
        if f:
            cur = f.group(1)
            body[cur] = []
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
            if ins:
                body[cur].append(ins)
    for fn, ins in body.items():
        d = demangle(fn)
        short = next((w for w in WANT if re.search(r"\b%s\b" % w, d)), None)
        if not short:
            continue
        hist = collections.Counter()
        for i in ins:
            op = i.split()[0] if not i.startswith("@") else i.split()[1]
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    hist[k] += 1
                    break
        sig = re.sub(r"\(anonymous namespace\)::", "", d)
        sig = sig.split("(")[0]
        print(f"\n## {sig}\n   instructions: {len(ins)}   " + "  ".join(f"{k}:{v}" for k, v in hist.items()))
    print("\n# (no UTMALDG / UTCMMA / LDTM expected: nothing on this path is a dense contraction; streaming is LDG.E.128 with"
          " L1 no-allocate, see DESIGN.md)")


if __name__ == "__main__":
    main()
