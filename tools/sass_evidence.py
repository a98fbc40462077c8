#!/usr/bin/env python
"""SASS evidence for the tensor-core / TMA / TMEM kernels of libb2g.so (cuobjdump -sass, sm_100a):
  python tools/sass_evidence.py > profiles/r2_sass_tensor_core_kernels.txt
Per kernel that contains a tcgen05 MMA: the instruction count and how often each tensor-core / TMA / TMEM / bulk-copy
mnemonic occurs, with the first four MMA lines."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-modal-gnn_b200", "libb2g.so")
PAT = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);")
KEEP = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "FENCE.VIEW.ASYNC",
        "UTMACMDFLUSH", "UTMACCTL", "STG.E.ENL2.256")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    print("# SASS evidence (cuobjdump -sass multi-modal-gnn_b200/libb2g.so, sm_100a) -- tensor-core / TMA / TMEM instructions per kernel")
    print("# UTCHMMA = tcgen05.mma (kind::tf32 and kind::f16 share the mnemonic; the kind is in the instruction descriptor), UTMALDG / UTMASTG =")
    print("# cp.async.bulk.tensor load / store (TMA), UBLKCP = cp.async.bulk (contiguous), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,")
    print("# UTCATOMSWS = tcgen05.alloc, SYNCS = mbarrier\n")
    name, body = None, []
    funcs = []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs.append((name, body))
            name, body = m.group(1), []
        elif name:
            mm = PAT.match(line)
            if mm:
                body.append((line.rstrip(), mm.group(1)))
    if name:
        funcs.append((name, body))
    for name, body in sorted(funcs):
        ops = [b.split()[1] if b.startswith("@") else b.split()[0] for _, b in body]
        if not any(o.startswith("UTCHMMA") or o.startswith("UTCQMMA") for o in ops):
            continue
        cnt = collections.Counter(o for o in ops if any(o.startswith(k) for k in KEEP))
        print(f"## {name}  ({len(body)} instructions)")
        for o, c in sorted(cnt.items()):
            print(f"   {o:40s} x{c}")
        shown = 0
        for raw, b in body:
            if "UTCHMMA" in b and shown < 4:
                print("    " + raw.split("/*", 1)[0].strip() + " " + raw[raw.index("/*"):].split("/* 0x")[0].strip())
                shown += 1
        print()


if __name__ == "__main__":
    main()
