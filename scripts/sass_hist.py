#!/usr/bin/env python
"""SASS listing helper: opcode histogram of one kernel of an object file, whole kernel and
the hot loop (the largest backward-branch span).

    python scripts/sass_hist.py <object> <mangled-or-substring> [--dump out.sass]
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
    funcs = sorted(set(re.findall(r"\.text\.(\S+)", names)))
    cand = [f for f in funcs if pat in f]
    if not cand:
        sys.exit(f"no function matching {pat}; have e.g. {funcs[:5]}")
    fun = cand[0]
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True,
                          text=True).stdout
    if "--dump" in sys.argv:
        open(sys.argv[sys.argv.index("--dump") + 1], "w").write(sass)
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    # hot loop: the backward branch with the largest span
    best = (0, 0, 0)
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr and addr - int(t.group(1), 16) > best[0]:
                best = (addr - int(t.group(1), 16), int(t.group(1), 16), addr)
    print(f"function {fun}: {len(ins)} instructions; hot loop {best[1]:#x}..{best[2]:#x}")
    loop = [i for i in ins if best[1] <= i[0] <= best[2]]
    for title, seq in (("whole kernel", ins), ("hot loop", loop)):
        c = collections.Counter(op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(
            ("IMAD", "ATOMS", "LDS", "STS")) and "." in op else "") for _, op, _ in seq)
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
        print(f"--- {title}: {len(seq)} instructions, {fp64} fp64 (DFMA/DADD/DMUL/DSETP)")
        for k, v in c.most_common():
            print(f"{v:6d}  {k}")


if __name__ == "__main__":
    main()
