"""Count the instruction mix of the main streaming loop of a kernel (largest backward branch)."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.findall(r"Function : (\S+)", names)
for fn in funcs:
    if not re.search(pat, fn):
        continue
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout.splitlines()
    ins = []
    for l in txt:
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    best = None
    for a, op, rest in ins:
        if op.startswith("BRA"):
            m = re.search(r"(0x[0-9a-f]+)", rest)
            if m:
                t = int(m.group(1), 16)
                if t < a and (best is None or a - t > best[1] - best[0]):
                    best = (t, a)
    if not best:
        continue
    body = [op for a, op, _ in ins if best[0] <= a <= best[1]]
    c = collections.Counter(op.split(".")[0] for op in body)
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(fn[:60], "loop instrs", len(body), "fp64", fp64, dict(c.most_common(12)))
