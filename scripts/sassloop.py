"""Print instruction mix of the hot loop(s) of a kernel in the built library (dev helper)."""
import collections, re, subprocess, sys
lib = sys.argv[3] if len(sys.argv) > 3 else "bio_b200/lib/libb200sketch.so"
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n")[0]
    if pat not in name:
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print(name, len(ins), "instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                j = addr[tgt]
                body = [x[1] for x in ins[j:i + 1]]
                l128 = sum("LDS.128" in b for b in body)
                if (l128 >= 4 or sum("LDS.64" in b for b in body) >= 8) and len(body) < 2500:
                    c = collections.Counter()
                    for b in body:
                        op = b.split()[1] if b.startswith("@") else b.split()[0]
                        c[op.split(".")[0]] += 1
                    l64 = sum("LDS.64" in b for b in body)
                    print(f" loop {j}-{i} len {len(body)} LDS.128={l128} LDS.64={l64}")
                    print("  ", c.most_common())
                    if len(sys.argv) > 2:
                        print("\n".join(body[: int(sys.argv[2])]))
