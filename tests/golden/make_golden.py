"""Extracts the known-answer vectors the reference's own tests hold for the sketching path into
JSON fixtures (run in the build container, where /root/reference exists; the GPU box only sees
the committed JSON).

  sketches/sketch_test.go:33-76    TestMinimizer: 5 canonical ntHash minimizers (k=5, w=3)
  sketches/sketch_test.go:78-117   TestSyncmer: input + the two commented-out expected values
  sketches/iterator_test.go:31-145 Kmer/Hash/SimHash iterator inputs and expected COUNTS
  seq/codon_tables_test.go:26-134  6 nt -> aa translation vectors (frames 1,-1,-2,-3; tables 1, 11)
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def codon_vectors():
    src = read("seq/codon_tables_test.go")
    out = []
    for m in re.finditer(r"codonTableTest\{(.*?)\n\t\t\}\)", src, re.S):
        body = m.group(1)
        table = int(re.search(r"table:\s*(\d+)", body).group(1))
        lits = re.findall(r"ReplaceAllString\(`(.*?)`", body, re.S)
        nt, aa = (re.sub(r"\s", "", x) for x in lits[:2])
        frame = int(re.search(r"frame:\s*(-?\d+)", body).group(1))
        flag = lambda name: bool(re.search(name + r":\s*true", body))
        out.append(dict(table=table, nt=nt, aa=aa, frame=frame, trim=flag("trim"), clean=flag("clean"),
                        allow_unknown=flag("allowUnknownCodon"), mark_init=flag("markInitCodonAsM")))
    return out


def sketch_vectors():
    src = read("sketches/sketch_test.go")
    t = src[src.index("func TestMinimizer"):src.index("func TestSyncmer")]
    seq = re.search(r'_s := "([ACGTacgt]+)"', t).group(1)
    k = int(re.search(r"k := (\d+)", t).group(1))
    w = int(re.search(r"w := (\d+)", t).group(1))
    vals = [int(x) for x in re.findall(r"codes\[\d\] == (\d{10,})", t)]
    # positions from the trailing comments "// <kmer> idx"? fall back to the oracle-free listing below
    mini = dict(seq=seq, k=k, w=w, values=vals)
    t2 = src[src.index("func TestSyncmer"):src.index("func BenchmarkMinimizerSketch")]
    seq2 = re.search(r'_s := "([ACGTacgt]+)"', t2).group(1)
    k2 = int(re.search(r"k := (\d+)", t2).group(1))
    s2 = int(re.search(r"s := (\d+)", t2).group(1))
    commented = [int(x) for x in re.findall(r"//\s*codes\[\d\] == (\d{10,})", t2)]
    return dict(minimizer=mini, syncmer=dict(seq=seq2, k=k2, s=s2, commented_values=commented))


def iterator_vectors():
    src = read("sketches/iterator_test.go")
    out = {}
    for name in ("TestKmerIterator", "TestHashIterator"):
        i = src.index("func " + name)
        t = src[i:src.index("\nfunc ", i + 10)]
        seq = re.search(r'_s := "([A-Za-z]+)"', t).group(1)
        k = int(re.search(r"k := (\d+)", t).group(1))
        out[name] = dict(seq=seq, k=k, expected_count=len(seq) - k + 1)
    return out


if __name__ == "__main__":
    data = dict(source="shenwei356/bio @ 7b48836e", codon=codon_vectors(), sketch=sketch_vectors(),
                iterator=iterator_vectors())
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(data, f, indent=1)
    print("codon vectors:", len(data["codon"]), "minimizer values:", data["sketch"]["minimizer"]["values"])
    print("syncmer:", {k: v for k, v in data["sketch"]["syncmer"].items() if k != "seq"})
    print("iterator:", {k: (v["k"], v["expected_count"]) for k, v in data["iterator"].items()})
