"""ORACLE build step — test infrastructure only.

Writes the definitions of named member functions of pclomp::NormalDistributionsTransform, exactly as they stand in the reference's
include/ndt_omp/ndt_omp_impl2.hpp, to a temporary include file that oracle/ndt_ref_harness.cpp compiles (oracle/build_ref.sh).  Nothing is
written into the repository: the output path is a temporary directory of the build.

usage: extract_ref_functions.py <ndt_omp_impl2.hpp> <out.inc> name [name ...]
"""
import re
import sys


def extract(text, name):
    out = []
    pat = re.compile(r"pclomp::NormalDistributionsTransform<PointSource, PointTarget>::" + re.escape(name) + r"\s*\(")
    for m in pat.finditer(text):
        start = text.rfind("template", 0, m.start())
        brace = text.index("{", m.end())
        depth, i = 0, brace
        while True:
            c = text[i]
            if c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        out.append(text[start:i + 1])
    if not out:
        raise SystemExit("extract_ref_functions: %s not found" % name)
    return out


def main():
    src, dst, names = sys.argv[1], sys.argv[2], sys.argv[3:]
    text = open(src, encoding="utf-8", errors="replace").read()
    with open(dst, "w", encoding="utf-8") as f:
        for n in names:
            for body in extract(text, n):
                f.write("// ---- %s, from %s\n%s\n\n" % (n, src, body))


if __name__ == "__main__":
    main()
