"""ORACLE build step — test infrastructure only.

Writes the definitions of named member functions of a registration class (pclomp:: / pclpca::NormalDistributionsTransform,
pclomp_ground::NormalDistributionsTransformGround), exactly as they stand in the reference's *_impl2.hpp / ndt_ground_impl.hpp, to a temporary include file that oracle/ndt_ref_harness.cpp compiles (oracle/build_ref.sh).  Nothing is
written into the repository: the output path is a temporary directory of the build.

usage: extract_ref_functions.py <impl.hpp> <out.inc> <qualified class, e.g. pclomp::NormalDistributionsTransform> name [name ...]
"""
import re
import sys


def extract(text, cls, name):
    out = []
    prefix = cls if "<" in cls else cls + "<PointSource, PointTarget>"
    pat = re.compile(re.escape(prefix) + r"::" + re.escape(name) + r"\s*\(")
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
    src, dst, cls, names = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
    text = open(src, encoding="utf-8", errors="replace").read()
    with open(dst, "w", encoding="utf-8") as f:
        for n in names:
            for body in extract(text, cls, n):
                f.write("// ---- %s, from %s\n%s\n\n" % (n, src, body))


if __name__ == "__main__":
    main()
