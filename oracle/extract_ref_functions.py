"""ORACLE build step — test infrastructure only.

Writes the definitions of named member functions of a registration class (pclomp:: / pclpca::NormalDistributionsTransform,
pclomp_ground::NormalDistributionsTransformGround), exactly as they stand in the reference's *_impl2.hpp / ndt_ground_impl.hpp, to a temporary include file that oracle/ndt_ref_harness.cpp compiles (oracle/build_ref.sh).  Nothing is
written into the repository: the output path is a temporary directory of the build.

usage: extract_ref_functions.py <impl.hpp> <out.inc> <qualified class, e.g. pclomp::NormalDistributionsTransform, or - for free functions> name [name ...]
"""
import re
import sys


def extract_free(text, name):
    """definitions (not declarations) of a free function `name`, with a preceding template header if there is one"""
    out = []
    for m in re.finditer(r"\b" + re.escape(name) + r"\s*\(", text):
        j = m.end()
        depth = 1
        while depth:                       # matching parenthesis of the parameter list
            c = text[j]
            depth += (c == "(") - (c == ")")
            j += 1
        k = j
        while text[k] in " \t\r\n":
            k += 1
        if text[k] != "{":
            continue                       # a declaration or a call
        start = text.rfind("\n", 0, m.start()) + 1
        head = text[:start].rstrip()
        t = head.rfind("template")
        if t >= 0 and not re.search(r"[;{}]", head[t:]):
            start = t                      # template <...> line(s) directly above
        elif re.search(r"[;{}]\s*$", head) is None and head.rfind("\n") >= 0:
            start = head.rfind("\n") + 1   # return type on the line above
        depth, i = 0, k
        while True:
            c = text[i]
            if c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        line = text[start:m.start()]
        if re.search(r"\b(return|=)\b|[=.]\s*$", line):
            continue                       # `x = name(...) {` cannot happen, but a call inside an expression can precede a brace
        out.append(text[start:i + 1])
    if not out:
        raise SystemExit("extract_ref_functions: free function %s not found" % name)
    return out


def extract(text, cls, name):
    if cls == "-":
        return extract_free(text, name)
    out = []
    plain = cls.startswith("=")            # "=Class": a member of a non-template class, definition starts on the line of its return type
    prefix = cls[1:] if plain else (cls if "<" in cls else cls + "<PointSource, PointTarget>")
    # "name()" selects the overload without parameters
    pat = re.compile(re.escape(prefix) + r"::" + (re.escape(name[:-2]) + r"\s*\(\s*\)" if name.endswith("()") else re.escape(name) + r"\s*\("))
    for m in pat.finditer(text):
        start = text.rfind("\n", 0, m.start()) + 1 if plain else text.rfind("template", 0, m.start())
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
