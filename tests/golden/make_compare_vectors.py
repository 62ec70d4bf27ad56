"""Generate tests/golden/compare_vectors.json from the reference's comparison and transpose phpt tests
(tests/logic/003..008-*.phpt, tests/manipulation/001-ndarray-transpose.phpt).  Run HERE (needs /root/reference):
``python tests/golden/make_compare_vectors.py``.  The committed JSON travels; /root/reference is never read at test time."""
import ast
import json
import os
import re

REF = os.environ.get("NB200_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compare_vectors.json")
FILES = [("logic/003-ndarray-equal.phpt", "equal"), ("logic/004-ndarray-greater.phpt", "greater"),
         ("logic/005-ndarray-greater_equal.phpt", "greater_equal"), ("logic/006-ndarray-less.phpt", "less"),
         ("logic/007-ndarray-less_equal.phpt", "less_equal"), ("logic/008-ndarray-not_equal.phpt", "not_equal"),
         ("manipulation/001-ndarray-transpose.phpt", "transpose")]


def sections(path):
    text = open(path).read()
    return text.split("--FILE--")[1].split("--EXPECT--")[0], text.split("--EXPECT--")[1].strip()


def parse_print_r(text):
    """print_r output of (nested) PHP arrays, several in a row -> list of nested Python lists."""
    toks = re.findall(r"Array|\(|\)|\[\d+\] => ?|-?[\d.]+(?:E[+-]?\d+)?", text)
    pos = 0

    def value():
        nonlocal pos
        if toks[pos] == "Array":
            pos += 1
            assert toks[pos] == "("
            pos += 1
            items = []
            while toks[pos] != ")":
                assert toks[pos].startswith("[")
                pos += 1
                items.append(value())
            pos += 1
            return items
        v = float(toks[pos])
        pos += 1
        return v

    out = []
    while pos < len(toks):
        out.append(value())
    return out


def split_args(text):
    """top-level comma split (array literals nest)"""
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
            continue
        depth += ch == "["
        depth -= ch == "]"
        cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def main():
    vectors = []
    for rel, fn in FILES:
        code, expect = sections(os.path.join(REF, "tests", rel))
        env = {}
        for m in re.finditer(r"\$(\w+)\s*=\s*(?:\\?NDArray|nd)::array\((.*)\);", code):
            env[m.group(1)] = ast.literal_eval(m.group(2))
        calls = []
        for m in re.finditer(r"::%s\((.*?)\)->toArray\(\)" % fn, code):
            args = []
            for tok in split_args(m.group(1)):
                args.append(env[tok[1:]] if tok.startswith("$") else ast.literal_eval(tok))
            calls.append(args)
        got = parse_print_r(expect)
        assert len(calls) == len(got), (rel, len(calls), len(got))
        for args, e in zip(calls, got):
            vectors.append({"file": "tests/" + rel, "op": fn, "args": args, "expect": e})
    json.dump({"source": "NumPower/numpower tests/logic/003..008-*.phpt, tests/manipulation/001-ndarray-transpose.phpt --EXPECT-- blocks",
               "vectors": vectors}, open(OUT, "w"), indent=1)
    print(f"wrote {OUT}: {len(vectors)} vectors")


if __name__ == "__main__":
    main()
