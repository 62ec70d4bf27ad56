"""Generate tests/golden/logic_vectors.json from the reference's logic phpt tests (tests/logic/001-ndarray-all.phpt,
002-ndarray-allclose.phpt).  Run HERE (needs /root/reference): ``python tests/golden/make_logic_vectors.py``.
The committed JSON travels; /root/reference is never read at test time."""
import ast
import json
import os
import re

import numpy as np

REF = os.environ.get("NB200_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "logic_vectors.json")


def sections(path):
    text = open(path).read()
    return text.split("--FILE--")[1].split("--EXPECT--")[0], text.split("--EXPECT--")[1].strip()


def operand(tok, env):
    tok = tok.strip()
    m = re.fullmatch(r"\$(\w+)\[(\d+)\]", tok)
    if m:
        return np.asarray(env[m.group(1)], dtype=np.float64)[int(m.group(2))].tolist()
    return env[tok.lstrip("$")]


def main():
    vectors = []
    for name, fn in (("001-ndarray-all.phpt", "all"), ("002-ndarray-allclose.phpt", "allclose")):
        code, expect = sections(os.path.join(REF, "tests", "logic", name))
        env = {}
        for m in re.finditer(r"\$(\w+)\s*=\s*(?:\\?NDArray|nd)::array\((.*)\);", code):
            env[m.group(1)] = ast.literal_eval(m.group(2))
        calls = [[operand(t, env) for t in m.group(1).split(",")] for m in re.finditer(r"::%s\(([^()]*)\)" % fn, code)]
        if fn == "all":
            got = [int(ch) for ch in expect]                     # print_r of ints, concatenated
        else:
            got = [int(v == "true") for v in re.findall(r"bool\((\w+)\)", expect)]
        assert len(calls) == len(got), (name, calls, got)
        for args, e in zip(calls, got):
            vectors.append({"file": "tests/logic/" + name, "op": fn, "args": args, "expect": e})
    json.dump({"source": "NumPower/numpower tests/logic/*.phpt --EXPECT-- blocks", "vectors": vectors}, open(OUT, "w"), indent=1)
    print(f"wrote {OUT}: {len(vectors)} vectors")


if __name__ == "__main__":
    main()
