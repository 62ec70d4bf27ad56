"""Generate tests/golden/phpt_vectors.json from the reference's own phpt tests.

Run HERE (needs /root/reference): ``python tests/golden/make_phpt_vectors.py``.
It parses the --FILE-- section of every hot-path phpt (tests/math/*.phpt,
tests/linalg/001-ndarray-matmul.phpt) into (op, operands, kwargs) records and the
--EXPECT-- section into the printed numbers (PHP precision=14), and stores both.
The committed JSON is what travels; /root/reference is never read at test time.
"""
import ast
import glob
import json
import os
import re
import sys

import numpy as np

REF = os.environ.get("NB200_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "phpt_vectors.json")

NUM = r"-?(?:\d+\.?\d*(?:E[-+]?\d+)?|NAN|INF)"
BINOPS = {"+": "add", "-": "sub", "*": "mul", "/": "div", "%": "mod", "**": "pow"}


def parse_operand(tok, env):
    tok = tok.strip()
    m = re.fullmatch(r"\$(\w+)\[(\d+)\]", tok)
    if m:
        return {"value": np.asarray(env[m.group(1)], dtype=np.float64)[int(m.group(2))].tolist(), "view": True}
    m = re.fullmatch(r"\$(\w+)", tok)
    if m:
        return {"value": env[m.group(1)]}
    return {"value": ast.literal_eval(tok)}


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "[(":
            depth += 1
        if ch in "])":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def parse_expect(text, want):
    """`want` = how many numbers the calls print; PHP's print_r of consecutive scalar
    results concatenates them without a separator (e.g. "422" = 4, 2, 2), so when the
    count is short, all-digit bare tokens are split into single digits."""
    vals = []
    bare = []
    for line in text.splitlines():
        line = line.strip()
        if "=>" in line:
            rhs = line.split("=>", 1)[1].strip()
            if rhs == "Array":
                continue
            vals.append(rhs)
        else:
            line = line.replace("Array", " ")
            for t in re.findall(NUM, line):
                bare.append(len(vals))
                vals.append(t)
    if len(vals) < want:
        out = []
        for i, v in enumerate(vals):
            if i in bare and v.isdigit() and len(v) > 1:
                out.extend(list(v))
            else:
                out.append(v)
        vals = out
    conv = {"NAN": float("nan"), "INF": float("inf"), "-INF": float("-inf")}
    return [conv[v] if v in conv else float(v) for v in vals]


def result_size(rec):
    vs = [np.asarray(o["value"], dtype=np.float64) for o in rec["operands"]]
    op = rec["op"]
    if op in BINOPS.values():
        return int(np.prod(np.broadcast_shapes(*[v.shape for v in vs])))
    if op == "matmul":
        return vs[0].shape[0] * vs[1].shape[1]
    if op in ("sum", "prod", "max", "min"):
        ax = rec["kwargs"].get("axis")
        return 1 if ax is None else int(vs[0].size // vs[0].shape[ax])
    return int(vs[0].size)


def main():
    files = sorted(glob.glob(os.path.join(REF, "tests/math/*.phpt"))) + \
        [os.path.join(REF, "tests/linalg/001-ndarray-matmul.phpt")]
    records = []
    for path in files:
        text = open(path).read()
        body = text.split("--FILE--", 1)[1]
        code, expect = body.split("--EXPECT--", 1)
        env, recs = {}, []
        for line in code.splitlines():
            line = line.strip().rstrip(";")
            m = re.fullmatch(r"\$(\w+) = \\NDArray::array\((.*)\)", line)
            if m:
                env[m.group(1)] = ast.literal_eval(m.group(2))
                continue
            m = re.fullmatch(r"print_r\(\((.+?) (\*\*|[-+*/%]) (.+)\)->toArray\(\)\)", line)
            if m:
                recs.append({"op": BINOPS[m.group(2)], "kwargs": {},
                             "operands": [parse_operand(m.group(1), env), parse_operand(m.group(3), env)]})
                continue
            m = re.fullmatch(r"print_r\(\\NDArray::(\w+)\((.*)\)->toArray\(\)\)", line) or \
                re.fullmatch(r"print_r\(\\NDArray::(\w+)\((.*)\)\)", line)
            if m:
                ops, kw = [], {}
                for a in split_args(m.group(2)):
                    km = re.fullmatch(r"\s*(\w+):\s*(.+)", a)
                    if km:
                        kw[km.group(1)] = ast.literal_eval(km.group(2).strip())
                    else:
                        ops.append(parse_operand(a, env))
                recs.append({"op": m.group(1), "kwargs": kw, "operands": ops})
                continue
            if line and not line.startswith("<?php") and not line.startswith("?>"):
                raise SystemExit(f"unparsed line in {path}: {line}")
        nums = parse_expect(expect, sum(result_size(r) for r in recs))
        pos = 0
        for r in recs:
            n = result_size(r)
            r["expected"] = nums[pos:pos + n]
            r["file"] = os.path.relpath(path, REF)
            pos += n
            records.append(r)
        if pos != len(nums):
            raise SystemExit(f"{path}: consumed {pos} of {len(nums)} expected numbers")

    def enc(o):
        if isinstance(o, float) and o != o:
            return "NAN"
        return o
    def walk(o):
        if isinstance(o, list):
            return [walk(x) for x in o]
        if isinstance(o, dict):
            return {k: walk(v) for k, v in o.items()}
        return enc(o)
    json.dump({"source": "NumPower/numpower tests/*.phpt --EXPECT-- blocks (PHP precision=14)",
               "records": walk(records)}, open(OUT, "w"), indent=1)
    print(f"wrote {len(records)} records to {OUT}")


if __name__ == "__main__":
    sys.exit(main())
