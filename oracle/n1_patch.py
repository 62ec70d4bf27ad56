"""TEST INFRASTRUCTURE — applies the Level-1 host patches to copies of four reference files at BUILD
time: one inserted call per hot-path function, right after its opening brace.  Inputs are read from /root/reference,
outputs go to oracle/_ref/n1_src/ (git-ignored); no reference source is stored in this repo.

    python oracle/n1_patch.py <reference_dir> <out_dir>
"""
import os
import re
import sys

DECL = ('\n#include <nb200.h>\n'
        'NDArray *nb200_glue_binary(int op, NDArray *a, NDArray *b);\n'
        'NDArray *nb200_glue_reduce(NDArray *array, int *axis, NDArray *(*operation)(NDArray *, NDArray *));\n'
        'NDArray *nb200_glue_argminmax(NDArray *op, int axis, bool keepdims, bool is_argmax);\n'
        'NDArray *nb200_glue_matmul(NDArray *a, NDArray *b);\n'
        'NDArray *nb200_glue_max_axis(NDArray *target, int axis);\n'
        'NDArray *nb200_glue_dot(NDArray *nda, NDArray *ndb);\n'
        'NDArray *nb200_glue_to_gpu(NDArray *target);\n'
        'NDArray *nb200_glue_to_cpu(NDArray *target);\n')

# file -> [(regex matching the function's definition line incl. "{", code inserted after it)]
PATCHES = {
    "src/ndmath/arithmetics.c": [
        (r"^NDArray_Add_Float\(NDArray\* a, NDArray\* b\) \{", "NB200_ADD"),
        (r"^NDArray_Subtract_Float\(NDArray\* a, NDArray\* b\) \{", "NB200_SUB"),
        (r"^NDArray_Multiply_Float\(NDArray\* a, NDArray\* b\) \{", "NB200_MUL"),
        (r"^NDArray_Divide_Float\(NDArray\* a, NDArray\* b\) \{", "NB200_DIV"),
        (r"^NDArray_Mod_Float\(NDArray\* a, NDArray\* b\) \{", "NB200_MOD"),
        (r"^NDArray_Pow_Float\(NDArray\* a, NDArray\* b\) \{", "NB200_POW"),
    ],
    "src/ndarray.c": [(r"^reduce\(NDArray \*array, int \*axis, NDArray \*\(\*operation\)\(NDArray \*, NDArray \*\)\) \{", "REDUCE"),
                      (r"^NDArray_MaxAxis\(NDArray \*target, int axis\) \{", "MAXAXIS"),
                      (r"^NDArray_Maximum\(NDArray \*a, NDArray \*b\) \{", "NB200_MAXIMUM"),
                      (r"^NDArray_Minimum\(NDArray \*a, NDArray \*b\) \{", "NB200_MINIMUM"),
                      (r"^NDArray_ToGPU\(NDArray \*target\) \{", "TOGPU"),
                      (r"^NDArray_ToCPU\(NDArray \*target\) \{", "TOCPU")],
    "src/ndmath/calculation.c": [(r"^NDArray_ArgMinMaxCommon\(NDArray \*op, int axis, bool keepdims, bool is_argmax\) \{", "ARG")],
    "src/ndmath/linalg.c": [(r"^NDArray_Matmul\(NDArray \*a, NDArray \*b\) \{", "MATMUL"),
                            (r"^NDArray_Dot\(NDArray \*nda, NDArray \*ndb\) \{", "DOT")],
}


def snippet(kind):
    if kind.startswith("NB200_"):
        return f"    {{ NDArray *nb200_r = nb200_glue_binary({kind}, a, b); if (nb200_r != NULL) return nb200_r; }}\n"
    if kind == "REDUCE":
        return "    { NDArray *nb200_r = nb200_glue_reduce(array, axis, operation); if (nb200_r != NULL) return nb200_r; }\n"
    if kind == "ARG":
        return ("    if (NDArray_DEVICE(op) == NDARRAY_DEVICE_GPU) return nb200_glue_argminmax(op, axis, keepdims, is_argmax);\n")
    if kind == "TOGPU":
        return "    { NDArray *nb200_r = nb200_glue_to_gpu(target); if (nb200_r != NULL) return nb200_r; }\n"
    if kind == "TOCPU":
        return "    { NDArray *nb200_r = nb200_glue_to_cpu(target); if (nb200_r != NULL) return nb200_r; }\n"
    if kind == "MAXAXIS":
        return "    if (NDArray_DEVICE(target) == NDARRAY_DEVICE_GPU) return nb200_glue_max_axis(target, axis);\n"
    if kind == "DOT":
        return "    { NDArray *nb200_r = nb200_glue_dot(nda, ndb); if (nb200_r != NULL) return nb200_r; }\n"
    if kind == "MATMUL":
        return "    { NDArray *nb200_r = nb200_glue_matmul(a, b); if (nb200_r != NULL) return nb200_r; }\n"
    raise ValueError(kind)


def main(ref, out):
    for rel, plist in PATCHES.items():
        text = open(os.path.join(ref, rel)).read()
        lines = text.splitlines(keepends=True)
        done = 0
        res = []
        # declarations go after the file's header block: the line after the last #include of the first 80 lines, moved past the
        # #endif of a conditional block (#ifdef HAVE_GD ... in ndarray.c) the include may sit in
        last_inc = max(i for i, l in enumerate(lines[:80]) if l.startswith("#include"))
        depth = 0
        for l in lines[:last_inc + 1]:
            if re.match(r"#\s*if", l):
                depth += 1
            elif re.match(r"#\s*endif", l):
                depth -= 1
        while depth > 0:
            last_inc += 1
            if re.match(r"#\s*if", lines[last_inc]):
                depth += 1
            elif re.match(r"#\s*endif", lines[last_inc]):
                depth -= 1
        for i, l in enumerate(lines):
            res.append(l)
            if i == last_inc:
                res.append(DECL)
            for rx, kind in plist:
                if re.match(rx, l):
                    res.append(snippet(kind))
                    done += 1
        if done != len(plist):
            raise SystemExit(f"{rel}: applied {done} of {len(plist)} patches (anchor not found)")
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").write("".join(res))
        print(f"patched {rel}: {done} insertion(s)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
