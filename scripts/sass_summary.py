"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA use per kernel of libnb200.so (cuobjdump -sass; no GPU needed).
    python scripts/sass_summary.py > profiles/r2_sass_tcgen05.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "numpower_b200", "libnb200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "USETMAXREG", "ACQBULK",
        "LDG.E.128", "STG.E.128", "LDGSTS", "HMMA", "FFMA", "REDUX", "SHFL", "BAR.SYNC", "ATOM", "RED."]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k) or (k.endswith(".") and k[:-1] == op.split(".")[0]):
                    per[cur][k] += 1
            if op.startswith("UTCHMMA") and "2CTA" in line:
                per[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and "2CTA" in line:
                per[cur]["UTMALDG.2CTA"] += 1
            if op.startswith("UTCBAR") and "MULTICAST" in line:
                per[cur]["UTCBAR.MULTICAST"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# UTCHMMA = tcgen05.mma (kind::f16 / tf32), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA load),")
    print("# UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg\n")
    tot = collections.Counter()
    fam = collections.OrderedDict()   # HBM-bound kernel families: instantiations aggregated by template name
    for (mangled, cnt), name in zip(per.items(), names):
        name = re.sub(r"\(CUtensorMap_st.*", "", name).replace("nb200::", "").replace("void ", "")
        keys = [k for k in cnt if k != "_total" and cnt[k]]
        if not keys:
            continue
        tot.update({k: cnt[k] for k in keys})
        if any(k.startswith(("UTC", "LDTM", "STTM", "UTMA")) for k in keys):
            print(f"{name[:110]:<112} instr {cnt['_total']:>6}  " + "  ".join(f"{k} {cnt[k]}" for k in sorted(keys)))
        else:
            base = re.sub(r"[<(].*", "", name)
            f = fam.setdefault(base, collections.Counter())
            f["_kernels"] += 1
            f.update(cnt)
    print("\n# kernels without tensor-core / TMA instructions (HBM-bound families), instantiations summed per template:")
    for base, cnt in fam.items():
        keys = [k for k in cnt if not k.startswith("_") and cnt[k]]
        print(f"{base[:60]:<62} x{cnt['_kernels']:<4} instr {cnt['_total']:>7}  " + "  ".join(f"{k} {cnt[k]}" for k in sorted(keys)))
    print("\n# library totals: " + "  ".join(f"{k} {tot[k]}" for k in sorted(tot)))


if __name__ == "__main__":
    main()
