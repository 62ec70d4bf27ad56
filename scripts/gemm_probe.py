"""GPU bring-up probe for the tcgen05 GEMM: one kernel variant per subprocess (a trap or timeout
in one variant must not take the others down).  Writes gpurun_out/gemm_probe.jsonl.

    python scripts/gemm_probe.py            # driver: spawns children
    python scripts/gemm_probe.py child <variant> <precision>
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
PNAME = {0: "x3", 1: "x1", 2: "bf16x3", 3: "auto", 4: "fp16x3", 5: "fp16x3u"}
VARIANTS = {"auto": 0, "cg1_bn256": 384, "cg1_bn128": 320, "cg2_bn256": 640, "cg2_bn128": 576}


def child(variant: str, precision: int, sizes):
    import ctypes as C
    import torch
    if VARIANTS[variant]:
        os.environ["NB200_GEMM_VARIANT"] = str(VARIANTS[variant])
    import numpower_b200 as nb
    lib = nb.lib()
    nb._lib.check(lib.nb200_init(0))
    nb._lib.check(lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    res = []
    for (M, K, N) in sizes:
        g = torch.Generator(device="cuda").manual_seed(M + K + N)
        a = torch.rand(M, K, device="cuda", generator=g)
        b = torch.rand(K, N, device="cuda", generator=g)
        c = torch.full((M, N), float("nan"), device="cuda")
        rc = lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), M, N, K, K, N, N, precision)
        torch.cuda.synchronize()
        rec = {"variant": variant, "precision": PNAME[precision], "M": M, "K": K, "N": N, "rc": rc}
        if rc != 0:
            rec["error"] = lib.nb200_last_error().decode()
            res.append(rec)
            continue
        truth = a.double() @ b.double()
        rel = ((c.double() - truth) / truth)
        rec.update(max_rel=float(rel.abs().max()), mean_signed_rel=float(rel.mean()), rms_rel=float(rel.pow(2).mean().sqrt()),
                   nan=int(torch.isnan(c).sum()))
        # what would exact tf32-truncated / tf32-rounded single-pass inputs give?
        if precision == 1 and M <= 1024:
            at = (a.view(torch.int32) & ~0x1FFF).view(torch.float32).double()
            bt = (b.view(torch.int32) & ~0x1FFF).view(torch.float32).double()
            rel_t = ((c.double() - at @ bt) / truth)
            ar = ((a.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).double()
            br = ((b.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).double()
            rel_r = ((c.double() - ar @ br) / truth)
            rec.update(max_rel_vs_truncated_inputs=float(rel_t.abs().max()), mean_vs_truncated=float(rel_t.mean()),
                       max_rel_vs_rounded_inputs=float(rel_r.abs().max()), mean_vs_rounded=float(rel_r.mean()))
        if rec["max_rel"] < 1e-2:
            reps = 20 if M >= 2048 else 50
            for _ in range(3):
                lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), M, N, K, K, N, N, precision)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), M, N, K, K, N, N, precision)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rec.update(ms=ms, useful_tflops=2.0 * M * N * K / ms / 1e9, pipe_tflops=(1 if precision == 1 else 3) * 2.0 * M * N * K / ms / 1e9)
        res.append(rec)
        print(json.dumps(rec), flush=True)
    return res


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        sizes = [(256, 256, 256), (1024, 1024, 1024), (1000, 520, 776), (4096, 4096, 4096)]
        if int(sys.argv[3]) in (2, 4):
            sizes = [(256, 256, 256), (333, 77, 129), (1000, 520, 776), (1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 8192)]
        if len(sys.argv) > 4:
            sizes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[4:]]
        child(sys.argv[2], int(sys.argv[3]), sizes)
        return
    variants = sys.argv[1:] or list(VARIANTS)
    with open(os.path.join(OUT, "gemm_probe.jsonl"), "a") as f:
        for v in variants:
            for prec in [int(x) for x in os.environ.get("PROBE_PRECISIONS", "1,0").split(",")]:
                t0 = time.time()
                try:
                    p = subprocess.run([sys.executable, __file__, "child", v, str(prec)], capture_output=True, text=True, timeout=240)
                    out, err, code = p.stdout, p.stderr[-2000:], p.returncode
                except subprocess.TimeoutExpired as e:
                    out, err, code = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT", -9
                for line in out.splitlines():
                    if line.startswith("{"):
                        f.write(line + "\n")
                f.write(json.dumps({"variant": v, "precision": PNAME[prec], "exit": code, "secs": round(time.time() - t0, 1),
                                    "stderr_tail": err if code != 0 else ""}) + "\n")
                f.flush()
                print(v, prec, "exit", code, flush=True)


if __name__ == "__main__":
    main()
