"""Device-side timeline of one nd::matmul call (nb200_trace_enable: %globaltimer stamps written by the kernels themselves).
Unlike an ncu launch list this shows the kernels as they overlap under programmatic dependent launch.

    python scripts/gemm_timeline.py [precision] MxKxN ...      (NB200_PDL=0 in the environment switches PDL off)
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import numpower_b200 as nb

SLOTS = ["prep_first_cta_start", "prep_phaseA_done", "prep_phaseB1_done", "prep_barrier_passed", "prep_last_cta_done",
         "gemm_first_cta_enter", "gemm_first_cta_past_wait", "gemm_last_cta_done", "post_enter", "post_done",
         "fallback_enter", "fallback_done"]
FIRST = {0, 5, 6, 8, 10}


def main():
    prec = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
    sizes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:] if "x" in s] or [(4096, 4096, 4096)]
    lib = nb.lib()
    assert lib.nb200_init(0) == 0
    assert lib.nb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    for (M, K, N) in sizes:
        g = torch.Generator(device="cuda").manual_seed(M + K + N)
        a = torch.rand(M, K, device="cuda", generator=g)
        b = torch.rand(K, N, device="cuda", generator=g)
        c = torch.empty(M, N, device="cuda")
        call = lambda: lib.nb200_sgemm(c.data_ptr(), a.data_ptr(), b.data_ptr(), M, N, K, K, N, N, prec)
        for _ in range(5):
            assert call() == 0, lib.nb200_last_error()
        torch.cuda.synchronize()
        init = torch.tensor([(1 << 63) - 1 if i in FIRST else 0 for i in range(16)], dtype=torch.int64, device="cuda")
        runs = []
        for _ in range(3):
            tr = init.clone()
            assert lib.nb200_trace_enable(C.c_void_p(tr.data_ptr())) == 0
            assert call() == 0
            torch.cuda.synchronize()
            v = tr.cpu().tolist()
            t0 = min(x for i, x in enumerate(v[:12]) if x not in (0, (1 << 63) - 1))
            runs.append({name: (round((v[i] - t0) / 1e3, 2) if v[i] not in (0, (1 << 63) - 1) else None) for i, name in enumerate(SLOTS)})
        lib.nb200_trace_enable(None)
        # the same stamps inside a back-to-back loop (calls 15 and 16 of 20): sustained clocks, and the gap between two calls
        tr1, tr2 = init.clone(), init.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(20):
            if i == 15: lib.nb200_trace_enable(C.c_void_p(tr1.data_ptr()))
            elif i == 16: lib.nb200_trace_enable(C.c_void_p(tr2.data_ptr()))
            elif i == 17: lib.nb200_trace_enable(None)
            call()
        e1.record()
        torch.cuda.synchronize()
        v1, v2 = tr1.cpu().tolist(), tr2.cpu().tolist()
        loop = {name: (round((v1[i] - v1[0]) / 1e3, 2) if v1[i] not in (0, (1 << 63) - 1) else None) for i, name in enumerate(SLOTS)}
        loop["next_call_prep_start"] = round((v2[0] - v1[0]) / 1e3, 2)
        print(json.dumps({"M": M, "K": K, "N": N, "precision": prec, "pdl": os.environ.get("NB200_PDL", "1"),
                          "ms_per_call_back_to_back": e0.elapsed_time(e1) / 20, "timeline_us": runs[-1], "timeline_in_loop_us": loop}), flush=True)


if __name__ == "__main__":
    main()
