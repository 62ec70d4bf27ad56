"""Runs nb200_debug_tcgen05_probe for the 4 flag combinations and reports which stage of the
tcgen05 path (TMA image, TMEM st/ld, MMA) matches expectations.  Output: gpurun_out/tcgen05_probe.json"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpower_b200 as nb

lib = nb.lib()   # libnb200.so first: the probe library links against it
dbg = C.CDLL(os.path.join(ROOT, "numpower_b200", "libnb200_debug.so"))   # bring-up probes live outside the product library
dbg.nb200_debug_tcgen05_probe.restype = C.c_int
dbg.nb200_debug_tcgen05_probe.argtypes = [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_void_p] * 4
assert lib.nb200_init(0) == 0


def unswizzle_image(img_bytes_as_f32, rows_of_128B, atom32=False):
    """img: flat float32 array of rows_of_128B*32; returns logical [row][32] after undoing the 128B swizzle
    (16B atoms: chunk ^= row%8; 32B atoms: 32B-chunk ^= row%4)."""
    out = np.empty((rows_of_128B, 32), np.float32)
    for r in range(rows_of_128B):
        for c16 in range(8):
            off = r * 128 + c16 * 16
            sw = off ^ ((((off >> 7) & 3) << 5) if atom32 else (((off >> 7) & 7) << 4))
            out[r, c16 * 4:(c16 + 1) * 4] = img_bytes_as_f32[sw // 4: sw // 4 + 4]
    return out


res = []
K = 64
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand(128, K, device="cuda", generator=g)
B = torch.rand(K, 128, device="cuda", generator=g)
Bt = B.t().contiguous()
# (flags, b_layout, lbo, sbo, kstep)
CASES = [(2, 2, 16, 1024, 32),            # B K-major reference point (worked in run 1)
         (0, 1, 4096, 512, 1024), (1, 1, 4096, 512, 1024),     # MN-major, SW128 with 32B atoms: derived setting
         (0, 1, 512, 4096, 1024), (0, 1, 4096, 1024, 1024), (0, 1, 4096, 256, 1024), (0, 1, 1024, 512, 1024)]
for flags, b_layout, lbo, sbo, kstep in CASES:
    smem = torch.full((8192,), -7.0, device="cuda")
    stld = torch.full((128 * 32,), -7.0, device="cuda")
    acc = torch.full((128 * 128,), -7.0, device="cuda")
    info = torch.zeros(16, dtype=torch.int32, device="cuda")
    rc = dbg.nb200_debug_tcgen05_probe(A.data_ptr(), B.data_ptr(), Bt.data_ptr(), K, 128, K, flags, b_layout, lbo, sbo, kstep, smem.data_ptr(), stld.data_ptr(),
                                       acc.data_ptr(), info.data_ptr())
    rec = {"flags": flags, "b_layout": b_layout, "lbo": lbo, "sbo": sbo, "kstep": kstep, "rc": rc, "err": lib.nb200_last_error().decode() if rc else ""}
    if rc == 0:
        torch.cuda.synchronize()
        s = smem.cpu().numpy()
        a_img = unswizzle_image(s[:4096], 128)
        rec["A_smem_matches"] = bool(np.array_equal(a_img, A[:, :32].cpu().numpy()))
        rec["A_smem_first_row"] = a_img[0, :8].tolist()
        rec["A_expected_first_row"] = A[0, :8].cpu().tolist()
        bimg = s[4096:]
        if flags & 2:
            b_img = unswizzle_image(bimg, 128)   # [n][k]
            rec["B_smem_matches"] = bool(np.array_equal(b_img, Bt[:, :32].cpu().numpy()))
        else:
            ok = True
            for j in range(4):
                chunk = unswizzle_image(bimg[j * 1024:(j + 1) * 1024], 32, atom32=(b_layout == 1))  # [k][32 n]
                ok = ok and np.array_equal(chunk, B[:32, j * 32:(j + 1) * 32].cpu().numpy())
            rec["B_smem_matches"] = bool(ok)
        exp_st = (np.arange(128)[:, None] * 100 + np.arange(32)[None, :]).astype(np.float32)
        rec["tmem_st_ld_roundtrip"] = bool(np.array_equal(stld.cpu().numpy().reshape(128, 32), exp_st))
        got = acc.cpu().numpy().reshape(128, 128).astype(np.float64)
        truth = (A[:, :32].double() @ B[:32, :].double()).cpu().numpy()
        rec["acc_max_rel_err"] = float(np.abs(got - truth).max() / np.abs(truth).max())
        rec["acc_sample"] = got[0, :4].tolist()
        rec["truth_sample"] = truth[0, :4].tolist()
        rec["acc_all_zero"] = bool((got == 0).all())
        rec["acc_untouched(-7)"] = bool((got == -7).all())
        rec["info"] = [int(x) & 0xFFFFFFFF for x in info.cpu().tolist()[:9]]
    res.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tcgen05_probe.json"), "w"), indent=1)
