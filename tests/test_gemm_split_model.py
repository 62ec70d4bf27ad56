"""CPU model of the two error-compensated GEMM modes (numpower_b200/csrc/sgemm_tcgen05.cu): the operand splits and the
three retained products are restated in numpy with exact (fp64) accumulation, so what remains is the algorithmic error of
the split itself — the part the kernels cannot do better than.  Checks the bounds DESIGN.md §5 quotes:

  TF32x3:  a = hi + lo, hi = trunc_tf32(a), lo = rna_tf32(a - hi); |lo| <= 2^-10 |a|, remainder <= 2^-21 |a|;
           dropped lo.lo (<= 2^-20) and the two remainders                       -> <= 2^-19 per product: GUARANTEES 1e-5
  BF16x3:  a = a1 + a2 + r, a1 = rn_bf16(a), a2 = rn_bf16(a - a1); |a2| <= 2^-8 |a|, |r| <= 2^-17 |a|;
           dropped a2.b2 (<= 2^-16) and the two remainders                       -> <= 2^-15 per product, zero-mean:
           averages out over K on random data, adds up coherently on constant matrices -> opt-in mode, not NB200_GEMM_AUTO
"""
import numpy as np
import pytest


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def rn_bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32 (cvt.rn.bf16.f32)."""
    b = _bits(x).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def trunc_tf32(x):
    return (_bits(x) & np.uint32(0xFFFFE000)).view(np.float32)


def rna_tf32(x):
    """round-to-nearest, ties away from zero, to TF32 (cvt.rna.tf32.f32)."""
    b = _bits(x).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def split_bf16(a):
    a1 = rn_bf16(a)
    a2 = rn_bf16((a - a1).astype(np.float32))      # a - a1 is exact in fp32
    return a1, a2


def split_tf32(a):
    hi = trunc_tf32(a)
    return hi, rna_tf32((a - hi).astype(np.float32))


def three_products(ah, al, bh, bl):
    d = np.float64
    return ah.astype(d) @ bh.astype(d) + ah.astype(d) @ bl.astype(d) + al.astype(d) @ bh.astype(d)


def test_bf16_split_remainder_bound():
    r = np.random.default_rng(0)
    a = (r.random(1 << 16, dtype=np.float32) * 2 - 1) * np.exp2(r.integers(-20, 20, 1 << 16)).astype(np.float32)
    a1, a2 = split_bf16(a)
    assert ((a - a1).astype(np.float32).astype(np.float64) == a.astype(np.float64) - a1.astype(np.float64)).all()   # exact subtraction
    rem = np.abs(a.astype(np.float64) - a1.astype(np.float64) - a2.astype(np.float64))
    assert (rem <= np.abs(a) * 2.0 ** -17).all()
    assert (np.abs(a2) <= np.abs(a) * 2.0 ** -8).all()
    # both parts really are bf16 values
    assert ((_bits(a1) & 0xFFFF) == 0).all() and ((_bits(a2) & 0xFFFF) == 0).all()


def test_tf32_split_remainder_bound():
    r = np.random.default_rng(1)
    a = (r.random(1 << 16, dtype=np.float32) * 2 - 1) * np.exp2(r.integers(-20, 20, 1 << 16)).astype(np.float32)
    hi, lo = split_tf32(a)
    rem = np.abs(a.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64))
    assert (rem <= np.abs(a) * 2.0 ** -21).all()
    assert (np.abs(lo) <= np.abs(a) * 2.0 ** -10).all()
    assert ((_bits(hi) & 0x1FFF) == 0).all() and ((_bits(lo) & 0x1FFF) == 0).all()


@pytest.mark.parametrize("k", [1, 8, 128, 1024])
def test_bf16x3_algorithmic_error_random_data(k):
    """Per product up to 2^-16 + 2 * 2^-17 = 2^-15; zero-mean, so on random data it shrinks along K."""
    r = np.random.default_rng(k)
    a, b = r.random((96, k), dtype=np.float32), r.random((k, 80), dtype=np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    a1, a2 = split_bf16(a)
    b1, b2 = split_bf16(b)
    signed = (three_products(a1, a2, b1, b2) - exact) / exact
    assert np.abs(signed).max() <= 2.0 ** -15
    if k >= 128:
        assert np.abs(signed).max() <= 3e-6      # what the GPU measurements show as well (1.2-2.5e-6 incl. accumulation)
    assert abs(signed.mean()) <= 2e-6


def test_bf16x3_coherent_inputs_break_1e5_and_tf32x3_does_not():
    """A constant-matrix GEMM is one product repeated K times: its split error does not average out.  This is the reason
    NB200_GEMM_AUTO resolves to TF32x3 (include/nb200.h) and BF16x3 is opt-in."""
    r = np.random.default_rng(7)
    n = 200_000
    a = ((r.random(n, dtype=np.float32) + 0.5) * np.exp2(r.integers(-3, 3, n))).astype(np.float32)
    b = ((r.random(n, dtype=np.float32) + 0.5) * np.exp2(r.integers(-3, 3, n))).astype(np.float32)
    d = np.float64
    exact = a.astype(d) * b.astype(d)
    a1, a2 = split_bf16(a)
    b1, b2 = split_bf16(b)
    e_bf16 = np.abs(a1.astype(d) * b1 + a1.astype(d) * b2 + a2.astype(d) * b1 - exact) / exact
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    e_tf32 = np.abs(ah.astype(d) * bh + ah.astype(d) * bl + al.astype(d) * bh - exact) / exact
    assert e_tf32.max() <= 2.0 ** -19 < 1e-5          # guaranteed
    assert e_bf16.max() <= 2.0 ** -15
    assert e_bf16.max() > 1e-5 and 0.01 < (e_bf16 > 1e-5).mean() < 0.2   # a few per cent of constant pairs violate 1e-5


@pytest.mark.parametrize("k", [1, 8, 1024])
def test_tf32x3_algorithmic_error(k):
    r = np.random.default_rng(100 + k)
    a, b = r.random((96, k), dtype=np.float32), r.random((k, 80), dtype=np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    err = np.abs(three_products(ah, al, bh, bl) - exact) / exact
    assert err.max() <= 2.0 ** -19


def test_bf16_split_special_values_match_the_kernel_rules():
    """split_bf16() in sgemm_tcgen05.cu: +-inf -> (inf, 0); finite values that would round up to inf are truncated."""
    big = np.array([3.4e38, -3.4e38], np.float32)
    assert np.isinf(rn_bf16(big)).all()                      # plain rounding overflows ...
    trunc = (_bits(big) & np.uint32(0xFFFF0000)).view(np.float32)
    assert np.isfinite(trunc).all()                          # ... truncation (what the kernel does) does not
    lo = rn_bf16((big - trunc).astype(np.float32))
    assert (np.abs(big.astype(np.float64) - trunc.astype(np.float64) - lo.astype(np.float64)) <= np.abs(big) * 2.0 ** -16).all()


# ------------------------------------------------------------------ FP16x3 (scaled half parts), numpy restatement
def scale_exp(max_abs):
    """sgemm_tcgen05.cu scale_exp(): power of two bringing the largest finite magnitude into [2^14, 2^15)."""
    m = np.asarray(max_abs, dtype=np.float64)
    e = np.zeros(m.shape, dtype=np.int64)
    ok = (m > 0) & np.isfinite(m)
    e[ok] = 14 - np.floor(np.log2(m[ok])).astype(np.int64)
    return e


def split_f16_scaled(x, e):
    """x * 2^e -> (hi, lo): IEEE half values, lo stored times 2^11 as in split_f16(); returned as float64 with the 2^-11
    already applied.  Also returns the eligibility the kernel computes (every non-zero element >= 2^-14 after scaling)."""
    xs = np.ldexp(x.astype(np.float64), e)                 # exact
    hi = xs.astype(np.float32).astype(np.float16)          # cvt.rn.f16.f32
    lo = ((xs - hi.astype(np.float64)) * 2048.0).astype(np.float32).astype(np.float16)
    eligible = bool(((xs == 0) | (np.abs(xs) >= 2.0 ** -14)).all())
    return hi.astype(np.float64), lo.astype(np.float64) / 2048.0, eligible


def fp16x3_model(a, b):
    ea = scale_exp(np.abs(a).max(axis=1))[:, None]
    eb = scale_exp(np.abs(b).max(axis=0))[None, :]
    ah, al, ok_a = split_f16_scaled(a, ea)
    bh, bl, ok_b = split_f16_scaled(b, eb)
    assert np.abs(ah).max() <= 32768 and np.abs(bh).max() <= 32768      # never overflows fp16
    if not (ok_a and ok_b):
        return None                                                      # the kernel's gated TF32x3 fallback takes over
    acc = ah @ bh + ah @ bl + al @ bh
    return np.ldexp(acc, -(ea + eb))


def test_fp16x3_coherent_inputs_stay_inside_1e5():
    """The case that breaks BF16x3 (one product repeated): 11-bit parts keep it at TF32x3 level."""
    r = np.random.default_rng(7)
    n = 100_000
    a = ((r.random(n, dtype=np.float32) + 0.5) * np.exp2(r.integers(-3, 3, n))).astype(np.float32)
    b = ((r.random(n, dtype=np.float32) + 0.5) * np.exp2(r.integers(-3, 3, n))).astype(np.float32)
    exact = a.astype(np.float64) * b.astype(np.float64)
    got = np.array([fp16x3_model(a[i:i + 1, None], b[None, i:i + 1])[0, 0] for i in range(0, n, 50)])
    err = np.abs(got - exact[::50]) / exact[::50]
    assert err.max() <= 2.0 ** -20


@pytest.mark.parametrize("k", [8, 256])
def test_fp16x3_random_and_wide_dynamic_range(k):
    r = np.random.default_rng(k)
    a, b = r.random((64, k), dtype=np.float32), r.random((k, 48), dtype=np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    assert (np.abs(fp16x3_model(a, b) - exact) / exact).max() <= 2.0 ** -20
    # rows of A / columns of B scaled by 2^-60 .. 2^60: the per-row / per-column exponents absorb it exactly
    a2 = (a * np.exp2(r.integers(-60, 61, size=(64, 1))).astype(np.float32)).astype(np.float32)
    b2 = (b * np.exp2(r.integers(-60, 61, size=(1, 48))).astype(np.float32)).astype(np.float32)
    exact2 = a2.astype(np.float64) @ b2.astype(np.float64)
    assert (np.abs(fp16x3_model(a2, b2) - exact2) / np.abs(exact2)).max() <= 2.0 ** -20


def test_fp16x3_eligibility_window_and_gather():
    """Inside the window (every element within 2^-28 of its row / column maximum) each element keeps a 2^-22 relative split
    error, so even a gather through a permutation matrix is accurate per element; outside it the model (like the kernel)
    hands over to TF32x3."""
    r = np.random.default_rng(3)
    a = ((r.random((32, 128), dtype=np.float32) + 0.5) * np.exp2(r.integers(-26, 1, size=(32, 128))).astype(np.float32)).astype(np.float32)
    perm = r.permutation(128)
    pm = np.zeros((128, 128), np.float32)
    pm[perm, np.arange(128)] = 1.0
    got = fp16x3_model(a, pm)
    assert got is not None
    assert (np.abs(got - a[:, perm].astype(np.float64)) / a[:, perm]).max() <= 2.0 ** -21
    a[3, 4] = a[3].max() * np.float32(2.0 ** -40)
    assert fp16x3_model(a, pm) is None


def fp16x3_model_with_fixup(a, b):
    """Out-of-window elements (fp16_fixup_kernel in sgemm_tcgen05.cu): the call stays on the FP16 path
    and the few elements the half parts cannot represent are repaired with a sparse rank-1 update.  For such an element of A the split leaves a
    residual d = a_ik - (hi + lo * 2^-11) / 2^e_i, and C[i, :] += d * B[k, :] (fp32 axpy) restores its full contribution;
    symmetrically C[:, j] += A[:, k] * d for an element of B.  (The kernel takes the element out of the GEMM entirely,
    hi = lo = 0 and d = a: a lo-only remnant would meet only the partner's 11-bit hi part.)  Returns (C, repaired)."""
    d64 = np.float64
    ea = scale_exp(np.abs(a).max(axis=1))[:, None]
    eb = scale_exp(np.abs(b).max(axis=0))[None, :]
    xa, xb = np.ldexp(a.astype(d64), ea), np.ldexp(b.astype(d64), eb)
    out_a = (xa != 0) & (np.abs(xa) < 2.0 ** -14)                # the window test of split_f16()
    out_b = (xb != 0) & (np.abs(xb) < 2.0 ** -14)
    ah, al, _ = split_f16_scaled(np.where(out_a, np.float32(0), a), ea)
    bh, bl, _ = split_f16_scaled(np.where(out_b, np.float32(0), b), eb)
    c = np.ldexp(ah @ bh + ah @ bl + al @ bh, -(ea + eb))
    ra = np.where(out_a, a.astype(d64), 0.0)                     # the records carry the whole element
    rb = np.where(out_b, b.astype(d64), 0.0)
    for i, k in zip(*np.nonzero(out_a)):
        c[i, :] += ra[i, k] * b[k, :].astype(d64)
    for k, j in zip(*np.nonzero(out_b)):
        c[:, j] += a[:, k].astype(d64) * rb[k, j]
    return c, int(out_a.sum() + out_b.sum())


def test_fp16x3_sparse_fixup_model_restores_out_of_window_elements():
    r = np.random.default_rng(11)
    a = (r.random((48, 96), dtype=np.float32) + 0.5).astype(np.float32)
    b = (r.random((96, 40), dtype=np.float32) + 0.5).astype(np.float32)
    # out-of-window elements whose products dominate an output: the partner column is zero everywhere else
    a[5, 7] = np.float32(2.0 ** -40)
    b[:, 3] = 0.0
    b[7, 3] = 1.0                                                # C[5, 3] == a[5, 7] exactly
    b[20, 9] = np.float32(3.0 * 2.0 ** -45)                      # an out-of-window element of B
    exact = a.astype(np.float64) @ b.astype(np.float64)
    assert fp16x3_model(a, b) is None                            # without the repair: ineligible (the gated TF32x3 fallback)
    c, repaired = fp16x3_model_with_fixup(a, b)
    assert repaired == 2
    assert (np.abs(c - exact) / np.abs(exact)).max() <= 2.0 ** -20


# ------------------------------------------------------------------ FP16x3, merged 256x256 tile: scale-input-d accumulation
def fp16x3_merged_model(a, b, kc=64):
    """GemmCfg<SCALED, MERGED> (sgemm_tcgen05.cu): the lo parts are stored times 2^11, so the two cross products of a k-block
    come out 2^11 too large; they are accumulated FIRST (fp32, in TMEM) and the first a_hi.b_hi MMA of the block carries
    tcgen05.mma's scale-input-d = 11 (D = A.B + D * 2^-11, exact power of two).  Blocks of kc = 64 along K are then added
    round-to-nearest in fp32 registers.  Returns None when an element is outside the window (repair / fallback path)."""
    ea = scale_exp(np.abs(a).max(axis=1))[:, None]
    eb = scale_exp(np.abs(b).max(axis=0))[None, :]
    ah, al, ok_a = split_f16_scaled(a, ea)       # al, bl returned with the 2^-11 already applied
    bh, bl, ok_b = split_f16_scaled(b, eb)
    if not (ok_a and ok_b):
        return None
    tot = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for k0 in range(0, a.shape[1], kc):
        sl = slice(k0, k0 + kc)
        cross = ((al[:, sl] * 2048.0) @ bh[sl] + ah[:, sl] @ (bl[sl] * 2048.0)).astype(np.float32)    # what TMEM holds, x 2^11
        d = (cross.astype(np.float64) / 2048.0).astype(np.float32)                                    # scale-input-d: exact
        d = (d.astype(np.float64) + ah[:, sl] @ bh[sl]).astype(np.float32)
        tot = (tot + d).astype(np.float32)                                                            # FADD.RN in the epilogue warps
    return np.ldexp(tot.astype(np.float64), -(ea + eb))


@pytest.mark.parametrize("k", [8, 64, 200, 1024])
def test_fp16x3_merged_scale_input_d_matches_the_split_model(k):
    """The merged tile computes the same three products as the 256x128 tile (fp16x3_model), only the accumulation order differs:
    both stay within the 3 * 2^-22 split bound plus fp32 accumulation rounding."""
    r = np.random.default_rng(300 + k)
    a, b = r.random((48, k), dtype=np.float32), r.random((k, 40), dtype=np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    got = fp16x3_merged_model(a, b)
    assert (np.abs(got - exact) / exact).max() <= 3 * 2.0 ** -22 + (k / 64 + 2) * 2.0 ** -24
    assert (np.abs(got - fp16x3_model(a, b)) / exact).max() <= (k / 64 + 2) * 2.0 ** -24
    # coherent inputs: the bound is per product, so constants do not accumulate an error beyond it
    a[:] = 1.0004883
    b[:] = 0.7501221
    exact = a.astype(np.float64) @ b.astype(np.float64)
    assert (np.abs(fp16x3_merged_model(a, b) - exact) / exact).max() <= 3 * 2.0 ** -22 + (k / 64 + 2) * 2.0 ** -24


# ------------------------------------------------------------------ FP16x3U: half hi parts + UNSCALED half lo parts
def split_f16u_scaled(x, e):
    """split_f16<MIX = true>() in sgemm_tcgen05.cu: hi = rn_f16(x * 2^e), lo = rn_f16(x * 2^e - hi) with NO 2^11 factor (a lo part
    below 2^-14 is a subnormal half); the window closes at 2^-6 instead of 2^-14."""
    xs = np.ldexp(x.astype(np.float64), e)
    hi = xs.astype(np.float32).astype(np.float16)
    lo = (xs - hi.astype(np.float64)).astype(np.float32).astype(np.float16)      # xs - hi is exact in fp32
    eligible = bool(((xs == 0) | (np.abs(xs) >= 2.0 ** -6)).all())
    return hi.astype(np.float64), lo.astype(np.float64), eligible


def fp16x3u_model(a, b):
    ea = scale_exp(np.abs(a).max(axis=1))[:, None]
    eb = scale_exp(np.abs(b).max(axis=0))[None, :]
    ah, al, ok_a = split_f16u_scaled(a, ea)
    bh, bl, ok_b = split_f16u_scaled(b, eb)
    if not (ok_a and ok_b):
        return None
    return np.ldexp(ah @ bh + ah @ bl + al @ bh, -(ea + eb))


def test_fp16u_split_remainder_bound():
    """remainder <= max(2^-22 |x'|, 2^-25): 2^-22 relative down to 2^-3 (2^-17 of the maximum's binade), 2^-19 at the window edge 2^-6."""
    r = np.random.default_rng(21)
    x = ((r.random(1 << 16, dtype=np.float32) + 0.5) * np.exp2(r.integers(-19, 1, 1 << 16)).astype(np.float32)).astype(np.float32)
    x[0] = 1.4999999                                                  # the row maximum: the exponent brings it to [2^14, 2^15)
    e = scale_exp(np.abs(x).max(keepdims=True))[:, None]
    hi, lo, ok = split_f16u_scaled(x[None, :], e)
    assert ok
    xs = np.ldexp(x.astype(np.float64), int(e[0, 0]))
    rem = np.abs(xs - hi[0] - lo[0])
    assert (rem <= np.maximum(np.abs(xs) * 2.0 ** -22, 2.0 ** -25)).all()
    assert (rem <= np.abs(xs) * 2.0 ** -19).all()
    big = np.abs(xs) >= 2.0 ** -3
    assert (rem[big] <= np.abs(xs[big]) * 2.0 ** -22).all()


def test_fp16x3u_coherent_inputs_stay_inside_1e5():
    """One product repeated K times (what breaks BF16x3).  Constant matrices put every element AT its row / column maximum, where the
    split error is 2^-22: the same class as FP16x3."""
    r = np.random.default_rng(7)
    n = 100_000
    a = ((r.random(n, dtype=np.float32) + 0.5) * np.exp2(r.integers(-3, 3, n))).astype(np.float32)
    b = ((r.random(n, dtype=np.float32) + 0.5) * np.exp2(r.integers(-3, 3, n))).astype(np.float32)
    exact = a.astype(np.float64) * b.astype(np.float64)
    got = np.array([fp16x3u_model(a[i:i + 1, None], b[None, i:i + 1])[0, 0] for i in range(0, n, 20)])
    err = np.abs(got - exact[::20]) / exact[::20]
    assert err.max() <= 2.0 ** -20


def test_fp16x3u_worst_case_pair_at_the_window_edge():
    """Both factors 2^-20 below their row / column maximum: per-product error <= 2^-18 + 2^-22 = 4.1e-6 < 1e-5."""
    r = np.random.default_rng(9)
    k = 512
    a = np.zeros((k, 2), np.float32)
    b = np.zeros((2, k), np.float32)
    a[:, 0] = 1.5; b[0, :] = 1.5                                      # the maxima (they meet only each other)
    a[:, 1] = ((r.random(k) + 1.0) * 2.0 ** -20).astype(np.float32)  # x' in [2^-6, 2^-5): window edge
    b[1, :] = ((r.random(k) + 1.0) * 2.0 ** -20).astype(np.float32)
    a[:, 0] = 0.0                                                     # ... and now only the small elements contribute
    a[0, 0] = 1.5
    got = fp16x3u_model(a, b)
    assert got is not None
    exact = a.astype(np.float64) @ b.astype(np.float64)
    sel = exact[1:, :] > 0
    err = np.abs(got[1:, :] - exact[1:, :])[sel] / exact[1:, :][sel]
    assert err.max() <= 2.0 ** -18 + 2.0 ** -22 < 1e-5


@pytest.mark.parametrize("k", [8, 256, 2048])
def test_fp16x3u_random_dynamic_range_and_gather(k):
    r = np.random.default_rng(300 + k)
    a, b = r.random((64, k), dtype=np.float32), r.random((k, 48), dtype=np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    got = fp16x3u_model(a, b)
    if got is not None:                                               # (U[0,1) data: an element below 2^-21 of its maximum is possible)
        err = np.abs(got - exact) / exact
        assert err.max() <= 2.0 ** -18
        if k >= 256:
            assert err.max() <= 5e-7
    a2 = ((a + 0.01) * np.exp2(r.integers(-60, 61, size=(64, 1))).astype(np.float32)).astype(np.float32)
    b2 = ((b + 0.01) * np.exp2(r.integers(-60, 61, size=(1, 48))).astype(np.float32)).astype(np.float32)
    exact2 = a2.astype(np.float64) @ b2.astype(np.float64)
    assert (np.abs(fp16x3u_model(a2, b2) - exact2) / np.abs(exact2)).max() <= 2.0 ** -18
    # gather: every output is ONE input element with its own split error (<= 2^-19 inside the window, here down to 2^-19 of the maximum)
    g = ((r.random((32, 128), dtype=np.float32) + 0.5) * np.exp2(r.integers(-19, 1, size=(32, 128))).astype(np.float32)).astype(np.float32)
    g[:, 0] = 1.4
    perm = r.permutation(128)
    pm = np.zeros((128, 128), np.float32)
    pm[perm, np.arange(128)] = 1.0
    got = fp16x3u_model(g, pm)
    assert got is not None
    assert (np.abs(got - g[:, perm].astype(np.float64)) / g[:, perm]).max() <= 2.0 ** -19
    g[3, 4] = g[3].max() * np.float32(2.0 ** -23)                     # outside the (narrower) window: repair / fallback path
    assert fp16x3u_model(g, pm) is None
