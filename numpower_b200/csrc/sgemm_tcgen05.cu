// nd::matmul on the Blackwell tensor pipe: C[M,N] = A[M,K] . B[K,N], row-major fp32 in/out,
// TF32 multiplies on tcgen05 with fp32 accumulation in TMEM.  Replaces the cublasSgemm call
// of NDArray_FMatmul (src/ndmath/linalg.c:54-72 in /root/reference); the CPU oracle is
// cblas_sgemm (linalg.c:75-79).
//
// Structure (one persistent CTA — or CTA pair — per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) global -> smem ring, mbarrier tx
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (smem descriptors),
//               tcgen05.commit releases smem slots / publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld TMEM -> registers -> st.global.v4, overlapped with the next
//               tile's mainloop through a 2-deep TMEM accumulator ring
//
// Operand layouts in shared memory (per pipeline stage, BK = 32 fp32 = one 128-byte swizzle row):
//   A tile  (BM x BK)  K-major : rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO = 1024)
//   B tile  (BK x BN)  MN-major: B is row-major [K,N], i.e. N is contiguous -> loaded untransposed
//           as BN/32 chunks of [32 k-rows x 128 B]; chunk stride = LBO = 4096 B.  For 4-byte MN-major
//           operands the tensor core only accepts the "128B swizzle with 32-byte atoms" layout
//           (UMMA layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: 32-byte chunk index ^= row % 4),
//           whose atom is 4 k-rows x 128 B: SBO = 512 B.  One K=8 MMA consumes two such atoms of every
//           chunk (start address advances 1024 B per k-step).  [Measured on B200: the 16-byte-atom
//           SWIZZLE_128B layout silently yields an all-zero accumulator for MN-major tf32.]
//
// Precision modes (include/nb200.h):
//   TF32X1  operands are fed as raw fp32 (the tensor core reads the top 19 bits)
//   TF32X3  error-compensated: a = a_hi + a_lo with a_hi = trunc_tf32(a) (what the tensor core reads from the raw
//           array) and a_lo = rna_tf32(a - a_hi) from the pre-pass; C = a_lo.b_hi + a_hi.b_lo + a_hi.b_hi —
//           three MMAs per k-step, chunked round-to-nearest accumulation (GemmCfg::CHUNKED), fp32-class accuracy
//
// Flops: 2*M*N*K useful (x3 executes 6*M*N*K on the tensor pipe).
#include "common.cuh"
#include <cuda.h>
#include <cstring>
#include <math_constants.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace nb200 {

// ------------------------------------------------------------------ configuration
template <int CG_, int BN_, int PASSES_, bool INK_ = false, bool BF16_ = false, bool MERGED_ = false, bool SCALED_ = false, bool MIXLO_ = false>
struct GemmCfg {
    // MIXLO (FP16x3U = SCALED with UNSCALED lo parts): hi = rn_f16(a') as in FP16x3, lo = rn_f16(a' - hi) without the 2^11 factor, so
    // all three products of a k-block carry the same scale and accumulate into ONE accumulator per chunk => the MERGED 256x256 tile
    // works exactly as for BF16x3 (K = 256 chunks, no scale-input-d).  Price: a lo part below 2^-14 is a subnormal half (absolute
    // error 2^-25), so the split error per element is max(2^-22 |a'|, 2^-25): 2^-22 relative for elements within 2^-17 of the row /
    // column maximum, growing to 2^-19 at |a'| = 2^-6; elements below 2^-6 (2^-20 .. 2^-21 of the maximum instead of FP16x3's 2^-28)
    // take the sparse-repair path.
    // [Measured on B200: tcgen05.mma.kind::f16 with DIFFERENT A and B formats (f16 x bf16) raises "illegal instruction", although the
    //  instruction descriptor has separate format fields - a bfloat16 lo part next to a half hi part is therefore not an option.]
    static constexpr bool MIXLO = MIXLO_;
    static_assert(!MIXLO_ || SCALED_, "unscaled lo parts belong to the scaled half-precision mode");
    // SCALED (FP16x3): the 16-bit operand parts are IEEE half (11-bit significands, TF32x3-class split error) of
    // A scaled per row and B scaled per column by powers of two (so that every row / column uses the top of fp16's
    // exponent range); the epilogue undoes the scaling with one exact scalbnf per output element.
    static constexpr bool SCALED = SCALED_;
    static_assert(!SCALED_ || BF16_, "scaled mode: 16-bit operands");
    // SCALED stores the lo parts as IEEE half times 2^11 (so that they use the same fp16 binades as the hi parts): the two
    // cross products come out 2^11 too large.
    //  * separate cross-term accumulator (256x128 tiles): scaled back (exactly) when it is folded into the total;
    //  * SCALED + MERGED (256x256 tiles, ONE accumulator per chunk): a chunk is a single k-block (K = 64); its cross
    //    products are issued first and the first a_hi.b_hi MMA of the chunk carries tcgen05.mma's scale-input-d = 11
    //    (D = A.B + D * 2^-11, exact), so the chunk accumulator ends up holding hh + 2^-11 (lh + hl).
    static constexpr float CROSS_SCALE = (SCALED_ && !MERGED_ && !MIXLO_) ? 1.0f / 2048.0f : 1.0f;
    static constexpr bool SCALE_D = SCALED_ && MERGED_ && !MIXLO_;   // fold 2^11-scaled cross products with scale-input-d, one k-block per chunk
    // MERGED (BF16x3, BN = 256): all three products of a chunk accumulate into ONE 256-column TMEM accumulator (2-deep
    // ring = all 512 columns) and the running total lives in the registers of eight epilogue warps.  Twice the flops
    // per staged byte and no cross-accumulator hand-off between tiles; costs 48 instead of 32 truncating accumulation
    // steps per (256-long) chunk.
    static constexpr bool MERGED = MERGED_;
    // BF16 (x3 only): operands are pre-split into two bfloat16 arrays each and multiplied with kind::f16 MMAs at twice
    // the TF32 rate; everything else (ring, chunked accumulation, epilogue) is shared with TF32x3.
    static constexpr bool BF16 = BF16_;
    static constexpr int ESZ = BF16_ ? 2 : 4;            // operand element size in shared memory
    // INK (TF32x3 only): the lo parts are computed INSIDE the kernel by four converter warps (smem raw tile ->
    // a_lo / b_lo smem tiles) instead of by the split pre-pass: TMA moves only the raw operands (half the L2->SM
    // bytes) and the 8 B/element pre-pass traffic disappears.
    static constexpr bool INK = INK_;
    static constexpr int CG = CG_;                       // CTAs cooperating on one MMA (cta_group)
    static constexpr int BN = BN_;                       // tile columns (per CTA pair when CG == 2)
    static constexpr int PASSES = PASSES_;               // 1 = TF32x1, 3 = TF32x3
    static constexpr int BM = 128;                       // rows per CTA (TMEM lanes)
    static constexpr int BK = 128 / ESZ;                 // elements per stage along K = one 128-byte swizzle row
    static constexpr int UMMA_K = 32 / ESZ;              // 32 B of K per instruction (tf32: 8, bf16: 16)
    // MN-major B tile = BN_CTA/B_CHUNK_N chunks of [BK k-rows x 128 B]
    static constexpr int B_CHUNK_N = 128 / ESZ;          // n-columns per chunk
    static constexpr int B_CHUNK_BYTES = BK * 128;       // = LBO
    static constexpr int B_SBO = BF16_ ? 1024 : 512;     // 16-bit: 8-row swizzle atoms; 32-bit: 4-row atoms (32-byte-atom layout)
    static constexpr int B_KSTEP = UMMA_K * 128;         // start-address advance per MMA k-step
    static constexpr int BN_CTA = BN / CG;               // B columns staged by each CTA
    static constexpr int NPART = PASSES == 3 ? 2 : 1;    // hi (+ lo)
    static constexpr int A_BYTES = BM * 128;             // 16 KiB
    static constexpr int B_BYTES = BN_CTA * BK * ESZ;
    static constexpr int STAGE_BYTES = NPART * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
    static constexpr int ACC_STAGES = 2;
    // TF32x3 keeps fp32-class accuracy only if the tensor core's truncating (round-toward-zero) fp32
    // accumulation is kept short: measured on B200 the bias grows ~ -1e-8 * K (4.3e-5 at K = 4096).
    // CHUNKED mode therefore (a) accumulates the leading product a_hi.b_hi in chunks of KC along K into a
    // 2-deep TMEM ring that the epilogue warps drain into a running total with round-to-nearest adds
    // (tcgen05.ld / FADD / tcgen05.st), and (b) keeps the 2^-11-smaller cross terms in their own
    // accumulator.  TMEM columns: [0,BN) [BN,2BN) main ring | [2BN,3BN) cross terms | [3BN,4BN) running total.
    static constexpr bool CHUNKED = PASSES == 3;
    // K elements per accumulation chunk (32 MMA k-steps; MERGED: 16 x 3 products; SCALED + MERGED: one k-block, see above)
    static constexpr int KC = SCALE_D ? BK : (BF16_ && !MERGED_) ? 512 : 256;
    static constexpr int KB_PER_CHUNK = KC / BK;
    static constexpr int TMEM_COLS = (CHUNKED && !MERGED_) ? 4 * BN : ACC_STAGES * BN;   // 256 or 512 (power of two)
    static_assert(!CHUNKED || MERGED_ || BN == 128, "chunked x3 uses 128-column tiles (4 x 128 TMEM columns)");
    static_assert(!MERGED_ || (BF16_ && BN == 256 && PASSES_ == 3), "merged accumulation exists for BF16x3 with 256-column tiles");
    static constexpr int EPI_WARPS = MERGED_ ? 8 : 4;
    // MERGED epilogue: accumulator columns per tcgen05.ld (each load is followed by a wait of a few hundred cycles).  One k-block
    // per chunk (SCALED) drains a 256-column accumulator every 12 MMAs: 64 columns per load, two waits per chunk.
    static constexpr int EPI_LD = SCALE_D ? 64 : 16;
    // MERGED: 12 warps = 3 warpgroups.  Warpgroup 0 = TMA producer, MMA issuer and two idle warps (TMEM allocation); warpgroups
    // 1-2 = eight epilogue warps that keep 128 running totals + the TMEM load in flight in registers: setmaxnreg moves registers
    // from warpgroup 0 (72 per thread) to the epilogue warpgroups (216 per thread; 128 * 72 + 256 * 216 <= 64 Ki).  A 10-warp CTA
    // would be allocated as 12 warps anyway (168 registers per thread, spills in the epilogue).
    static constexpr int THREADS = MERGED_ ? 384 : (INK_ ? 320 : 192);     // INK: + 4 converter warps
    static constexpr int EPI_WARP0 = MERGED_ ? 4 : 2;                     // first epilogue warp
    static constexpr int TMA_BYTES = INK_ ? (A_BYTES + B_BYTES) : STAGE_BYTES;   // bytes the TMA lands per stage per CTA
    static_assert(!INK_ || (PASSES_ == 3 && !BF16_), "in-kernel split only exists for TF32x3");
    static_assert(!BF16_ || PASSES_ == 3, "bf16 operands are only used by the x3 error-compensated mode");
    // SCALED + MERGED: per-tile column scale factors (two floats per column, double-buffered by tile parity) behind the barrier block
    static constexpr int COLFAC_BYTES = (SCALED_ && MERGED_) ? 2 * 256 * 8 : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 384 /*barriers*/ + COLFAC_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
};

struct GemmParams {
    float *C;
    int64_t M, N, K, ldc, strideC;
    int64_t batch;
    int tiles_m, tiles_n;      // tiles per matrix
    int group_m;               // tile raster: row-tiles per band (0 = plain m-fastest order)
    int64_t full_items;        // MERGED: work items >= full_items are half tiles (128 of the 256 columns), two per tile
    int64_t total_tiles;
    int a_batched, b_batched;  // 0: operand shared across the batch (coordinate 0)
    int early_cross;           // chunked epilogue: release the cross accumulator before writing C (A/B switch)
    // Device-side gate: the kernel returns at once unless (*gate == gate_gen) == (gate_want != 0).  Lets the host enqueue both
    // the FP16x3 product and its TF32x3 fallback and have the split pre-pass decide, without a host round trip.
    const int *gate;
    int gate_gen, gate_want;
    const unsigned int *row_max, *col_max;   // SCALED: |max| bit patterns per row of A / column of B (see scale_exp)
    const int *nonfinite;      // x3 modes: the split pre-pass stores nonfinite_gen here when an operand holds +-inf (see lo_part)
    int nonfinite_gen;         // this call's tag (a fresh value per call instead of a memset per call)
    unsigned int *debug;       // [0] = timeout flag, [1..] = info
    unsigned long long *trace; // optional timeline stamps (common.cuh), nullptr normally
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// shared::cluster address of the same smem offset in CTA rank 0 of the cluster (the MMA leader)
__device__ __forceinline__ uint32_t map_to_leader(uint32_t addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(0));
    return r;
}
// arrive on the barrier at the same offset in the leader CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_leader(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int *debug, uint32_t tag) {
    for (uint32_t spin = 0; spin < (1u << 26); spin++)
        if (mbar_try_wait(bar, parity)) return;
    if (debug) {
        debug[0] = 1u;
        debug[1] = tag;
        debug[2] = blockIdx.x;
        debug[3] = parity;
    }
    __threadfence_system();
    asm volatile("trap;");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still running (once every CTA of the predecessor has called launch_dependents or exited);
// griddepcontrol.wait blocks until the predecessor has completed and its writes are visible.  Every kernel of a chain executes
// the wait before it exits, so completion stays ordered along the stream.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    if constexpr (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    } else {
        // both CTAs of the pair post their bytes on the LEADER's barrier
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(dst), "l"(map), "r"(map_to_leader(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc]
template <int CG, bool BF16 = false>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (BF16 && CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else if constexpr (BF16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// kind::f16 with scale-input-d: D = A.B + D * 2^-11 (SCALED + MERGED: folds the 2^11-scaled cross products of a chunk)
template <int CG>
__device__ __forceinline__ void umma_f16_scale_d11(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
    }
}
// all previously issued MMAs complete -> one arrival on `bar` (in both CTAs of a pair when CG == 2)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"((uint16_t)0x3) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (64-bit), SM100 format:
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1       | [49,52) base offset = 0          | [61,64) layout: 2 = SWIZZLE_128B,
//                                                                           1 = SWIZZLE_128B with 32-byte atoms
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// Instruction descriptor (32-bit): D=f32 (bits 4-5 =1), A=B=tf32 (bits 7-9, 10-12 =2), A K-major (bit 15 =0),
// B MN-major (bit 16 =1), N>>3 at bits 17-22, M>>4 at bits 24-28.
// kind::f16 uses the same fields with A=B=bf16 (format code 1).
// fmt: 0 = f16, 1 = bf16 (both kind::f16), 2 = tf32.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt_a, uint32_t fmt_b) {
    return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// FP16x3 scaling: the power of two that brings a row's / column's largest finite magnitude (given as its bit pattern,
// collected by absmax_*_kernel) into [2^14, 2^15), the top binade below fp16's 65504.  0 for empty / all-zero vectors.
// The exponent lies in [-113, 100]: scaling UP is capped (vectors whose largest magnitude is below 2^-86 are simply not brought
// all the way up) so that a row and a column exponent always combine into two representable power-of-two factors.
__device__ __forceinline__ int scale_exp(unsigned int max_bits) {
    if (max_bits == 0u || max_bits >= 0x7F800000u) return 0;
    const int e = 14 - ((int)(max_bits >> 23) - 127);          // >= -113; subnormal maxima read as 2^-127 and hit the cap
    return e > 100 ? 100 : e;
}
// x * 2^e, -226 <= e <= 226, as two exact multiplications (2^e itself may not be representable)
__device__ __forceinline__ float scale_pow2(float x, int e) {
    const int e1 = e / 2, e2 = e - e1;
    return x * __int_as_float((e1 + 127) << 23) * __int_as_float((e2 + 127) << 23);
}

// Tile order within one matrix: bands of `group_m` row-tiles, inside a band m fastest.  The tiles in flight at any time
// (one per CTA pair) then cover ~group_m x (clusters / group_m) tiles: every A row-block and B column-block they touch is
// shared by several of them while it is still in L2, and consecutive waves of a band reuse the band's A rows.
__device__ __forceinline__ void tile_coords(int64_t r, const GemmParams &p, int &m_tile, int &n_tile) {
    const int gm = p.group_m;
    if (gm <= 0 || gm >= p.tiles_m) { m_tile = (int)(r % p.tiles_m); n_tile = (int)(r / p.tiles_m); return; }
    const int64_t per_band = (int64_t)gm * p.tiles_n;
    const int band = (int)(r / per_band);
    const int within = (int)(r - band * per_band);
    const int rows = (p.tiles_m - band * gm) < gm ? (p.tiles_m - band * gm) : gm;   // last band may be short
    m_tile = band * gm + within % rows;
    n_tile = within / rows;
}

// ------------------------------------------------------------------ the kernel
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
sgemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const GemmParams p) {
    constexpr int CG = Cfg::CG, BN = Cfg::BN, BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
    constexpr int PASSES = Cfg::PASSES, NPART = Cfg::NPART;
    // Gated fallback (gate_want = 1: run only if the pre-pass marked the call): normally not needed, leave before any set-up.
    const int tslot = (p.gate != nullptr && p.gate_want != 0) ? 10 : 5;
    if (threadIdx.x == 0) trace_min(p.trace, tslot);
    if (p.gate != nullptr && p.gate_want != 0) {
        pdl_wait();
        if (*reinterpret_cast<const volatile int *>(p.gate) != p.gate_gen) {   // uniform over the grid
            if (threadIdx.x == 0) trace_max(p.trace, 11);
            return;
        }
    }
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    // barrier block: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t cross_empty_bar = bar_base + 8u * (2 * STAGES + 4);
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 5);
    auto conv_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 6 + s); };   // INK: lo parts of stage s ready (leader's copy is the one waited on)
    // stage layout: [A_hi][A_lo?][B_hi][B_lo?]
    auto a_smem = [&](int s, int part) { return smem_base + s * Cfg::STAGE_BYTES + part * Cfg::A_BYTES; };
    auto b_smem = [&](int s, int part) { return smem_base + s * Cfg::STAGE_BYTES + NPART * Cfg::A_BYTES + part * Cfg::B_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = (rank == 0);
    const int64_t cluster_id = blockIdx.x / CG;
    const int64_t num_clusters = gridDim.x / CG;
    const int num_kb = (int)((p.K + BK - 1) / BK);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmB_hi);
        if (PASSES == 3 && !Cfg::INK) { tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), Cfg::EPI_WARPS * CG); }
        mbar_init(cross_empty_bar, 4 * CG);
        if (Cfg::INK) for (int s = 0; s < STAGES; s++) mbar_init(conv_bar(s), 4 * CG);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<CG>(tmem_ptr_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

    // Everything above touched only this CTA's shared / tensor memory and overlapped the tail of the previous kernel in the stream
    // (the split pre-pass); the operands, the gate and the non-finite flag are read below.
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0 && tslot == 5) trace_min(p.trace, 6);
    // gate_want = 0 (FP16x3 product): skip the work if the pre-pass marked the call ineligible (the gated fallback runs instead)
    const bool skip = p.gate != nullptr && p.gate_want == 0 && *reinterpret_cast<const volatile int *>(p.gate) == p.gate_gen;   // uniform over the grid
    const int64_t total_tiles = skip ? 0 : p.total_tiles;
    const int64_t tiles_per_mat = (int64_t)p.tiles_m * p.tiles_n;
    auto run_epilogue = [&]() {
        // ===================== epilogue (warps 2..5; MERGED: 4..11, warps 8..11 take the upper 128 columns) =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        int epi_tile = 0;   // tiles this CTA has started (parity selects the column-factor buffer, SCALED + MERGED)
        (void)epi_tile;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.strideC & 3) == 0);
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        for (int64_t t = cluster_id; t < total_tiles; t += num_clusters) {
            const bool half_tile = Cfg::MERGED && t >= p.full_items;
            const int64_t tt = half_tile ? p.full_items + ((t - p.full_items) >> 1) : t;
            const int64_t b = tt / tiles_per_mat, r = tt % tiles_per_mat;
            int m_tile, n_tile;
            tile_coords(r, p, m_tile, n_tile);
            const int64_t row = (int64_t)m_tile * BM * CG + (int64_t)rank * BM + quarter * 32 + lane;
            const int64_t col0 = (int64_t)n_tile * BN + (half_tile ? (int64_t)((t - p.full_items) & 1) * (BN / 2) : 0);
            float *crow = p.C + b * p.strideC + row * p.ldc;
            // SCALED: C = acc * 2^-(row exponent + column exponent), one exact scalbnf per element
            int e_row = 0;
            const unsigned int *cmax = nullptr;
            if constexpr (Cfg::SCALED) {
                if (row < p.M) e_row = scale_exp(p.row_max[(p.a_batched ? b : 0) * p.M + row]);
                cmax = p.col_max + (p.b_batched ? b : 0) * p.N;
            }
            auto store_row = [&](int c, const uint32_t (&v)[32]) {
                const int64_t col = col0 + c * 32;
                if (row < p.M) {
                    if constexpr (Cfg::SCALED) {
                        if (vec_ok && col + 32 <= p.N && (p.N & 3) == 0) {   // column maxima and C rows 16-byte aligned
                            const uint4 *mb = reinterpret_cast<const uint4 *>(cmax + col);
                            float4 *dst = reinterpret_cast<float4 *>(crow + col);
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                const uint4 m4 = __ldg(mb + q);
                                dst[q] = make_float4(scale_pow2(__uint_as_float(v[4 * q]), -(e_row + scale_exp(m4.x))),
                                                     scale_pow2(__uint_as_float(v[4 * q + 1]), -(e_row + scale_exp(m4.y))),
                                                     scale_pow2(__uint_as_float(v[4 * q + 2]), -(e_row + scale_exp(m4.z))),
                                                     scale_pow2(__uint_as_float(v[4 * q + 3]), -(e_row + scale_exp(m4.w))));
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; q++)
                                if (col + q < p.N) crow[col + q] = scale_pow2(__uint_as_float(v[q]), -(e_row + scale_exp(cmax[col + q])));
                        }
                    } else if (vec_ok && col + 32 <= p.N) {
                        float4 *dst = reinterpret_cast<float4 *>(crow + col);
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                 __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; q++)
                            if (col + q < p.N) crow[col + q] = __uint_as_float(v[q]);
                    }
                }
            };
            if constexpr (Cfg::MERGED) {
                // running total of this thread's row (128 of the tile's 256 columns) in registers, chunks added round-to-nearest
                const int half = (warp - Cfg::EPI_WARP0) >> 2;
                const bool idle = half_tile && half == 1;   // half tiles only have the lower 128 accumulator columns
                float tot[128];
                // SCALED: the 2^-(row exponent + column exponent) of every output as FOUR exact power-of-two factors,
                // x * (c1 * r1) * (c2 * r2) with 2^-ec = c1 * c2 and 2^-er = r1 * r2 split in halves (the intermediate lies between x
                // and the result: no spurious overflow).  The 256 epilogue threads compute the tile's 256 column pairs ONCE, into
                // shared memory (double-buffered by tile parity, one named barrier per tile), instead of every thread redoing the
                // exponent arithmetic for its 128 columns at the end of the tile, when the next tile's first chunks are waiting.
                float r1 = 1.f, r2 = 1.f;
                uint32_t colfac = 0;
                if constexpr (Cfg::SCALED) {
                    const int et = (int)threadIdx.x - Cfg::EPI_WARP0 * 32;   // 0 .. 255
                    colfac = bar_base + 384u + (uint32_t)((epi_tile & 1) * 2048);
                    const int64_t cc = col0 + et;
                    const int ec = cc < p.N ? scale_exp(cmax[cc]) : 0;
                    const int c1e = (-ec) / 2, c2e = -ec - c1e;
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(colfac + (uint32_t)et * 8u),
                                 "f"(__int_as_float((c1e + 127) << 23)), "f"(__int_as_float((c2e + 127) << 23)) : "memory");
                    const int r1e = (-e_row) / 2, r2e = -e_row - r1e;
                    r1 = __int_as_float((r1e + 127) << 23);
                    r2 = __int_as_float((r2e + 127) << 23);
                    asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps (barrier 0 belongs to __syncthreads)
                    epi_tile++;
                }
                for (int kb0 = 0; kb0 < num_kb; kb0 += Cfg::KB_PER_CHUNK) {
                    const bool first = kb0 == 0, last = kb0 + Cfg::KB_PER_CHUNK >= num_kb;
                    mbar_wait(tfull_bar(acc), acc_phase, p.debug, 0x400u + acc);
                    tc_fence_after();
                    const uint32_t t_main = tmem_base + lane_sel + (uint32_t)(acc * BN + half * 128);
                    if (!idle) {
#pragma unroll
                        for (int c = 0; c < 128 / Cfg::EPI_LD; c++) {   // EPI_LD columns per TMEM load (one wait each): 128 totals + EPI_LD in flight
                            uint32_t v[Cfg::EPI_LD];
                            if constexpr (Cfg::EPI_LD == 64) tmem_ld_32x64(t_main + (uint32_t)(c * 64), v);
                            else tmem_ld_32x16(t_main + (uint32_t)(c * 16), v);
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < Cfg::EPI_LD; q++)
                                tot[c * Cfg::EPI_LD + q] = first ? __uint_as_float(v[q]) : __fadd_rn(tot[c * Cfg::EPI_LD + q], __uint_as_float(v[q]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_leader(tempty_bar(acc));
                        else mbar_arrive_local(tempty_bar(acc));
                    }
                    if (last && row < p.M && !idle) {
                        const int64_t colh = col0 + half * 128;
                        if constexpr (Cfg::SCALED) {
                            const uint32_t cf = colfac + (uint32_t)(half * 128) * 8u;
#pragma unroll
                            for (int q = 0; q < 64; q++) {   // two columns (four factors) per 16-byte shared load, same address in every lane
                                float c1a, c2a, c1b, c2b;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c1a), "=f"(c2a), "=f"(c1b), "=f"(c2b) : "r"(cf + (uint32_t)q * 16u));
                                tot[2 * q] = tot[2 * q] * (c1a * r1) * (c2a * r2);
                                tot[2 * q + 1] = tot[2 * q + 1] * (c1b * r1) * (c2b * r2);
                            }
                        }
                        if (vec_ok && colh + 128 <= p.N) {
                            float4 *dst = reinterpret_cast<float4 *>(crow + colh);
#pragma unroll
                            for (int q = 0; q < 32; q++) dst[q] = make_float4(tot[4 * q], tot[4 * q + 1], tot[4 * q + 2], tot[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 128; q++)
                                if (colh + q < p.N) crow[colh + q] = tot[q];
                        }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            } else if constexpr (!Cfg::CHUNKED) {
                mbar_wait(tfull_bar(acc), acc_phase, p.debug, 0x400u + acc);
                tc_fence_after();
                const uint32_t taddr0 = tmem_base + lane_sel + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int c = 0; c < BN / 32; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr0 + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    store_row(c, v);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_leader(tempty_bar(acc));
                    else mbar_arrive_local(tempty_bar(acc));
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            } else {
                const uint32_t t_cross = tmem_base + lane_sel + (uint32_t)(2 * BN);
                const uint32_t t_total = tmem_base + lane_sel + (uint32_t)(3 * BN);
                for (int kb0 = 0; kb0 < num_kb; kb0 += Cfg::KB_PER_CHUNK) {
                    const bool first = kb0 == 0, last = kb0 + Cfg::KB_PER_CHUNK >= num_kb;
                    mbar_wait(tfull_bar(acc), acc_phase, p.debug, 0x400u + acc);
                    tc_fence_after();
                    const uint32_t t_main = tmem_base + lane_sel + (uint32_t)(acc * BN);
                    if (!last) {
                        // total (+)= this chunk, round-to-nearest, all in TMEM
#pragma unroll 1
                        for (int c = 0; c < BN / 32; c++) {
                            uint32_t m[32], x[32];
                            tmem_ld_32x32(t_main + (uint32_t)(c * 32), m);
                            if (!first) tmem_ld_32x32(t_total + (uint32_t)(c * 32), x);
                            tmem_ld_wait();
                            if (!first) {
#pragma unroll
                                for (int q = 0; q < 32; q++) m[q] = __float_as_uint(__fadd_rn(__uint_as_float(x[q]), __uint_as_float(m[q])));
                            }
                            tmem_st_32x32(t_total + (uint32_t)(c * 32), m);
                        }
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_leader(tempty_bar(acc));
                            else mbar_arrive_local(tempty_bar(acc));
                        }
                    } else if (!p.early_cross) {
                        // last chunk, single pass: C = (total + chunk) + cross; the next tile's MMAs wait for all of it
#pragma unroll 1
                        for (int c = 0; c < BN / 32; c++) {
                            uint32_t m[32], x[32];
                            tmem_ld_32x32(t_main + (uint32_t)(c * 32), m);
                            if (!first) tmem_ld_32x32(t_total + (uint32_t)(c * 32), x);
                            tmem_ld_wait();
                            if (!first) {
#pragma unroll
                                for (int q = 0; q < 32; q++) m[q] = __float_as_uint(__fadd_rn(__uint_as_float(x[q]), __uint_as_float(m[q])));
                            }
                            tmem_ld_32x32(t_cross + (uint32_t)(c * 32), x);
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 32; q++) m[q] = __float_as_uint(__fadd_rn(__uint_as_float(m[q]), __uint_as_float(x[q]) * Cfg::CROSS_SCALE));
                            store_row(c, m);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) { mbar_arrive_leader(tempty_bar(acc)); mbar_arrive_leader(cross_empty_bar); }
                            else { mbar_arrive_local(tempty_bar(acc)); mbar_arrive_local(cross_empty_bar); }
                        }
                    } else {
                        // last chunk.  Pass 1 (TMEM only, a few hundred cycles): total (+)= cross terms, then hand the cross
                        // accumulator back so the NEXT tile's MMAs start while this tile is still being written out.
#pragma unroll 1
                        for (int c = 0; c < BN / 32; c++) {
                            uint32_t m[32], x[32];
                            tmem_ld_32x32(t_cross + (uint32_t)(c * 32), m);
                            if (!first) tmem_ld_32x32(t_total + (uint32_t)(c * 32), x);
                            tmem_ld_wait();
                            if constexpr (Cfg::SCALED && !Cfg::MIXLO) {   // the lo parts are stored times 2^11 (exact power-of-two factor)
#pragma unroll
                                for (int q = 0; q < 32; q++) m[q] = __float_as_uint(__uint_as_float(m[q]) * Cfg::CROSS_SCALE);
                            }
                            if (!first) {
#pragma unroll
                                for (int q = 0; q < 32; q++) m[q] = __float_as_uint(__fadd_rn(__uint_as_float(x[q]), __uint_as_float(m[q])));
                            }
                            tmem_st_32x32(t_total + (uint32_t)(c * 32), m);
                        }
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_leader(cross_empty_bar);
                            else mbar_arrive_local(cross_empty_bar);
                        }
                        // Pass 2: C = total + last chunk -> global memory
#pragma unroll 1
                        for (int c = 0; c < BN / 32; c++) {
                            uint32_t m[32], x[32];
                            tmem_ld_32x32(t_main + (uint32_t)(c * 32), m);
                            tmem_ld_32x32(t_total + (uint32_t)(c * 32), x);
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 32; q++) m[q] = __float_as_uint(__fadd_rn(__uint_as_float(x[q]), __uint_as_float(m[q])));
                            store_row(c, m);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_leader(tempty_bar(acc));
                            else mbar_arrive_local(tempty_bar(acc));
                        }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    };

    // MERGED: warpgroup-uniform branch so that each side is dominated by its setmaxnreg (register re-balancing, see GemmCfg::THREADS)
    if (Cfg::MERGED && warp >= 4) {
        if constexpr (Cfg::MERGED) {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
            run_epilogue();
        }
    } else {
    if constexpr (Cfg::MERGED) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = cluster_id; t < total_tiles; t += num_clusters) {
                // work item -> tile (+ which half of it when the last wave is split, see launch_gemm)
                const bool half_tile = Cfg::MERGED && t >= p.full_items;
                const int64_t tt = half_tile ? p.full_items + ((t - p.full_items) >> 1) : t;
                const int ncols_cta = half_tile ? Cfg::BN_CTA / 2 : Cfg::BN_CTA;   // B columns this CTA stages
                const int64_t b = tt / tiles_per_mat, r = tt % tiles_per_mat;
                int m_tile, n_tile;
                tile_coords(r, p, m_tile, n_tile);
                const int row0 = m_tile * BM * CG + (int)rank * BM;
                const int col0 = n_tile * BN + (half_tile ? (int)((t - p.full_items) & 1) * (BN / 2) : 0) + (int)rank * ncols_cta;
                const int ba = p.a_batched ? (int)b : 0, bb = p.b_batched ? (int)b : 0;
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, p.debug, 0x100u + stage);
                    const int k0 = kb * BK;
                    if constexpr (Cfg::INK) {
                        // raw operands only, signalled on THIS CTA's full barrier (its converter warps wait on it)
                        mbar_expect_tx(full_bar(stage), (uint32_t)Cfg::TMA_BYTES);
                        tma_load_3d<1>(&tmA_hi, full_bar(stage), a_smem(stage, 0), k0, row0, ba);
#pragma unroll
                        for (int j = 0; j < Cfg::BN_CTA / 32; j++)
                            tma_load_3d<1>(&tmB_hi, full_bar(stage), b_smem(stage, 0) + j * 4096, col0 + j * 32, k0, bb);
                    } else {
                        if (leader) mbar_expect_tx(full_bar(stage), (uint32_t)(NPART * (Cfg::A_BYTES + ncols_cta * BK * Cfg::ESZ)) * CG);
                        tma_load_3d<CG>(&tmA_hi, full_bar(stage), a_smem(stage, 0), k0, row0, ba);
                        if (PASSES == 3) tma_load_3d<CG>(&tmA_lo, full_bar(stage), a_smem(stage, 1), k0, row0, ba);
#pragma unroll
                        for (int j = 0; j < Cfg::BN_CTA / Cfg::B_CHUNK_N; j++) {
                            if (j * Cfg::B_CHUNK_N >= ncols_cta) break;
                            tma_load_3d<CG>(&tmB_hi, full_bar(stage), b_smem(stage, 0) + j * Cfg::B_CHUNK_BYTES, col0 + j * Cfg::B_CHUNK_N, k0, bb);
                            if (PASSES == 3)
                                tma_load_3d<CG>(&tmB_lo, full_bar(stage), b_smem(stage, 1) + j * Cfg::B_CHUNK_BYTES, col0 + j * Cfg::B_CHUNK_N, k0, bb);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t FMT = Cfg::SCALED ? 0u : (Cfg::BF16 ? 1u : 2u);   // operand format: f16 / bf16 (kind::f16), tf32
            constexpr uint32_t BL = Cfg::BF16 ? LAYOUT_SW128 : LAYOUT_SW128_BASE32B;
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, tile_phase = 0;
            for (int64_t t = cluster_id; t < total_tiles; t += num_clusters) {
                const int nn = (Cfg::MERGED && t >= p.full_items) ? BN / 2 : BN;
                const uint32_t idesc = make_idesc(BM * CG, nn, FMT, FMT);
                if constexpr (!Cfg::CHUNKED) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.debug, 0x300u + acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < num_kb; kb++) {
                        mbar_wait(full_bar(stage), phase, p.debug, 0x200u + stage);
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / Cfg::UMMA_K; k++) {
                                const uint64_t da = make_smem_desc(a_smem(stage, 0) + k * 32, 16, 1024, LAYOUT_SW128);
                                const uint64_t db = make_smem_desc(b_smem(stage, 0) + k * Cfg::B_KSTEP, Cfg::B_CHUNK_BYTES, Cfg::B_SBO, BL);
                                umma_tf32<CG, Cfg::BF16>(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                            umma_commit<CG>(empty_bar(stage));
                            if (kb == num_kb - 1) umma_commit<CG>(tfull_bar(acc));
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                } else {
                    // cross-term accumulator of the previous tile must have been read out
                    if constexpr (!Cfg::MERGED) mbar_wait(cross_empty_bar, tile_phase ^ 1u, p.debug, 0x500u);
                    // An operand with +-inf: a_hi = inf times b_lo = 0 would turn cblas_sgemm's inf into NaN.  The pre-pass
                    // flags such calls; the cross terms then use (a_lo, b_lo) — finite, ~2^-22 of the result — so the
                    // call degrades to TF32x1 accuracy but keeps IEEE inf/NaN propagation identical to the reference.
                    const int hi_part = (!Cfg::INK && *reinterpret_cast<const volatile int *>(p.nonfinite) == p.nonfinite_gen) ? 1 : 0;
                    const uint32_t d_cross = tmem_base + (uint32_t)(2 * BN);
                    for (int kb0 = 0; kb0 < num_kb; kb0 += Cfg::KB_PER_CHUNK) {
                        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.debug, 0x300u + acc);
                        tc_fence_after();
                        const uint32_t d_main = tmem_base + (uint32_t)(acc * BN);
                        const uint32_t d_x = Cfg::MERGED ? d_main : d_cross;   // where the cross products go
                        const int kb1 = kb0 + Cfg::KB_PER_CHUNK < num_kb ? kb0 + Cfg::KB_PER_CHUNK : num_kb;
                        for (int kb = kb0; kb < kb1; kb++) {
                            if (Cfg::INK) mbar_wait(conv_bar(stage), phase, p.debug, 0x600u + stage);
                            else mbar_wait(full_bar(stage), phase, p.debug, 0x200u + stage);
                            tc_fence_after();
                            if (elect_one()) {
                                // a_lo.b_hi and a_hi.b_lo -> cross accumulator (whole K); a_hi.b_hi -> this chunk's accumulator
#pragma unroll
                                for (int k = 0; k < BK / Cfg::UMMA_K; k++) {
                                    const uint64_t da = make_smem_desc(a_smem(stage, 1) + k * 32, 16, 1024, LAYOUT_SW128);
                                    const uint64_t db = make_smem_desc(b_smem(stage, hi_part) + k * Cfg::B_KSTEP, Cfg::B_CHUNK_BYTES, Cfg::B_SBO, BL);
                                    umma_tf32<CG, Cfg::BF16>(d_x, da, db, idesc, ((Cfg::MERGED ? kb - kb0 : kb) | k) != 0 ? 1u : 0u);
                                }
#pragma unroll
                                for (int k = 0; k < BK / Cfg::UMMA_K; k++) {
                                    const uint64_t da = make_smem_desc(a_smem(stage, hi_part) + k * 32, 16, 1024, LAYOUT_SW128);
                                    const uint64_t db = make_smem_desc(b_smem(stage, 1) + k * Cfg::B_KSTEP, Cfg::B_CHUNK_BYTES, Cfg::B_SBO, BL);
                                    umma_tf32<CG, Cfg::BF16>(d_x, da, db, idesc, 1u);
                                }
#pragma unroll
                                for (int k = 0; k < BK / Cfg::UMMA_K; k++) {
                                    const uint64_t da = make_smem_desc(a_smem(stage, 0) + k * 32, 16, 1024, LAYOUT_SW128);
                                    const uint64_t db = make_smem_desc(b_smem(stage, 0) + k * Cfg::B_KSTEP, Cfg::B_CHUNK_BYTES, Cfg::B_SBO, BL);
                                    if constexpr (Cfg::SCALE_D) {   // chunk == this k-block: fold its cross products (x 2^11) first
                                        if (k == 0) umma_f16_scale_d11<CG>(d_main, da, db, idesc);
                                        else umma_tf32<CG, true>(d_main, da, db, idesc, 1u);
                                    } else {
                                        umma_tf32<CG, Cfg::BF16>(d_main, da, db, idesc, (Cfg::MERGED || kb != kb0 || k != 0) ? 1u : 0u);
                                    }
                                }
                                umma_commit<CG>(empty_bar(stage));
                                if (kb == kb1 - 1) umma_commit<CG>(tfull_bar(acc));
                            }
                            __syncwarp();
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                    }
                    tile_phase ^= 1u;
                }
            }
        }
    } else if (Cfg::INK && warp >= 6) {
        // ===================== lo-part converters (warps 6..9, TF32x3 in-kernel split) =====================
        // a_lo = rna_tf32(a - trunc_tf32(a)) element for element: the swizzled layout of the raw tile carries over
        // unchanged because the lo tile sits at the same offset modulo 1024 B.
        const int ct = threadIdx.x - 192;   // 0..127
        int stage = 0;
        uint32_t phase = 0;
        auto lo_of = [](float a) {
            const float hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a - hi));
            return __uint_as_float(r);
        };
        auto convert = [&](uint32_t src, uint32_t dst, int bytes) {
#pragma unroll 4
            for (int off = ct * 16; off < bytes; off += 128 * 16) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + off));
                v.x = lo_of(v.x); v.y = lo_of(v.y); v.z = lo_of(v.z); v.w = lo_of(v.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        };
        for (int64_t t = cluster_id; t < total_tiles; t += num_clusters) {
            for (int kb = 0; kb < num_kb; kb++) {
                mbar_wait(full_bar(stage), phase, p.debug, 0x700u + stage);
                convert(a_smem(stage, 0), a_smem(stage, 1), Cfg::A_BYTES);
                convert(b_smem(stage, 0), b_smem(stage, 1), Cfg::B_BYTES);
                fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_leader(conv_bar(stage));
                    else mbar_arrive_local(conv_bar(stage));
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (!Cfg::MERGED && warp >= Cfg::EPI_WARP0 && warp < Cfg::EPI_WARP0 + Cfg::EPI_WARPS) {
        if constexpr (!Cfg::MERGED) run_epilogue();
    }
    }

    // ---- teardown: everyone done with TMEM before it is released
    __syncwarp();  // reconverge the single-lane producer / issuer warps before aligned barriers
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
    }
    if (threadIdx.x == 0) trace_max(p.trace, tslot == 5 ? 7 : 11);
}

// ------------------------------------------------------------------ TF32 split pre-pass (x3)
// The tensor core reads an fp32 operand as TF32 by TRUNCATION (measured), so the raw array already is
// a_hi = trunc_tf32(a).  The pre-pass only writes the remainder  a_lo = rna_tf32(a - trunc_tf32(a))
// (a - trunc(a) is exact in fp32; the final rounding to TF32 is unbiased, relative 2^-21 of a).
// One launch covers both operands: 4 B read + 4 B written per operand element.
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Device word "an operand of call #gen contains +-inf".  Every x3 call gets a fresh tag (gemm_reset_nonfinite), the split
// kernels store the tag when they meet an inf and the GEMM compares against it: no per-call memset on the stream.
static int *nonfinite_flag() { return reinterpret_cast<int *>(ctx().dev_result) + 8; }
int gemm_reset_nonfinite() {
    int &gen = ctx().nonfinite_gen;   // lives in the context: a re-initialised context starts from a cleared word again
    if (gen == 0) NB_CUDA(cudaMemsetAsync(nonfinite_flag(), 0, 2 * sizeof(int), ctx().stream));   // once: defined start values ([1] = FP16x3 eligibility)
    if (++gen == 0x7FFFFFFF) gen = 1;
    return NB200_OK;
}
// +-inf has no finite remainder (inf - inf = NaN): its lo part is 0 and the call is flagged, see the MMA issuer.
__device__ __forceinline__ float lo_part(float a, int *nonfinite, int gen) {
    if (fabsf(a) == CUDART_INF_F) { *nonfinite = gen; return 0.f; }
    const float hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
    return to_tf32(a - hi);
}
__device__ __forceinline__ void split_tf32_body(const float *__restrict__ in0, float *__restrict__ lo0, int64_t n0,
                                                const float *__restrict__ in1, float *__restrict__ lo1, int64_t n1,
                                                int *__restrict__ nonfinite, int gen) {
    const int64_t g0 = (n0 + 3) >> 2, g1 = (n1 + 3) >> 2;   // 4-element groups (spans are padded to a multiple of 4)
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < g0 + g1; i += (int64_t)gridDim.x * 256) {
        const bool second = i >= g0;
        const int64_t j = second ? i - g0 : i;
        const float *in = second ? in1 : in0;
        float *lo = second ? lo1 : lo0;
        const int64_t n = second ? n1 : n0;
        if ((j << 2) + 4 <= n) {
            float4 a = ld_ew(reinterpret_cast<const float4 *>(in) + j), l;
            l.x = lo_part(a.x, nonfinite, gen); l.y = lo_part(a.y, nonfinite, gen); l.z = lo_part(a.z, nonfinite, gen); l.w = lo_part(a.w, nonfinite, gen);
            reinterpret_cast<float4 *>(lo)[j] = l;
        } else {
            for (int64_t e = j << 2; e < n; e++) lo[e] = lo_part(in[e], nonfinite, gen);
        }
    }
}
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ in0, float *__restrict__ lo0, int64_t n0,
                                                         const float *__restrict__ in1, float *__restrict__ lo1, int64_t n1,
                                                         int *__restrict__ nonfinite, int gen) {
    pdl_launch_dependents();   // the GEMM that follows may set itself up while the split is still running (it waits before reading)
    split_tf32_body(in0, lo0, n0, in1, lo1, n1, nonfinite, gen);
}

// ------------------------------------------------------------------ BF16 split pre-pass (BF16x3)
// a = a1 + a2 + r with a1 = rn_bf16(a), a2 = rn_bf16(a - a1) (a - a1 is exact in fp32), |r| <= 2^-18 |a|.
// The matrices are rewritten as packed bf16 [batch*rows][ld_out] (ld_out = cols rounded up to 8 so that every row
// starts 16-byte aligned for the TMA); 4 B read + 4 B written per element, like the TF32 pre-pass.
struct SplitSpan {
    const float *in;
    __nv_bfloat16 *hi, *lo;
    int64_t rows_per, nrows, cols, ld_in, stride_in, ld_out;   // nrows = batch * rows_per
    int64_t gpr, groups;                                        // 4-element groups per row / in total
    int vec;                                                    // 16-byte aligned source rows
    int flat;                                                   // contiguous, cols % 8 == 0: output index == input index
};
__device__ __forceinline__ void split_bf16(float a, __nv_bfloat16 &h, __nv_bfloat16 &l, int *nonfinite, int gen) {
    uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) << 16;
    const uint32_t ab = __float_as_uint(a);
    if ((hb & 0x7FFFFFFFu) == 0x7F800000u) {
        if ((ab & 0x7FFFFFFFu) == 0x7F800000u) {   // +-inf: no finite remainder, see the MMA issuer
            *nonfinite = gen;
            h = __ushort_as_bfloat16((unsigned short)(hb >> 16));
            l = __ushort_as_bfloat16((unsigned short)0);
            return;
        }
        hb = ab & 0xFFFF0000u;                     // finite value that rounds up to inf in bf16: truncate instead
    }
    h = __ushort_as_bfloat16((unsigned short)(hb >> 16));
    l = __float2bfloat16_rn(a - __uint_as_float(hb));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// Contiguous operands whose row length is a multiple of 8 keep their flat indexing (ld_out == cols): 8 elements per
// thread, 2 x 16-byte loads and 2 x 16-byte stores, one group per thread, no index arithmetic.
__global__ void __launch_bounds__(256) split_bf16_flat_kernel(const float *__restrict__ in0, __nv_bfloat16 *__restrict__ hi0,
                                                              __nv_bfloat16 *__restrict__ lo0, int64_t g0,
                                                              const float *__restrict__ in1, __nv_bfloat16 *__restrict__ hi1,
                                                              __nv_bfloat16 *__restrict__ lo1, int64_t g1,
                                                              int *__restrict__ nonfinite, int gen) {
    pdl_launch_dependents();   // see split_tf32_kernel
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < g0 + g1; i += (int64_t)gridDim.x * 256) {
        const bool second = i >= g0;
        const int64_t j = second ? i - g0 : i;
        const float4 *src = reinterpret_cast<const float4 *>(second ? in1 : in0) + 2 * j;
        const float4 a = ld_ew(src), b = ld_ew(src + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t hp[4], lp[4], any = 0;
        // fast path, two elements per cvt: h = rn_bf16(v), l = rn_bf16(v - h).  A remainder with an all-ones exponent
        // (h overflowed to inf, or v is inf/NaN) sends the whole group through the careful per-element path.
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
            hp[q] = *reinterpret_cast<const uint32_t *>(&h2);
            const float r0 = v[2 * q] - __uint_as_float(hp[q] << 16), r1 = v[2 * q + 1] - __uint_as_float(hp[q] & 0xFFFF0000u);
            any |= __float_as_uint(r0) | __float_as_uint(r1);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
            lp[q] = *reinterpret_cast<const uint32_t *>(&l2);
        }
        if ((any & 0x7F800000u) == 0x7F800000u) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(v[2 * q], h0, l0, nonfinite, gen);
                split_bf16(v[2 * q + 1], h1, l1, nonfinite, gen);
                hp[q] = pack_bf16(h0, h1);
                lp[q] = pack_bf16(l0, l1);
            }
        }
        const uint4 hv = make_uint4(hp[0], hp[1], hp[2], hp[3]), lv = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        reinterpret_cast<uint4 *>(second ? hi1 : hi0)[j] = hv;
        reinterpret_cast<uint4 *>(second ? lo1 : lo0)[j] = lv;
    }
}
__global__ void __launch_bounds__(256) split_bf16_kernel(const SplitSpan s0, const SplitSpan s1, int *__restrict__ nonfinite, int gen) {
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < s0.groups + s1.groups; i += (int64_t)gridDim.x * 256) {
        const bool second = i >= s0.groups;
        const SplitSpan &s = second ? s1 : s0;
        const int64_t j = second ? i - s0.groups : i;
        const int64_t r = j / s.gpr, c = (j - r * s.gpr) << 2;
        const int64_t b = r / s.rows_per, rr = r - b * s.rows_per;
        const float *src = s.in + b * s.stride_in + rr * s.ld_in + c;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (s.vec && c + 4 <= s.cols) {
            const float4 a = ld_ew(reinterpret_cast<const float4 *>(src));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else {
            for (int e = 0; e < 4; e++) if (c + e < s.cols) v[e] = src[e];
        }
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; e++) split_bf16(v[e], h[e], l[e], nonfinite, gen);
        // ld_out is a multiple of 8 and c of 4: 8-byte aligned stores; columns in [cols, ld_out) receive zeros
        const int64_t o = r * s.ld_out + c;
        uint2 hv, lv;
        hv.x = pack_bf16(h[0], h[1]); hv.y = pack_bf16(h[2], h[3]);
        lv.x = pack_bf16(l[0], l[1]); lv.y = pack_bf16(l[2], l[3]);
        *reinterpret_cast<uint2 *>(s.hi + o) = hv;
        *reinterpret_cast<uint2 *>(s.lo + o) = lv;
    }
}

// ------------------------------------------------------------------ FP16x3 pre-pass: per-row / per-column |max|, scaled half split
// Largest FINITE magnitude of every row of A and every column of B, as fp32 bit patterns (monotonic as unsigned ints),
// combined with atomicMax into zero-initialised arrays.  +-inf / NaN are skipped here; the split flags them.
__device__ __forceinline__ unsigned int finite_abs_bits(float v) {
    const unsigned int b = __float_as_uint(v) & 0x7FFFFFFFu;
    return b < 0x7F800000u ? b : 0u;
}
__device__ __forceinline__ unsigned int abs4_bits(const float4 &v) {
    return max(max(finite_abs_bits(v.x), finite_abs_bits(v.y)), max(finite_abs_bits(v.z), finite_abs_bits(v.w)));
}
// one warp per row; rows = batch * rows_per
__global__ void __launch_bounds__(256) absmax_rows_kernel(const float *__restrict__ in, int64_t rows_per, int64_t nrows, int64_t cols,
                                                          int64_t ld_in, int64_t stride_in, unsigned int *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < nrows; r += (int64_t)gridDim.x * 8) {
        const int64_t b = r / rows_per, rr = r - b * rows_per;
        const float *src = in + b * stride_in + rr * ld_in;
        unsigned int m = 0;
        if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (cols & 3) == 0) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            const int64_t n4 = cols >> 2;
            int64_t c = lane;
            for (; c + 96 < n4; c += 128) {   // four independent 16-byte loads in flight per lane
                const float4 v0 = ldg_stream(s4 + c), v1 = ldg_stream(s4 + c + 32), v2 = ldg_stream(s4 + c + 64), v3 = ldg_stream(s4 + c + 96);
                m = max(m, max(max(abs4_bits(v0), abs4_bits(v1)), max(abs4_bits(v2), abs4_bits(v3))));
            }
            for (; c < n4; c += 32) m = max(m, abs4_bits(ldg_stream(s4 + c)));
        } else {
            for (int64_t c = lane; c < cols; c += 32) m = max(m, finite_abs_bits(ldg_stream(src + c)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if (lane == 0) out[r] = m;
    }
}
// threads along the contiguous columns, blockIdx.y splits the rows, blockIdx.z = batch
__global__ void __launch_bounds__(256) absmax_cols_kernel(const float *__restrict__ in, int64_t rows, int64_t cols, int64_t ld_in,
                                                          int64_t stride_in, unsigned int *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    const float *src = in + (int64_t)blockIdx.z * stride_in + c;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    unsigned int m = 0;
    for (int64_t r = r0; r < r1; r++) m = max(m, finite_abs_bits(ldg_stream(src + r * ld_in)));
    if (m) atomicMax(out + (int64_t)blockIdx.z * cols + c, m);
}
// same, four adjacent columns per thread with 16-byte loads (needs 16-byte aligned rows: ld % 4 == 0, cols % 4 == 0).
// Default-policy loads on purpose: the split that follows re-reads B, and whatever of it is still in L2 is not fetched
// from HBM again.  Eight loads in flight per thread; the host sizes the row segments so that ~2 blocks per SM run
// (few, long segments: the final atomicMax traffic is 4 per thread).
__global__ void __launch_bounds__(256) absmax_cols4_kernel(const float *__restrict__ in, int64_t rows, int64_t cols, int64_t ld_in,
                                                           int64_t stride_in, unsigned int *__restrict__ out) {
    const int64_t c4 = (int64_t)blockIdx.x * 256 + threadIdx.x;   // column group
    if (c4 * 4 >= cols) return;
    const float4 *src = reinterpret_cast<const float4 *>(in + (int64_t)blockIdx.z * stride_in) + c4;
    const int64_t ld4 = ld_in >> 2;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    unsigned int mx = 0, my = 0, mz = 0, mw = 0;
    int64_t r = r0;
    for (; r + 7 < r1; r += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = ld_ew(src + (r + u) * ld4);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            mx = max(mx, finite_abs_bits(v[u].x)); my = max(my, finite_abs_bits(v[u].y));
            mz = max(mz, finite_abs_bits(v[u].z)); mw = max(mw, finite_abs_bits(v[u].w));
        }
    }
    for (; r < r1; r++) {
        const float4 v = ld_ew(src + r * ld4);
        mx = max(mx, finite_abs_bits(v.x)); my = max(my, finite_abs_bits(v.y)); mz = max(mz, finite_abs_bits(v.z)); mw = max(mw, finite_abs_bits(v.w));
    }
    unsigned int *o = out + (int64_t)blockIdx.z * cols + c4 * 4;
    if (mx) atomicMax(o, mx);
    if (my) atomicMax(o + 1, my);
    if (mz) atomicMax(o + 2, mz);
    if (mw) atomicMax(o + 3, mw);
}
// a' = a * 2^e (e from the row or column |max|), hi = rn_f16(a'), |a'| < 2^15; lo = rn_f16((a' - hi) * 2^11): a' - hi is exact
// in fp32, |.| * 2^11 <= 2^15 so nothing overflows, and the remainder a' - hi - lo * 2^-11 is <= 2^-22 |a'|: the same class as
// TF32x3.  The bound needs hi to be a NORMAL half (|a'| >= 2^-14, i.e. the element lies within 2^-28 of its row / column maximum).
// Out-of-window elements do not abandon the call as long as there are few of them: the split records (row, column, a),
// zeroes the element's half parts, and the post kernel adds a times the partner row / column of the other (raw fp32)
// operand to C after the GEMM (a sparse rank-1 repair in full fp32 precision).  Only when a record list overflows is
// the call marked ineligible (nonfinite[1] = gen) and left to the gated fallback (see gemm_fp16x3).
constexpr int FIX_CAP = 4096;              // records per operand and chunk
struct FixList {
    unsigned int *count;                   // device counter (reset by the host before the operand is split)
    int4 *recs;                            // {row (over batch * rows), column, float bits of d, 0}
};
// MIX (GemmCfg::MIXLO): the lo part is stored UNSCALED, lo = rn_f16(a' - hi): remainder <= max(2^-22 |a'|, 2^-25), and the window
// closes at |a'| = 2^-6 (remainder <= 2^-19 |a'|) instead of 2^-14.
template <bool MIX>
__device__ __forceinline__ void split_f16(float a, int e, unsigned short &h, unsigned short &l, int *nonfinite, int gen,
                                          const FixList &fix, int64_t row, int64_t col) {
    const unsigned int ab = __float_as_uint(a) & 0x7FFFFFFFu;
    if (ab >= 0x7F800000u) {                       // +-inf (flag: see the MMA issuer) or NaN: hi carries it, lo = 0
        if (ab == 0x7F800000u) *nonfinite = gen;
        h = __half_as_ushort(__float2half_rn(a));
        l = 0;
        return;
    }
    const float x = scale_pow2(a, e);
    const __half hh = __float2half_rn(x);
    h = __half_as_ushort(hh);
    l = __half_as_ushort(__float2half_rn((x - __half2float(hh)) * (MIX ? 1.0f : 2048.0f)));
    if (ab != 0u && fabsf(x) < (MIX ? 0.015625f : 6.103515625e-05f)) {  // below 2^-14 (MIX: 2^-6): hi (MIX: lo) would lose bits as a subnormal half -> repair record
        // The element leaves the GEMM entirely (hi = lo = 0) and is carried by the record in full fp32: a lo-only
        // representation would multiply it with the partner's hi part alone, i.e. with 11 bits (measured 4.7e-4).
        h = 0;
        l = 0;
        const unsigned int idx = atomicAdd(fix.count, 1u);
        if (idx < (unsigned int)FIX_CAP && row < 0x7FFFFFFF) fix.recs[idx] = make_int4((int)row, (int)col, __float_as_int(a), 0);
        else nonfinite[1] = gen;
    }
}
struct SplitSpanF16 {
    SplitSpan s;
    const unsigned int *max_bits;   // per row (by_col == 0: index = output row) or per column (index = batch * cols + column)
    int by_col;
    FixList fix;
};
// General operands (any alignment / leading dimension; rows are repacked to ld_out = cols rounded up to 8): four elements per thread.
template <bool MIX>
__global__ void __launch_bounds__(256) split_f16_kernel(const SplitSpanF16 s0, const SplitSpanF16 s1, int *__restrict__ nonfinite, int gen) {
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < s0.s.groups + s1.s.groups; i += (int64_t)gridDim.x * 256) {
        const bool second = i >= s0.s.groups;
        const SplitSpanF16 &sp = second ? s1 : s0;
        const SplitSpan &s = sp.s;
        const int64_t j = second ? i - s0.s.groups : i;
        const int64_t r = j / s.gpr, c = (j - r * s.gpr) << 2;
        const int64_t b = r / s.rows_per, rr = r - b * s.rows_per;
        const float *src = s.in + b * s.stride_in + rr * s.ld_in + c;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (s.vec && c + 4 <= s.cols) {
            const float4 a = ld_ew(reinterpret_cast<const float4 *>(src));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else {
            for (int e = 0; e < 4; e++) if (c + e < s.cols) v[e] = src[e];
        }
        const int e_row = sp.by_col ? 0 : scale_exp(sp.max_bits[r]);
        unsigned short h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int ex = sp.by_col ? (c + e < s.cols ? scale_exp(sp.max_bits[b * s.cols + c + e]) : 0) : e_row;
            split_f16<MIX>(v[e], ex, h[e], l[e], nonfinite, gen, sp.fix, r, c + e);
        }
        const int64_t o = r * s.ld_out + c;
        uint2 hv, lv;
        hv.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16); hv.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
        lv.x = (uint32_t)l[0] | ((uint32_t)l[1] << 16); lv.y = (uint32_t)l[2] | ((uint32_t)l[3] << 16);
        *reinterpret_cast<uint2 *>(s.hi + o) = hv;
        *reinterpret_cast<uint2 *>(s.lo + o) = lv;
    }
}

// Four adjacent elements (row r, columns c .. c+3), each with the two exact power-of-two factors of its scaling (scale_factors),
// -> packed hi / lo half words.  Fast path: packed conversions; an element that is +-inf / NaN or (non-zero and) below 2^-14 after
// scaling sends the quad through the careful per-element path (flags, repair records), which is kept out of line.
struct Pow2Pair { float f1, f2; };   // 2^e as f1 * f2 (2^e itself may not be representable)
__device__ __forceinline__ Pow2Pair scale_factors(int e) {
    const int e1 = e / 2, e2 = e - e1;
    return {__int_as_float((e1 + 127) << 23), __int_as_float((e2 + 127) << 23)};
}
// returns {hi.x, hi.y, lo.x, lo.y} by value (pointer outputs would pin the caller's fast-path results to local memory)
template <bool MIX>
__device__ __noinline__ uint4 split4_careful(float4 v, int e0, int e1, int e2, int e3, int *nonfinite, int gen,
                                             unsigned int *fix_count, int4 *fix_recs, int64_t r, int64_t c) {
    const FixList fix{fix_count, fix_recs};
    unsigned short h[4], l[4];
    split_f16<MIX>(v.x, e0, h[0], l[0], nonfinite, gen, fix, r, c);
    split_f16<MIX>(v.y, e1, h[1], l[1], nonfinite, gen, fix, r, c + 1);
    split_f16<MIX>(v.z, e2, h[2], l[2], nonfinite, gen, fix, r, c + 2);
    split_f16<MIX>(v.w, e3, h[3], l[3], nonfinite, gen, fix, r, c + 3);
    return make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                      (uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}
// not in [2^-14, inf) (MIX: [2^-6, inf)) and not zero: the element needs the careful path
template <bool MIX>
__device__ __forceinline__ bool outside_window(float x) {
    const unsigned int b = __float_as_uint(x) & 0x7FFFFFFFu;
    constexpr unsigned int LOW = MIX ? 0x3C800000u : 0x38800000u;
    return (b - LOW) >= (0x7F800000u - LOW) && b != 0u;
}
template <bool MIX>
__device__ __forceinline__ bool split4_fast(const float4 &v, const Pow2Pair &s0, const Pow2Pair &s1, const Pow2Pair &s2, const Pow2Pair &s3,
                                            uint2 &hi, uint2 &lo) {
    const float x0 = v.x * s0.f1 * s0.f2, x1 = v.y * s1.f1 * s1.f2, x2 = v.z * s2.f1 * s2.f2, x3 = v.w * s3.f1 * s3.f2;
    const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    hi = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
    constexpr float LS = MIX ? 1.0f : 2048.0f;   // the lo parts are stored times 2^11 unless MIX
    const __half2 l01 = __floats2half2_rn((x0 - f01.x) * LS, (x1 - f01.y) * LS);
    const __half2 l23 = __floats2half2_rn((x2 - f23.x) * LS, (x3 - f23.y) * LS);
    lo = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
    return outside_window<MIX>(x0) | outside_window<MIX>(x1) | outside_window<MIX>(x2) | outside_window<MIX>(x3);
}

// ---- The flat FP16x3 pre-pass (contiguous operands, cols % 8 == 0, 16-byte aligned: output index == input index): ONE persistent
// launch whose CTAs are all co-resident (grid sized by the occupancy API); the CTAs are divided between the two operands in
// proportion to their sizes (interleaved, so every SM hosts both kinds) and both parts stream concurrently:
//   A-CTAs    rows of A, each held in REGISTERS between its |max| reduction and its split (32 / 64 / 128 / 256 threads per row,
//             two 8-element groups per thread; rows longer than 4096 are re-read from L1 / L2 by the CTA that just read them).
//             A is read from HBM once; row_max is written for the GEMM epilogue.
//   B-CTAs    phase 1: column maxima over (matrix, 1024-column block, row segment) tiles, eight 16-byte loads in flight per
//             thread, atomicMax into the zeroed col_max; barrier among the B-CTAs (arrival counter zeroed by the same memset as
//             col_max; bounded spin); phase 2: the SAME CTA splits the SAME tile with the same thread-to-column mapping, rows in
//             reverse order: the four scale factors of a thread's columns are computed once, and the tile comes back from the
//             SM's own L1 / the near L2 partition that phase 1 pulled it into instead of from HBM.
// Every CTA calls griddepcontrol.launch_dependents as soon as it no longer needs the rest of the grid (A-CTAs at once, B-CTAs
// after the barrier): the GEMM's prologue (barriers, TMEM allocation, tensor-map prefetch) overlaps the tail.
struct PrepCoop {
    SplitSpanF16 a, b;
    int64_t a_rows;            // rows of A over the batch (0: A is not prepared by this launch)
    int64_t b_mats;            // matrices of B (0: B is not prepared by this launch)
    int n_b;                   // CTAs working on B (0 <= n_b <= gridDim.x), evenly interleaved with the A-CTAs
    unsigned int *row_max_out;
    unsigned int *barrier;     // zero on entry
    unsigned int *zero_ptr;    // control block of the NEXT call (other parity): zero_words words are cleared here, so that call
    int64_t zero_words;        //   needs no memset.  Its last users (the previous call's kernels) are complete: stream order.
    unsigned int *debug;       // pinned host words: [0] = timeout flag, [1] = tag
    unsigned long long *trace; // optional timeline stamps (common.cuh)
};
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int abs8_bits(const float4 &a, const float4 &b) { return max(abs4_bits(a), abs4_bits(b)); }
// one 8-element group of a row of A (both quads share the row's factors) -> 16-byte stores
template <bool MIX>
__device__ __forceinline__ void store_group_a(const float4 &a, const float4 &b, int e, const Pow2Pair &sc, uint4 *hi, uint4 *lo, int *nonfinite,
                                              int gen, const FixList &fix, int64_t r, int64_t c, uint64_t pol) {
    uint2 h0, l0, h1, l1;
    const bool bad0 = split4_fast<MIX>(a, sc, sc, sc, sc, h0, l0), bad1 = split4_fast<MIX>(b, sc, sc, sc, sc, h1, l1);
    if (bad0) { const uint4 w = split4_careful<MIX>(a, e, e, e, e, nonfinite, gen, fix.count, fix.recs, r, c); h0 = make_uint2(w.x, w.y); l0 = make_uint2(w.z, w.w); }
    if (bad1) { const uint4 w = split4_careful<MIX>(b, e, e, e, e, nonfinite, gen, fix.count, fix.recs, r, c + 4); h1 = make_uint2(w.x, w.y); l1 = make_uint2(w.z, w.w); }
    st_l2(hi, make_uint4(h0.x, h0.y, h1.x, h1.y), pol);
    st_l2(lo, make_uint4(l0.x, l0.y, l1.x, l1.y), pol);
}
// Rows of A held in registers: TPR threads per row, 256 / TPR rows per CTA pass, up to two groups per thread (row length <= 16 * TPR).
// Software-pipelined over the passes: the loads of pass i+1 are issued before pass i is reduced, split and stored, so a CTA always
// has a row batch in flight (without it every pass was load -> wait -> compute -> store with nothing outstanding during the last
// three: ncu showed the pre-pass at 50 % of the DRAM rate, latency-bound).
template <int TPR, bool MIX>
__device__ __forceinline__ void prep_a_rows(const PrepCoop &q, int rank, int n_ctas, int *nonfinite, int gen, unsigned int (*red)[8], uint64_t pol) {
    constexpr int RPC = 256 / TPR;                         // rows per CTA pass
    const int t = threadIdx.x % TPR, rg = threadIdx.x / TPR, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n8 = q.a.s.cols >> 3;
    const int64_t stride = (int64_t)n_ctas * RPC;
    const float4 *in4 = reinterpret_cast<const float4 *>(q.a.s.in);
    auto load = [&](int64_t base, float4 (&dst)[2][2]) {
        const int64_t r = base + rg;
        if (r >= q.a_rows) return;
        const float4 *s4 = in4 + r * (2 * n8);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int64_t g = t + u * TPR;
            if (g < n8) { dst[u][0] = ldg_stream_l2(s4 + 2 * g, pol); dst[u][1] = ldg_stream_l2(s4 + 2 * g + 1, pol); }
        }
    };
    float4 v[2][2], nv[2][2];
    int64_t base = (int64_t)rank * RPC;
    if (base < q.a_rows) load(base, v);
    int it = 0;
    for (; base < q.a_rows; base += stride, it++) {   // uniform trip count over the CTA
        if (base + stride < q.a_rows) load(base + stride, nv);
        const int64_t r = base + rg;
        const bool live = r < q.a_rows;
        unsigned int m = 0;
#pragma unroll
        for (int u = 0; u < 2; u++)
            if (live && t + u * TPR < n8) m = max(m, abs8_bits(v[u][0], v[u][1]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if constexpr (TPR > 32) {
            if (lane == 0) red[it & 1][warp] = m;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < TPR / 32; w++) m = max(m, red[it & 1][rg * (TPR / 32) + w]);
        }
        if (live) {
            if (t == 0) q.row_max_out[r] = m;
            const int e = scale_exp(m);
            const Pow2Pair sc = scale_factors(e);
            uint4 *hi = reinterpret_cast<uint4 *>(q.a.s.hi) + r * n8, *lo = reinterpret_cast<uint4 *>(q.a.s.lo) + r * n8;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t g = t + u * TPR;
                if (g < n8) store_group_a<MIX>(v[u][0], v[u][1], e, sc, hi + g, lo + g, nonfinite, gen, q.a.fix, r, g << 3, pol);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) { v[u][0] = nv[u][0]; v[u][1] = nv[u][1]; }
    }
}
// rows longer than 4096: one CTA per row, pass 1 reduces the |max|, pass 2 re-reads the row (L1 / L2: this CTA just read it) and splits
template <bool MIX>
__device__ __forceinline__ void prep_a_rows_long(const PrepCoop &q, int rank, int n_ctas, int *nonfinite, int gen, unsigned int (*red)[8], uint64_t pol) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t n8 = q.a.s.cols >> 3;
    int it = 0;
    for (int64_t r = rank; r < q.a_rows; r += n_ctas, it++) {
        const float4 *s4 = reinterpret_cast<const float4 *>(q.a.s.in) + r * (2 * n8);
        unsigned int m = 0;
        for (int64_t g0 = 0; g0 < n8; g0 += 512) {
            float4 v[2][2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t g = g0 + t + u * 256;
                if (g < n8) { v[u][0] = ld_ew(s4 + 2 * g); v[u][1] = ld_ew(s4 + 2 * g + 1); }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (g0 + t + u * 256 < n8) m = max(m, abs8_bits(v[u][0], v[u][1]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if (lane == 0) red[it & 1][warp] = m;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; w++) m = max(m, red[it & 1][w]);
        if (t == 0) q.row_max_out[r] = m;
        const int e = scale_exp(m);
        const Pow2Pair sc = scale_factors(e);
        uint4 *hi = reinterpret_cast<uint4 *>(q.a.s.hi) + r * n8, *lo = reinterpret_cast<uint4 *>(q.a.s.lo) + r * n8;
        for (int64_t g0 = 0; g0 < n8; g0 += 512) {
            float4 v[2][2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t g = g0 + t + u * 256;
                if (g < n8) { v[u][0] = ld_ew(s4 + 2 * g); v[u][1] = ld_ew(s4 + 2 * g + 1); }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t g = g0 + t + u * 256;
                if (g < n8) store_group_a<MIX>(v[u][0], v[u][1], e, sc, hi + g, lo + g, nonfinite, gen, q.a.fix, r, g << 3, pol);
            }
        }
    }
}
template <bool MIX>
__global__ void __launch_bounds__(256, 3) prep16_coop_kernel(const PrepCoop q, int *__restrict__ nonfinite, int gen) {
    __shared__ unsigned int red[2][8];
    const int t = threadIdx.x;
    pdl_wait();   // launched with programmatic stream serialization: the operands / control blocks belong to earlier work until here
    if (t == 0) trace_min(q.trace, 0);
    if (blockIdx.x == gridDim.x - 1)
        for (int64_t w = t; w < q.zero_words; w += 256) q.zero_ptr[w] = 0u;
    // Everything except B's first read is stream-once traffic: marked evict_first so that B (read again in phase 2) stays in L2.
    const uint64_t pol = l2_policy_evict_first();
    // role: n_b of the gridDim.x CTAs work on B, spread evenly over the block indices
    const int G = (int)gridDim.x, i = (int)blockIdx.x;
    const int b_before = (int)(((int64_t)i * q.n_b) / G);
    const bool is_b = (int)(((int64_t)(i + 1) * q.n_b) / G) > b_before;
    if (!is_b) {
        pdl_launch_dependents();   // nothing here needs the rest of the grid
        const int rank = i - b_before, n_a = G - q.n_b;
        const int64_t n8 = q.a.s.cols >> 3;
        if (q.a_rows > 0) {
            if (n8 <= 64) prep_a_rows<32, MIX>(q, rank, n_a, nonfinite, gen, red, pol);
            else if (n8 <= 128) prep_a_rows<64, MIX>(q, rank, n_a, nonfinite, gen, red, pol);
            else if (n8 <= 256) prep_a_rows<128, MIX>(q, rank, n_a, nonfinite, gen, red, pol);
            else if (n8 <= 512) prep_a_rows<256, MIX>(q, rank, n_a, nonfinite, gen, red, pol);
            else prep_a_rows_long<MIX>(q, rank, n_a, nonfinite, gen, red, pol);
        }
        if (t == 0) { trace_max(q.trace, 1); trace_max(q.trace, 4); }
        return;
    }
    // ---------------- B-CTAs.  Tile decomposition shared by both phases.
    const int rank = b_before, n_b = q.n_b;
    const int64_t K = q.b.s.rows_per, N = q.b.s.cols, n4 = N >> 2;
    unsigned int *col_max = const_cast<unsigned int *>(q.b.max_bits);
    int cw = 256;                                 // threads across the columns (16 bytes each); the rest stack up along the rows
    while (cw > 32 && (cw >> 1) >= n4) cw >>= 1;
    const int ty = t / cw, tx = t - ty * cw, TY = 256 / cw;
    const int64_t ncb = (n4 + cw - 1) / cw;
    int64_t nseg = ((int64_t)n_b + q.b_mats * ncb - 1) / (q.b_mats * ncb);
    const int64_t max_seg = (K + 8 * TY - 1) / (8 * TY);
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    const int64_t per = (K + nseg - 1) / nseg;
    const int64_t tasks = q.b_mats * nseg * ncb;
    // phase 1: column maxima
    for (int64_t task = rank; task < tasks; task += n_b) {
        const int64_t cb = task % ncb, rest = task / ncb, seg = rest % nseg, mat = rest / nseg;
        const int64_t c4 = cb * cw + tx;
        if (c4 >= n4) continue;
        const float4 *src = reinterpret_cast<const float4 *>(q.b.s.in) + mat * K * n4 + c4;
        const int64_t r0 = seg * per, r1 = r0 + per < K ? r0 + per : K;
        unsigned int mx = 0, my = 0, mz = 0, mw = 0;
        int64_t r = r0 + ty;
        for (; r + 7 * TY < r1; r += 8 * TY) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = ld_ew(src + (r + u * TY) * n4);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                mx = max(mx, finite_abs_bits(v[u].x)); my = max(my, finite_abs_bits(v[u].y));
                mz = max(mz, finite_abs_bits(v[u].z)); mw = max(mw, finite_abs_bits(v[u].w));
            }
        }
        for (; r < r1; r += TY) {
            const float4 v = ld_ew(src + r * n4);
            mx = max(mx, finite_abs_bits(v.x)); my = max(my, finite_abs_bits(v.y)); mz = max(mz, finite_abs_bits(v.z)); mw = max(mw, finite_abs_bits(v.w));
        }
        unsigned int *o = col_max + mat * N + c4 * 4;
        if (mx) atomicMax(o, mx);
        if (my) atomicMax(o + 1, my);
        if (mz) atomicMax(o + 2, mz);
        if (mw) atomicMax(o + 3, mw);
    }
    // barrier among the B-CTAs (all CTAs of the grid are resident: the host sizes it with the occupancy API)
    __syncthreads();
    if (t == 0) {
        trace_max(q.trace, 2);
        __threadfence();
        atomicAdd(q.barrier, 1u);
        unsigned int spins = 0;
        while (ld_acquire_u32(q.barrier) < (unsigned int)n_b) {
            __nanosleep(32);
            if (++spins > (1u << 24)) {   // a protocol bug must surface as an error, never as a hung GPU
                if (q.debug) { q.debug[0] = 1u; q.debug[1] = 0x900u; q.debug[2] = blockIdx.x; }
                __threadfence_system();
                asm volatile("trap;");
            }
        }
    }
    __syncthreads();
    pdl_launch_dependents();
    if (t == 0) trace_max(q.trace, 3);
    // phase 2: split the same tiles, last task / last rows first (most recently read = most likely still in L1 / L2)
    const int64_t my_tasks = rank < tasks ? (tasks - 1 - rank) / n_b + 1 : 0;
    for (int64_t k = my_tasks - 1; k >= 0; k--) {
        const int64_t task = rank + k * n_b;
        const int64_t cb = task % ncb, rest = task / ncb, seg = rest % nseg, mat = rest / nseg;
        const int64_t c4 = cb * cw + tx;
        if (c4 >= n4) continue;
        const uint4 mb = __ldcg(reinterpret_cast<const uint4 *>(col_max + mat * N) + c4);
        const int e0 = scale_exp(mb.x), e1 = scale_exp(mb.y), e2 = scale_exp(mb.z), e3 = scale_exp(mb.w);
        const Pow2Pair s0 = scale_factors(e0), s1 = scale_factors(e1), s2 = scale_factors(e2), s3 = scale_factors(e3);
        const float4 *src = reinterpret_cast<const float4 *>(q.b.s.in) + mat * K * n4 + c4;
        uint2 *hi = reinterpret_cast<uint2 *>(q.b.s.hi) + mat * K * n4 + c4, *lo = reinterpret_cast<uint2 *>(q.b.s.lo) + mat * K * n4 + c4;
        const int64_t r0 = seg * per, r1 = r0 + per < K ? r0 + per : K;
        // rows r1-1-ty, r1-1-ty-TY, ... >= r0, eight in flight
        int64_t r = r1 - 1 - ty;
        for (; r - 7 * TY >= r0; r -= 8 * TY) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = ldg_stream_l2(src + (r - u * TY) * n4, pol);   // last use of the raw tile
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int64_t rr = r - u * TY;
                uint2 h, l;
                if (split4_fast<MIX>(v[u], s0, s1, s2, s3, h, l)) {
                    const uint4 w = split4_careful<MIX>(v[u], e0, e1, e2, e3, nonfinite, gen, q.b.fix.count, q.b.fix.recs, mat * K + rr, c4 * 4);
                    h = make_uint2(w.x, w.y); l = make_uint2(w.z, w.w);
                }
                st_l2(hi + rr * n4, h, pol);
                st_l2(lo + rr * n4, l, pol);
            }
        }
        for (; r >= r0; r -= TY) {
            const float4 v = ldg_stream_l2(src + r * n4, pol);
            uint2 h, l;
            if (split4_fast<MIX>(v, s0, s1, s2, s3, h, l)) {
                const uint4 w = split4_careful<MIX>(v, e0, e1, e2, e3, nonfinite, gen, q.b.fix.count, q.b.fix.recs, mat * K + r, c4 * 4);
                h = make_uint2(w.x, w.y); l = make_uint2(w.z, w.w);
            }
            st_l2(hi + r * n4, h, pol);
            st_l2(lo + r * n4, l, pol);
        }
    }
    if (t == 0) trace_max(q.trace, 4);
}

// (Tried and removed, measured on one box, profiles/r2_summary.md: doing the repair in the merged GEMM's own epilogue - records in
//  shared memory, the warp walking a matching row segment together - so that this kernel always leaves at once.  The tail after the
//  GEMM shrank from 12 to 2.5 us, but the read-modify-writes at the end of the affected tiles (a record touches 16 tiles) sit on the
//  epilogue's critical path: 4096^3 0.336 -> 0.340 ms, 2048^3 / 8192^3 unchanged.  Releasing the dependents of this kernel early
//  changed nothing either.)
// After the FP16x3 GEMM, one launch, exactly one of two jobs:
//  * the call stayed eligible: sparse repair of the recorded out-of-window elements (see split_f16), one block per
//    record.  A-record (i, k, d): C[i, :] += d * B[k, :]; B-record (k, j, d): C[:, j] += A[:, k] * d.  A record of an operand
//    that is shared by the whole batch applies to every matrix of the batch.  atomicAdd: several records may touch the
//    same element.  Normally there are no records and the kernel returns at once.
//  * the split marked the call ineligible (*gate == gen): write the TF32 lo parts of the raw operands for the gated
//    TF32x3 fallback GEMM that follows (lo0 == nullptr: the fallback is the SIMT kernel, nothing to prepare).
__global__ void __launch_bounds__(256) fp16_post_kernel(float *__restrict__ C, const float *__restrict__ A, const float *__restrict__ B,
                                                        int64_t batch, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                                                        int64_t ldc, int64_t sA, int64_t sB, int64_t sC, FixList fa, FixList fb,
                                                        float *lo0, int64_t n0, float *lo1, int64_t n1,
                                                        int *nonfinite, int gen, unsigned long long *trace) {
    if (threadIdx.x == 0) trace_min(trace, 8);
    // Launched with programmatic stream serialization: these reads only touch what the PRE-PASS wrote (complete before any CTA of
    // the GEMM passed its own griddepcontrol.wait, which precedes its launch_dependents).  In the normal case - eligible, no
    // records - the grid leaves at once, while the GEMM is still running; one thread stays to keep completion ordered.
    const bool fallback = *reinterpret_cast<const volatile int *>(nonfinite + 1) == gen;
    const unsigned int na = min(*reinterpret_cast<volatile unsigned int *>(fa.count), (unsigned int)FIX_CAP),
                       nb = min(*reinterpret_cast<volatile unsigned int *>(fb.count), (unsigned int)FIX_CAP);
    if (!fallback && na + nb == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { pdl_wait(); trace_max(trace, 9); }
        return;
    }
    pdl_wait();   // the GEMM (or, for an ineligible call, its immediate exit) is complete
    if (fallback) {
        if (lo0 != nullptr) split_tf32_body(A, lo0, n0, B, lo1, n1, nonfinite, gen);
        return;
    }
    // Every record is spread over `parts` CTAs (the whole grid works on the few records there are: a B-record walks a COLUMN of C and
    // of A - one cache line per element - and is pure latency; one CTA per record took 16 dependent rounds for M = 4096).
    const unsigned int nrec = na + nb;
    unsigned int parts = gridDim.x / nrec;
    if (parts < 1) parts = 1;
    for (unsigned int w = blockIdx.x; w < nrec * parts; w += gridDim.x) {
        const unsigned int rec = w / parts, part = w - rec * parts;
        const bool from_a = rec < na;
        const int4 q = from_a ? fa.recs[rec] : fb.recs[rec - na];
        const float d = __int_as_float(q.z);
        if (from_a) {
            const int64_t bi = sA ? q.x / M : 0, i = sA ? q.x - bi * M : q.x, k = q.y;
            for (int64_t bb = sA ? bi : 0; bb < (sA ? bi + 1 : batch); bb++) {
                const float *brow = B + (sB ? bb * sB : 0) + k * ldb;
                float *crow = C + bb * sC + i * ldc;
                for (int64_t j = (int64_t)part * 256 + threadIdx.x; j < N; j += (int64_t)parts * 256) atomicAdd(crow + j, d * brow[j]);
            }
        } else {
            const int64_t bk = sB ? q.x / K : 0, k = sB ? q.x - bk * K : q.x, j = q.y;
            for (int64_t bb = sB ? bk : 0; bb < (sB ? bk + 1 : batch); bb++) {
                const float *acol = A + (sA ? bb * sA : 0) + k;
                float *ccol = C + bb * sC + j;
                for (int64_t i = (int64_t)part * 256 + threadIdx.x; i < M; i += (int64_t)parts * 256) atomicAdd(ccol + i * ldc, acol[i * lda] * d);
            }
        }
    }
}

// ------------------------------------------------------------------ SIMT fp32 GEMM (small / unaligned shapes)
// 64x64 tile, BK = 16, 256 threads x (4x4) micro-tile, fp32 FMA, k in increasing order.
__global__ void __launch_bounds__(256) sgemm_simt_kernel(float *__restrict__ C, const float *__restrict__ A,
                                                         const float *__restrict__ B, int64_t M, int64_t N, int64_t K,
                                                         int64_t lda, int64_t ldb, int64_t ldc, int64_t sA, int64_t sB,
                                                         int64_t sC, const int *gate = nullptr, int gate_gen = 0) {
    if (gate != nullptr && *reinterpret_cast<const volatile int *>(gate) != gate_gen) return;   // FP16x3 fallback for operands the TF32 path cannot read
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int64_t b = blockIdx.z;
    A += b * sA; B += b * sB; C += b * sC;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
    float acc[4][4] = {};
    for (int64_t k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            int r = i / 16, c = i % 16;
            As[c][r] = (m0 + r < M && k0 + c < K) ? A[(m0 + r) * lda + k0 + c] : 0.f;
            int rk = i / 64, cn = i % 64;
            Bs[rk][cn] = (k0 + rk < K && n0 + cn < N) ? B[(k0 + rk) * ldb + n0 + cn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; bb[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
            if (r < M && c < N) C[r * ldc + c] = acc[i][j];
        }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map over (inner, rows, batch) of a row-major fp32 matrix stack; box = (32, box_rows, 1), 128B swizzle.
static int make_map(CUtensorMap *map, const void *base, int64_t inner, int64_t rows, int64_t ld, int64_t batch,
                    int64_t batch_stride, int box_rows, CUtensorMapSwizzle swizzle, bool bf16 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error(NB200_ECUDA, "cuTensorMapEncodeTiled unavailable");
    if (batch_stride == 0 || batch < 1) { batch = 1; batch_stride = rows * ld; }
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t esz = bf16 ? 2 : 4;
    cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)batch_stride * esz};
    cuuint32_t box[3] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows, 1};   // one 128-byte swizzle row
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(NB200_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return NB200_OK;
}

struct GemmArgs {
    float *C;
    const float *A, *B, *A_lo, *B_lo;   // A/B are the "hi" (or raw) operands
    int64_t batch, M, N, K, lda, ldb, ldc, sA, sB, sC;
    const unsigned int *row_max = nullptr, *col_max = nullptr;   // FP16x3 scaling inputs (GemmCfg::SCALED)
    int gate_want = -1;   // -1: ungated; 0: run unless the call was marked ineligible for FP16x3; 1: run only if it was
};

static bool pdl_enabled() {
    static const bool on = !(getenv("NB200_PDL") && atoi(getenv("NB200_PDL")) == 0);   // A/B switch, read once
    return on;
}

template <class Cfg>
static int launch_gemm(const GemmArgs &g) {
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    // g.A/A_lo/B/B_lo point at packed bf16 arrays (lda/ldb/sA/sB in bf16 elements) when Cfg::BF16
    const CUtensorMapSwizzle SWA = CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapSwizzle SWB = Cfg::BF16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if ((rc = make_map(&ma_hi, g.A, g.K, g.M, g.lda, g.batch, g.sA, Cfg::BM, SWA, Cfg::BF16)) != NB200_OK) return rc;
    if ((rc = make_map(&mb_hi, g.B, g.N, g.K, g.ldb, g.batch, g.sB, Cfg::BK, SWB, Cfg::BF16)) != NB200_OK) return rc;
    ma_lo = ma_hi;
    mb_lo = mb_hi;
    if (Cfg::PASSES == 3 && !Cfg::INK) {
        if ((rc = make_map(&ma_lo, g.A_lo, g.K, g.M, g.lda, g.batch, g.sA, Cfg::BM, SWA, Cfg::BF16)) != NB200_OK) return rc;
        if ((rc = make_map(&mb_lo, g.B_lo, g.N, g.K, g.ldb, g.batch, g.sB, Cfg::BK, SWB, Cfg::BF16)) != NB200_OK) return rc;
    }
    GemmParams p;
    p.C = g.C; p.M = g.M; p.N = g.N; p.K = g.K; p.ldc = g.ldc; p.strideC = g.sC; p.batch = g.batch;
    p.tiles_m = (int)((g.M + Cfg::BM * Cfg::CG - 1) / (Cfg::BM * Cfg::CG));
    p.tiles_n = (int)((g.N + Cfg::BN - 1) / Cfg::BN);
    p.total_tiles = (int64_t)p.tiles_m * p.tiles_n * g.batch;
    // measured on B200 (BF16x3): 4096^3 0.308 -> 0.300 ms, 8192^3 2.40 -> 2.22 ms against the plain m-fastest order
    p.full_items = p.total_tiles;
    if (Cfg::MERGED) {
        // Last wave poorly filled (at most half of the CTA pairs busy): split its tiles into two 128-column halves each, so
        // that wave keeps (nearly) every pair busy for half a tile time.  4096^3: 256 tiles on 74 pairs = 3.46 -> 3.5 waves
        // instead of 4.
        const int64_t pairs = ctx().num_sms / Cfg::CG;
        const int64_t rem = p.total_tiles % pairs;
        static const bool split_tail = getenv("NB200_GEMM_NO_SPLIT_TAIL") == nullptr;
        if (split_tail && rem > 0 && 2 * rem <= pairs && g.N % Cfg::BN == 0) {   // (ragged last n-tile: keep whole tiles)
            p.full_items = p.total_tiles - rem;
            p.total_tiles = p.full_items + 2 * rem;
        }
    }
    static const int group_m = getenv("NB200_GEMM_GROUP_M") ? atoi(getenv("NB200_GEMM_GROUP_M")) : 8;
    p.group_m = group_m;
    static const int early = getenv("NB200_GEMM_EARLY_CROSS") ? atoi(getenv("NB200_GEMM_EARLY_CROSS")) : 1;
    p.early_cross = early;
    p.row_max = g.row_max; p.col_max = g.col_max;
    p.gate = g.gate_want >= 0 ? nonfinite_flag() + 1 : nullptr;
    p.gate_gen = ctx().nonfinite_gen;
    p.gate_want = g.gate_want > 0 ? 1 : 0;
    p.nonfinite = nonfinite_flag();
    p.nonfinite_gen = ctx().nonfinite_gen;
    p.a_batched = g.sA != 0;
    p.b_batched = g.sB != 0;
    // pinned host memory (device-visible under UVA): survives a trap so the host can report which wait timed out
    p.debug = reinterpret_cast<unsigned int *>(ctx().host_result) + 4;
    p.trace = ctx().trace;
    auto kern = sgemm_tf32_kernel<Cfg>;
    // the opt-in shared-memory size is a per-device function attribute (nb200_set_device may move the context)
    static bool attr_set[64] = {};
    const int dev = ctx().device;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        NB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    int64_t clusters = (ctx().num_sms - ctx().gemm_sm_reserve) / Cfg::CG;   // persistent: one CTA (pair) per SM, minus the SMs left to concurrent transfer kernels
    if (clusters < 1) clusters = 1;
    if (clusters > p.total_tiles) clusters = p.total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * Cfg::CG));
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the prologue overlaps the tail of the pre-pass (pdl_wait in the kernel)
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    NB_CUDA(cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, p));
    ctx().launches++;
    return NB200_OK;
}

static int launch_split(const float *in0, float *lo0, int64_t n0, const float *in1, float *lo1, int64_t n1) {
    int64_t groups = ((n0 + 3) >> 2) + ((n1 + 3) >> 2);
    if (groups == 0) return NB200_OK;
    int64_t blocks = (groups + 255) / 256;   // one 4-element group per thread, non-persistent (see common.cuh)
    if (blocks > 0x7FFFFFFF) blocks = 0x7FFFFFFF;
    split_tf32_kernel<<<(unsigned)blocks, 256, 0, ctx().stream>>>(in0, lo0, n0, in1, lo1, n1, nonfinite_flag(), ctx().nonfinite_gen);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

static inline int64_t span(int64_t batch, int64_t stride, int64_t rows, int64_t ld, int64_t cols) {
    int64_t one = (rows - 1) * ld + cols;
    return stride == 0 ? one : (batch - 1) * stride + one;
}
static inline int64_t round4(int64_t x) { return (x + 3) & ~int64_t(3); }

// variant: 0 = auto; otherwise (CG<<8 | BN) for experiments (NB200_GEMM_VARIANT env)
static int g_variant = -1;
static int gemm_variant() {
    if (g_variant < 0) {
        const char *e = getenv("NB200_GEMM_VARIANT");
        g_variant = e ? atoi(e) : 0;
    }
    return g_variant;
}

// Default tiles: CTA pairs (cta_group::2) whenever there is more than one 128-row block;
// TF32x1 256x256 per pair, TF32x3 256x128 per pair (chunked accumulation needs 4 x BN TMEM columns).
template <int PASSES>
static int dispatch_cfg(const GemmArgs &g) {
    int v = gemm_variant();
    int cg = v ? (v >> 8) : (g.M > 128 ? 2 : 1);
    int bn = v ? (v & 0xFF) * 2 : (PASSES == 3 ? 128 : (g.N > 128 ? 256 : 128));   // encoded as BN/2
    // (Measured and rejected: single-CTA 128-row tiles when a call has fewer pair tiles than half the CTA pairs, e.g. the 256-row
    // blocks of nb200_sgemm_host - 32 pair tiles for 74 pairs.  Twice the busy SMs, but without the pair's operand multicast every
    // CTA reads its own B panel: a block's product took 110-125 us instead of 90 and the whole call 3.02 instead of 2.78 ms.)
    if constexpr (PASSES == 3) {
        if (g.A_lo == nullptr) {   // in-kernel split (default): no lo arrays exist
            if (cg == 2) return launch_gemm<GemmCfg<2, 128, 3, true>>(g);
            return launch_gemm<GemmCfg<1, 128, 3, true>>(g);
        }
        if (cg == 2) return launch_gemm<GemmCfg<2, 128, 3>>(g);
        return launch_gemm<GemmCfg<1, 128, 3>>(g);
    } else {
        if (cg == 2 && bn == 256) return launch_gemm<GemmCfg<2, 256, 1>>(g);
        if (cg == 2 && bn == 128) return launch_gemm<GemmCfg<2, 128, 1>>(g);
        if (cg == 1 && bn == 128) return launch_gemm<GemmCfg<1, 128, 1>>(g);
        return launch_gemm<GemmCfg<1, 256, 1>>(g);
    }
}

static bool tensor_path_ok(const GemmArgs &g) {
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return al(g.A) && al(g.B) && (g.lda % 4 == 0) && (g.ldb % 4 == 0) && (g.sA % 4 == 0) && (g.sB % 4 == 0) &&
           g.M * g.N * g.K >= (int64_t)64 * 64 * 64 && g.K >= 32 && g.N >= 32;
}

static inline int64_t round8(int64_t x) { return (x + 7) & ~int64_t(7); }
// Batched x3 calls process the batch in chunks whose operand parts fit this much workspace (default 16 GiB of the 180 GB;
// NB200_GEMM_WS_BUDGET_MB overrides it, read per call so that tests can force several chunks on small problems).
static int64_t gemm_ws_budget() {
    const char *e = getenv("NB200_GEMM_WS_BUDGET_MB");
    const int64_t mb = e ? atoll(e) : 0;
    return mb > 0 ? mb << 20 : (int64_t)16 << 30;
}
static SplitSpan make_span(const float *in, __nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t batch, int64_t rows, int64_t cols,
                           int64_t ld_in, int64_t stride_in) {
    SplitSpan s;
    s.in = in; s.hi = hi; s.lo = lo;
    s.rows_per = rows; s.nrows = batch * rows; s.cols = cols; s.ld_in = ld_in; s.stride_in = stride_in;
    s.ld_out = round8(cols);
    s.gpr = s.ld_out >> 2;
    s.groups = s.nrows * s.gpr;
    s.vec = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (ld_in % 4 == 0) && (stride_in % 4 == 0);
    s.flat = s.vec && cols % 8 == 0 && ld_in == cols && (batch <= 1 || stride_in == rows * cols);
    return s;
}
static int launch_split_bf16(const SplitSpan &s0, const SplitSpan &s1) {
    const int64_t groups = s0.groups + s1.groups;
    if (groups == 0) return NB200_OK;
    if ((s0.flat || s0.groups == 0) && (s1.flat || s1.groups == 0)) {
        const int64_t g0 = s0.groups >> 1, g1 = s1.groups >> 1;   // 8-element groups
        int64_t blocks = (g0 + g1 + 255) / 256;
        if (blocks > 0x7FFFFFFF) blocks = 0x7FFFFFFF;
        split_bf16_flat_kernel<<<(unsigned)blocks, 256, 0, ctx().stream>>>(s0.in, s0.hi, s0.lo, g0, s1.in, s1.hi, s1.lo, g1, nonfinite_flag(), ctx().nonfinite_gen);
        NB_LAUNCH_CHECK();
        return NB200_OK;
    }
    int64_t blocks = (groups + 255) / 256;
    if (blocks > 0x7FFFFFFF) blocks = 0x7FFFFFFF;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, ctx().stream>>>(s0, s1, nonfinite_flag(), ctx().nonfinite_gen);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

// BF16x3 tile choice for CTA pairs.  Measured on B200: the merged 256x256 tile is ~7.5 % faster per flop than 256x128
// (8192^3: 2.06 vs 2.21 ms) but has half as many tiles, so it loses when the last wave over the 74 CTA pairs is poorly
// filled (4096^3: 256 tiles = 3.46 waves, 0.308 vs 0.297 ms).  Pick by wave-quantisation efficiency.
static int bf16_pair_bn(int64_t batch, int64_t M, int64_t N) {
    if (N <= 128) return 128;
    const int64_t clusters = ctx().num_sms / 2 > 0 ? ctx().num_sms / 2 : 1;
    auto eff = [&](int bn) {
        const int64_t tiles = batch * ((M + 255) / 256) * ((N + bn - 1) / bn);
        const int64_t rem = tiles % clusters;
        double waves = (double)(tiles / clusters);
        if (rem) waves += (bn == 256 && 2 * rem <= clusters && N % 256 == 0) ? 0.5 : 1.0;   // the merged kernel splits a thin last wave
        return (double)tiles / (waves * (double)clusters);
    };
    return eff(256) * 1.07 > eff(128) ? 256 : 128;
}

// BF16x3: C = a1.b1 + (a2.b1 + a1.b2) with bf16 pairs (a1, a2), (b1, b2) from the pre-pass, same chunked kernel.
// Dropped terms (a2.b2 and the two split remainders) are each <= 2^-18 |a||b| per product and zero-mean.
static int gemm_bf16x3(const GemmArgs &g) {
    const int64_t lda = round8(g.K), ldb = round8(g.N);
    const int64_t per_a = g.M * lda, per_b = g.K * ldb;          // bf16 elements per matrix
    int64_t chunk = g.batch;
    const int64_t budget = gemm_ws_budget();
    if (g.batch > 1) {
        const int64_t per = 4 * ((g.sA ? per_a : 0) + (g.sB ? per_b : 0));
        if (per > 0 && per * chunk > budget) chunk = budget / per;
        if (chunk < 1) chunk = 1;
    }
    { const int rcf = gemm_reset_nonfinite(); if (rcf != NB200_OK) return rcf; }
    for (int64_t b0 = 0; b0 < g.batch; b0 += chunk) {
        const int64_t nb = g.batch - b0 < chunk ? g.batch - b0 : chunk;
        // workspace offsets follow the FULL chunk size: a shared operand, split with the first chunk, must stay where it is
        // when the last chunk is shorter
        const int64_t na = (g.sA ? chunk : 1) * per_a, nbb = (g.sB ? chunk : 1) * per_b;
        int rc = ensure_gemm_ws((na + nbb) * 4 + 1024);
        if (rc != NB200_OK) return rc;
        __nv_bfloat16 *ws = static_cast<__nv_bfloat16 *>(ctx().gemm_ws);
        __nv_bfloat16 *a_hi = ws, *a_lo = ws + na, *b_hi = ws + 2 * na, *b_lo = ws + 2 * na + nbb;
        const float *a_src = g.A + (g.sA ? b0 * g.sA : 0), *b_src = g.B + (g.sB ? b0 * g.sB : 0);
        const bool do_a = (b0 == 0 || g.sA), do_b = (b0 == 0 || g.sB);   // a shared operand is split once
        SplitSpan sa = make_span(a_src, a_hi, a_lo, do_a ? (g.sA ? nb : 1) : 0, g.M, g.K, g.lda, g.sA);
        SplitSpan sb = make_span(b_src, b_hi, b_lo, do_b ? (g.sB ? nb : 1) : 0, g.K, g.N, g.ldb, g.sB);
        if ((rc = launch_split_bf16(sa, sb)) != NB200_OK) return rc;
        GemmArgs c = g;
        c.batch = nb;
        c.A = reinterpret_cast<const float *>(a_hi); c.A_lo = reinterpret_cast<const float *>(a_lo);
        c.B = reinterpret_cast<const float *>(b_hi); c.B_lo = reinterpret_cast<const float *>(b_lo);
        c.lda = lda; c.ldb = ldb;
        c.sA = g.sA ? per_a : 0; c.sB = g.sB ? per_b : 0;
        c.C = g.C + b0 * g.sC;
        const int v = gemm_variant();
        const int cg = v ? (v >> 8) : (g.M > 128 ? 2 : 1);
        // NB200_GEMM_VARIANT: CG << 8 | BN / 2; default = CTA pairs with merged accumulation (BN = 256) once N > 128
        const int bn = v ? (v & 0xFF) * 2 : (cg == 2 ? bf16_pair_bn(nb, g.M, g.N) : 128);
        if (cg == 2 && bn == 256) rc = launch_gemm<GemmCfg<2, 256, 3, false, true, true>>(c);
        else rc = cg == 2 ? launch_gemm<GemmCfg<2, 128, 3, false, true>>(c) : launch_gemm<GemmCfg<1, 128, 3, false, true>>(c);
        if (rc != NB200_OK) return rc;
    }
    return NB200_OK;
}

// FP16x3: the BF16x3 pipeline with IEEE-half hi parts of row-scaled A / column-scaled B (GemmCfg::SCALED): 11-bit parts give
// the TF32x3 error class at the kind::f16 rate - provided every non-zero element lies within 2^-28 of its row / column
// maximum (split_f16).  That is decided ON THE DEVICE by the split pre-pass; the host enqueues both the gated FP16x3 GEMM
// (+ the sparse repair of the few elements outside the window) and the gated fallback, exactly one of which does the work.
// Tiles per CTA pair: 256x128 with the separate cross-term accumulator (default), or 256x256 with one merged accumulator per
// k-block whose cross products are folded with scale-input-d (GemmCfg; slower, see fp16_pair_bn).
// The pre-pass repacks the operands, so there is no alignment / leading-dimension requirement; the fallback is TF32x3 on the
// raw operands when the TF32 path can read them (tensor_path_ok) and the fp32 SIMT kernel otherwise.
// Workspace: [a_hi | a_lo | b_hi | b_lo (16-bit)] [fix counters (16 B) | col_max] [row_max] [a_lo32 | b_lo32 (fp32 lo parts
// of the TF32x3 fallback)] [fix records].
// Measured on B200 (profiles/r2_summary.md): the merged SCALED tile is correct (max rel. error 7e-7, the K = 64 chunks shorten
// the truncating accumulation) but SLOWER: draining a 256-column accumulator after every k-block reads 128 KB of TMEM per CTA
// per 12 MMAs, and tcgen05.ld moves 64 B/clk per SM (2048 clk against 1536 clk of MMA time): 4096^3 GEMM 398 us vs 265 us with the
// 256x128 tile.  Chunks of two k-blocks would need four 64 KB stages resident.  The 256x128 tile is therefore the default;
// NB200_FP16_TILE=256 selects the merged one (tests, experiments).
static int fp16_pair_bn(int64_t batch, int64_t M, int64_t N) {
    (void)batch; (void)M;
    const char *e = getenv("NB200_FP16_TILE");   // read per call
    return (e && atoi(e) == 256 && N > 128) ? 256 : 128;
}
static int launch_fp16_prepass(const float *a_src, const float *b_src, const GemmArgs &g, int64_t ba, int64_t bb, bool do_a, bool do_b,
                               SplitSpanF16 &sa, SplitSpanF16 &sb, unsigned int *row_max, unsigned int *col_max, unsigned int *barrier,
                               unsigned int *zero_ptr, int64_t zero_words, bool *zeroed_other, bool mix, int max_ctas_per_sm = 0) {
    *zeroed_other = false;
    const int64_t n_rows = ba * g.M;
    if (sa.s.groups + sb.s.groups == 0) return NB200_OK;
    const bool flat = (sa.s.flat || sa.s.groups == 0) && (sb.s.flat || sb.s.groups == 0);
    if (flat) {
        // one persistent launch, all CTAs co-resident (prep16_coop_kernel has a grid barrier between B's two phases)
        static int per_sm[64] = {};
        const int dev = ctx().device;
        int occ = (dev >= 0 && dev < 64) ? per_sm[dev] : 0;
        if (occ == 0) {
            int occ0 = 0, occ1 = 0;   // (both instantiations: the grid must be co-resident whichever runs)
            NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, prep16_coop_kernel<false>, 256, 0));
            NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, prep16_coop_kernel<true>, 256, 0));
            occ = occ0 < occ1 ? occ0 : occ1;
            if (occ < 1) return set_error(NB200_ECUDA, "sgemm: the FP16x3 pre-pass kernel does not fit on an SM");
            if (dev >= 0 && dev < 64) per_sm[dev] = occ;
        }
        PrepCoop q;
        q.a = sa; q.b = sb;
        q.a_rows = do_a ? n_rows : 0;
        q.b_mats = do_b ? bb : 0;
        q.row_max_out = row_max;
        q.barrier = barrier;
        q.zero_ptr = zero_ptr; q.zero_words = zero_words;
        q.debug = reinterpret_cast<unsigned int *>(ctx().host_result) + 4;
        q.trace = ctx().trace;
        // grid: every CTA resident; small problems get fewer CTAs (a CTA pass covers 256 / TPR rows of A or ~32 K elements of B)
        const int64_t n8 = g.K >> 3;
        const int64_t rpc = n8 <= 64 ? 8 : n8 <= 128 ? 4 : n8 <= 256 ? 2 : 1;
        const int64_t a_elems = q.a_rows * g.K, b_elems = q.b_mats * g.K * g.N;
        const int64_t a_ctas = (q.a_rows + rpc - 1) / rpc, b_ctas = (b_elems + 32767) / 32768;
        // (max_ctas_per_sm: the pipelined batched path runs this kernel BESIDE a persistent GEMM grid that leaves room for one of
        // these CTAs per SM; the grid barrier needs every CTA resident)
        int64_t grid = (int64_t)ctx().num_sms * (max_ctas_per_sm > 0 && max_ctas_per_sm < occ ? max_ctas_per_sm : occ);
        if (grid > a_ctas + b_ctas) grid = a_ctas + b_ctas;
        if (grid < 2) grid = 2;
        // CTAs per operand in proportion to the bytes each side moves through L2: A is read once and written once (8 B per element),
        // B is read twice and written once (12 B per element; NB200_PREP_BW overrides the B weight in percent of A's).  Measured with
        // the software-pipelined A rows (4096^2, pre-pass end): 100 -> 63.8 us, 150 -> 57.3, 200 -> 50.3, 250 -> 55.6.
        static const int64_t bw = getenv("NB200_PREP_BW") ? atoll(getenv("NB200_PREP_BW")) : 200;
        const int64_t wa = a_elems * 100, wb = b_elems * bw;
        int64_t n_b = a_elems == 0 ? grid : b_elems == 0 ? 0 : (int64_t)(((double)grid * (double)wb) / ((double)wa + (double)wb) + 0.5);
        if (a_elems > 0 && b_elems > 0) { if (n_b < 1) n_b = 1; if (n_b > grid - 1) n_b = grid - 1; }
        q.n_b = (int)n_b;
        cudaLaunchConfig_t pc = {};
        pc.gridDim = dim3((unsigned)grid);
        pc.blockDim = dim3(256);
        pc.stream = ctx().stream;
        cudaLaunchAttribute pa[1];
        pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // back-to-back calls: the launch latency hides behind the previous call's tail
        pa[0].val.programmaticStreamSerializationAllowed = 1;
        pc.attrs = pa;
        pc.numAttrs = pdl_enabled() ? 1 : 0;
        if (mix) NB_CUDA(cudaLaunchKernelEx(&pc, prep16_coop_kernel<true>, q, nonfinite_flag(), ctx().nonfinite_gen));
        else NB_CUDA(cudaLaunchKernelEx(&pc, prep16_coop_kernel<false>, q, nonfinite_flag(), ctx().nonfinite_gen));
        ctx().launches++;
        *zeroed_other = zero_words > 0;
        return NB200_OK;
    }
    // general operands (views, odd leading dimensions): separate |max| launches, then the repacking split
    if (do_a) {
        int64_t blocks = (n_rows + 7) / 8;
        if (blocks > 0x7FFFFFFF) blocks = 0x7FFFFFFF;
        absmax_rows_kernel<<<(unsigned)blocks, 256, 0, ctx().stream>>>(a_src, g.M, n_rows, g.K, g.lda, g.sA, row_max);
        NB_LAUNCH_CHECK();
    }
    if (do_b) {
        const bool vec4 = (g.N % 4 == 0) && sb.s.vec;
        const int64_t bx = vec4 ? (g.N / 4 + 255) / 256 : (g.N + 255) / 256;
        int64_t by = (2 * ctx().num_sms + bx * bb - 1) / (bx * bb);   // row segments: ~2 blocks per SM in total
        if (by < 1) by = 1;
        if (by > (g.K + 7) / 8) by = (g.K + 7) / 8;
        if (by > 65535) by = 65535;
        if (vec4)
            absmax_cols4_kernel<<<dim3((unsigned)bx, (unsigned)by, (unsigned)bb), 256, 0, ctx().stream>>>(b_src, g.K, g.N, g.ldb, g.sB, col_max);
        else
            absmax_cols_kernel<<<dim3((unsigned)bx, (unsigned)by, (unsigned)bb), 256, 0, ctx().stream>>>(b_src, g.K, g.N, g.ldb, g.sB, col_max);
        NB_LAUNCH_CHECK();
    }
    int64_t blocks = (sa.s.groups + sb.s.groups + 255) / 256;
    if (blocks > 0x7FFFFFFF) blocks = 0x7FFFFFFF;
    if (mix) split_f16_kernel<true><<<(unsigned)blocks, 256, 0, ctx().stream>>>(sa, sb, nonfinite_flag(), ctx().nonfinite_gen);
    else split_f16_kernel<false><<<(unsigned)blocks, 256, 0, ctx().stream>>>(sa, sb, nonfinite_flag(), ctx().nonfinite_gen);
    NB_LAUNCH_CHECK();
    return NB200_OK;
}

// ---- Tried and removed (measured on B200, profiles/r2_summary.md): software-pipelining a batched call over two streams and two
// workspace sets so that the pre-pass of chunk c+1 runs BESIDE the GEMM of chunk c.  Beside the 192-thread GEMM CTA (~220
// registers per thread, ~200 KB shared memory) an SM has room for ONE 256-thread pre-pass CTA, and at a quarter of its usual
// occupancy the latency-bound pre-pass of a chunk takes longer than the chunk's GEMM: the pipeline becomes pre-pass bound
// (128 x 2048^2 on one GPU: 6.9 -> 10.2 ms).  The merged 256x256 GEMM owns the whole register file, nothing co-resides with it.
// Hiding the pre-pass needs a pre-pass that keeps ~40 KB per SM in flight from one small CTA (TMA-staged through the ~26 KB of
// shared memory the GEMM leaves free) - future work.
template <bool MIX>
static int launch_fp16_gemm(const GemmArgs &c, int cg, bool merged) {
    if (MIX) {
        if (merged) return launch_gemm<GemmCfg<2, 256, 3, false, true, true, true, true>>(c);
        return cg == 2 ? launch_gemm<GemmCfg<2, 128, 3, false, true, false, true, true>>(c) : launch_gemm<GemmCfg<1, 128, 3, false, true, false, true, true>>(c);
    }
    if (merged) return launch_gemm<GemmCfg<2, 256, 3, false, true, true, true>>(c);
    return cg == 2 ? launch_gemm<GemmCfg<2, 128, 3, false, true, false, true>>(c) : launch_gemm<GemmCfg<1, 128, 3, false, true, false, true>>(c);
}

static int gemm_fp16x3(const GemmArgs &g, bool mix) {
    const bool raw_ok = tensor_path_ok(g);                       // the TF32x3 fallback can read the raw operands
    const int64_t lda = round8(g.K), ldb = round8(g.N);
    const int64_t per_a = g.M * lda, per_b = g.K * ldb;          // 16-bit elements per matrix
    int64_t chunk = g.batch;
    const int64_t budget = gemm_ws_budget();
    if (g.batch > 1) {
        const int64_t per = 4 * ((g.sA ? per_a : 0) + (g.sB ? per_b : 0)) + (raw_ok ? 4 * ((g.sA ? round4(g.sA) : 0) + (g.sB ? round4(g.sB) : 0)) : 0);
        if (per > 0 && per * chunk > budget) chunk = budget / per;
        if (chunk < 1) chunk = 1;
        if (chunk > 65535) chunk = 65535;                        // absmax_cols_kernel / the SIMT fallback put the batch on gridDim.z
    }
    { const int rcf = gemm_reset_nonfinite(); if (rcf != NB200_OK) return rcf; }
    const int v = gemm_variant();
    const int cg = v ? (v >> 8) : (g.M > 128 ? 2 : 1);
    // FP16x3U (mix): one accumulator per chunk, so the merged 256x256 tile is chosen exactly as for BF16x3 (wave quantisation)
    const int bn = v ? (v & 0xFF) * 2 : (cg == 2 ? (mix ? bf16_pair_bn(chunk, g.M, g.N) : fp16_pair_bn(chunk, g.M, g.N)) : 128);
    const bool merged = cg == 2 && bn == 256;
    // workspace offsets follow the FULL chunk size (a shared operand prepared with the first chunk must not move)
    const int64_t na = (g.sA ? chunk : 1) * per_a, nbb = (g.sB ? chunk : 1) * per_b;
    const int64_t rows_layout = round4((g.sA ? chunk : 1) * g.M), cols_layout = round4((g.sB ? chunk : 1) * g.N);   // 16-byte aligned sub-arrays
    const int64_t na32 = raw_ok ? round4(span(g.sA ? chunk : 1, g.sA, g.M, g.lda, g.K)) : 0, nb32 = raw_ok ? round4(span(g.sB ? chunk : 1, g.sB, g.K, g.ldb, g.N)) : 0;
    // Control blocks (16-byte slot: [0] A records, [1] B records, [2] pre-pass barrier; then col_max) live at the start of the
    // workspace, double-buffered by call parity.  A single-chunk call uses block `parity` and its pre-pass zeroes the other one for
    // the next call (Ctx::ctl_ready), so steady-state calls enqueue no memset; multi-chunk calls use block 0 with explicit memsets.
    const int64_t ctl_stride = (16 + cols_layout * 4 + 255) & ~int64_t(255);
    // (inside a stream capture the recorded call is replayed many times: it must carry its own memset and leave the cross-call state alone)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NB_CUDA(cudaStreamIsCapturing(ctx().stream, &cap));
    const bool single = chunk >= g.batch && cap == cudaStreamCaptureStatusNone;
    for (int64_t b0 = 0; b0 < g.batch; b0 += chunk) {
        const int64_t nb = g.batch - b0 < chunk ? g.batch - b0 : chunk;
        const int64_t ba = g.sA ? nb : 1, bb = g.sB ? nb : 1;   // matrices of each operand in this chunk
        const int64_t fix_bytes = 2 * (int64_t)FIX_CAP * (int64_t)sizeof(int4);
        Ctx &cx = ctx();
        const void *ws_before = cx.gemm_ws;
        const int64_t stride_before = cx.ctl_stride;
        int rc = ensure_gemm_ws(2 * ctl_stride + (na + nbb) * 4 + rows_layout * 4 + (na32 + nb32) * 4 + fix_bytes + 1024);
        if (rc != NB200_OK) return rc;
        if (cx.gemm_ws != ws_before || stride_before != ctl_stride || !single) cx.ctl_ready[0] = cx.ctl_ready[1] = 0;   // unknown contents
        cx.ctl_stride = ctl_stride;
        const int mine = single ? cx.ctl_parity : 0;
        char *wsb = static_cast<char *>(cx.gemm_ws);
        unsigned int *fix_cnt = reinterpret_cast<unsigned int *>(wsb + mine * ctl_stride);
        unsigned int *other_ctl = reinterpret_cast<unsigned int *>(wsb + (1 - mine) * ctl_stride);
        unsigned int *col_max = fix_cnt + 4;
        __nv_bfloat16 *ws = reinterpret_cast<__nv_bfloat16 *>(wsb + 2 * ctl_stride);   // (16-bit storage; the contents are IEEE half)
        __nv_bfloat16 *a_hi = ws, *a_lo = ws + na, *b_hi = ws + 2 * na, *b_lo = ws + 2 * na + nbb;
        unsigned int *row_max = reinterpret_cast<unsigned int *>(ws + 2 * na + 2 * nbb);
        float *a_lo32 = reinterpret_cast<float *>(row_max + rows_layout);
        float *b_lo32 = a_lo32 + na32;
        int4 *recs = reinterpret_cast<int4 *>(b_lo32 + nb32);
        FixList fix_a{fix_cnt, recs}, fix_b{fix_cnt + 1, recs + FIX_CAP};
        const float *a_src = g.A + (g.sA ? b0 * g.sA : 0), *b_src = g.B + (g.sB ? b0 * g.sB : 0);
        const bool do_a = (b0 == 0 || g.sA), do_b = (b0 == 0 || g.sB);   // a shared operand is prepared once (its records persist)
        const int64_t need = 16 + bb * g.N * 4;                          // counters + barrier + the column maxima (atomicMax targets)
        if (single) {
            if (cx.ctl_ready[mine] < need) NB_CUDA(cudaMemsetAsync(fix_cnt, 0, (size_t)need, cx.stream));
            cx.ctl_ready[mine] = 0;                                      // in use from here on
        } else if (do_b) {
            NB_CUDA(cudaMemsetAsync(do_a ? fix_cnt : fix_cnt + 1, 0, (size_t)(do_a ? 16 : 12) + (size_t)(bb * g.N) * 4, cx.stream));
        } else if (do_a) {
            NB_CUDA(cudaMemsetAsync(fix_cnt, 0, 4, cx.stream));
        }
        SplitSpanF16 sa, sb;
        sa.s = make_span(a_src, a_hi, a_lo, do_a ? ba : 0, g.M, g.K, g.lda, g.sA);
        sa.max_bits = row_max; sa.by_col = 0; sa.fix = fix_a;
        sb.s = make_span(b_src, b_hi, b_lo, do_b ? bb : 0, g.K, g.N, g.ldb, g.sB);
        sb.max_bits = col_max; sb.by_col = 1; sb.fix = fix_b;
        bool zeroed_other = false;
        if ((rc = launch_fp16_prepass(a_src, b_src, g, ba, bb, do_a, do_b, sa, sb, row_max, col_max, fix_cnt + 2, other_ctl, single ? need / 4 : 0,
                                      &zeroed_other, mix)) != NB200_OK) return rc;
        if (single) {
            if (zeroed_other) cx.ctl_ready[1 - mine] = need;
            cx.ctl_parity = 1 - mine;
        }
        // (1) FP16x3 product, runs unless the split marked the call ineligible
        GemmArgs c = g;
        c.batch = nb;
        c.A = reinterpret_cast<const float *>(a_hi); c.A_lo = reinterpret_cast<const float *>(a_lo);
        c.B = reinterpret_cast<const float *>(b_hi); c.B_lo = reinterpret_cast<const float *>(b_lo);
        c.lda = lda; c.ldb = ldb;
        c.sA = g.sA ? per_a : 0; c.sB = g.sB ? per_b : 0;
        c.C = g.C + b0 * g.sC;
        c.row_max = row_max; c.col_max = col_max;
        c.gate_want = 0;
        rc = mix ? launch_fp16_gemm<true>(c, cg, merged) : launch_fp16_gemm<false>(c, cg, merged);
        if (rc != NB200_OK) return rc;
        // (2) eligible: sparse repair of the recorded out-of-window elements (normally none: returns at once);
        //     ineligible: TF32 lo parts of BOTH raw operands of this chunk (a shared operand is split again with every chunk:
        //     an earlier chunk's fallback split may never have run)
        const int64_t s_a = span(ba, g.sA, g.M, g.lda, g.K), s_b = span(bb, g.sB, g.K, g.ldb, g.N);
        {
            cudaLaunchConfig_t pc = {};
            pc.gridDim = dim3((unsigned)(ctx().num_sms * 4));
            pc.blockDim = dim3(256);
            pc.stream = ctx().stream;
            cudaLaunchAttribute pa[1];
            pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pa[0].val.programmaticStreamSerializationAllowed = 1;
            pc.attrs = pa;
            pc.numAttrs = pdl_enabled() ? 1 : 0;
            NB_CUDA(cudaLaunchKernelEx(&pc, fp16_post_kernel, g.C + b0 * g.sC, a_src, b_src, nb, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.sA, g.sB, g.sC,
                                       fix_a, fix_b, raw_ok ? a_lo32 : (float *)nullptr, s_a, raw_ok ? b_lo32 : (float *)nullptr, s_b, nonfinite_flag(),
                                       ctx().nonfinite_gen, ctx().trace));
            ctx().launches++;
        }
        // (3) the gated fallback, runs only if the call was marked: TF32x3 on the raw operands, bit-identical to a TF32X3 call
        if (raw_ok) {
            GemmArgs f = g;
            f.batch = nb;
            f.A = a_src; f.A_lo = a_lo32; f.B = b_src; f.B_lo = b_lo32;
            f.C = g.C + b0 * g.sC;
            f.gate_want = 1;
            if ((rc = dispatch_cfg<3>(f)) != NB200_OK) return rc;
        } else {
            dim3 grid((unsigned)((g.N + 63) / 64), (unsigned)((g.M + 63) / 64), (unsigned)nb);
            sgemm_simt_kernel<<<grid, 256, 0, ctx().stream>>>(g.C + b0 * g.sC, a_src, b_src, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.sA, g.sB, g.sC,
                                                              nonfinite_flag() + 1, ctx().nonfinite_gen);
            NB_LAUNCH_CHECK();
        }
    }
    return NB200_OK;
}

// AUTO = the fastest mode whose error bound GUARANTEES 1e-5 against cblas_sgemm for every input: FP16x3U - half hi parts of
// row-scaled A / column-scaled B plus UNSCALED half lo parts, all three products of a 256-long k-chunk in one TMEM accumulator,
// merged 256x256 tile (per product <= 2^-18 + 2^-22 at the edge of its 2^-21 window, 3 * 2^-22 near the row / column maximum;
// measured ~2e-6), out-of-window elements repaired / gated TF32x3 fallback decided on the device.  Measured on B200 against FP16x3
// (2^11-scaled lo parts, 256x128 tile, 2^-28 window): 4096^3 0.334 vs 0.338 ms, 8192^3 2.36 vs 2.55 ms, 2048^3 0.074 vs 0.096 ms.
// TF32x3 has a tighter bound at half the tensor rate and is what AUTO uses for K < 128 (pre-pass not worth it) and what the
// fallback runs.  BF16x3 is faster still but its bound is only statistical: each product may be off by up to 2^-16 + 2 * 2^-17
// (dropped a2.b2 and the split remainders) - zero-mean, so it averages out over K for ordinary data (measured 1.2-2.5e-6) but adds
// up coherently for e.g. constant matrices (tests/test_gemm_split_model.py: 4.7 % of random constant pairs exceed 1e-5).  It has
// to be asked for: precision = NB200_GEMM_BF16X3 per call.
// NB200_GEMM_AUTO_MODE = tf32x3 | bf16x3 | fp16x3 | fp16x3u overrides what AUTO stands for at K >= 128.
int gemm_resolve_precision(int precision, int64_t K) {
    if (precision != NB200_GEMM_AUTO) return precision;
    static const char *mode = getenv("NB200_GEMM_AUTO_MODE");
    static const int fast = !mode ? NB200_GEMM_FP16X3U : strcmp(mode, "bf16x3") == 0 ? NB200_GEMM_BF16X3
                                  : strcmp(mode, "tf32x3") == 0 ? NB200_GEMM_TF32X3
                                  : strcmp(mode, "fp16x3") == 0 ? NB200_GEMM_FP16X3 : NB200_GEMM_FP16X3U;
    return K >= 128 ? fast : NB200_GEMM_TF32X3;
}

static int gemm_impl(GemmArgs g, int precision) {
    if (g.batch == 0 || g.M == 0 || g.N == 0) return NB200_OK;
    if (g.K == 0) {
        for (int64_t b = 0; b < g.batch; b++)
            for (int64_t r = 0; r < g.M; r++)
                NB_CUDA(cudaMemsetAsync(g.C + b * g.sC + r * g.ldc, 0, (size_t)g.N * 4, ctx().stream));
        return NB200_OK;
    }
    precision = gemm_resolve_precision(precision, g.K);
    // BF16x3 repacks its operands, so it has no alignment / leading-dimension requirements of its own
    const bool bf16_ok = g.M * g.N * g.K >= (int64_t)64 * 64 * 64 && g.K >= 32 && g.N >= 32;
    static const bool force_simt = getenv("NB200_GEMM_FORCE_SIMT") != nullptr;   // debugging switch, read once
    if (precision == NB200_GEMM_BF16X3 && bf16_ok && !force_simt) return gemm_bf16x3(g);
    if (precision == NB200_GEMM_FP16X3 && bf16_ok && !force_simt) return gemm_fp16x3(g, false);
    if (precision == NB200_GEMM_FP16X3U && bf16_ok && !force_simt) return gemm_fp16x3(g, true);
    // TF32X3 asked for on operands the TF32 path cannot read (4-byte aligned views, ld % 4 != 0): the FP16x3 pre-pass
    // repacks them, same class of guaranteed bound, so they stay on the tensor pipe instead of the fp32 SIMT kernel
    if (precision == NB200_GEMM_TF32X3 && bf16_ok && !tensor_path_ok(g) && g.K >= 128 && !force_simt) return gemm_fp16x3(g, true);
    if (precision == NB200_GEMM_BF16X3 || precision == NB200_GEMM_FP16X3 || precision == NB200_GEMM_FP16X3U) precision = NB200_GEMM_TF32X3;   // tiny shapes
    if (!tensor_path_ok(g) || force_simt) {
        for (int64_t b0 = 0; b0 < g.batch; b0 += 65535) {   // the batch rides on gridDim.z
            const int64_t nb = g.batch - b0 < 65535 ? g.batch - b0 : 65535;
            dim3 grid((unsigned)((g.N + 63) / 64), (unsigned)((g.M + 63) / 64), (unsigned)nb);
            sgemm_simt_kernel<<<grid, 256, 0, ctx().stream>>>(g.C + b0 * g.sC, g.A + b0 * g.sA, g.B + b0 * g.sB, g.M, g.N, g.K, g.lda, g.ldb, g.ldc,
                                                              g.sA, g.sB, g.sC);
            NB_LAUNCH_CHECK();
        }
        return NB200_OK;
    }
    if (precision == NB200_GEMM_TF32X1) return dispatch_cfg<1>(g);
    // Measured on B200 (profiles/r1_summary.md §4): producing the lo parts INSIDE the kernel (GemmCfg::INK, four converter
    // warps) halves TMA traffic but the extra LDS/STS contends with the tensor core's operand reads for shared-memory
    // bandwidth: 0.98 ms vs 0.53 ms at 4096^3.  The HBM-roofline pre-pass is therefore the default; NB200_GEMM_INKERNEL=1
    // selects the in-kernel variant for experiments.
    static const bool inkernel = getenv("NB200_GEMM_INKERNEL") != nullptr;
    if (inkernel) {
        g.A_lo = g.B_lo = nullptr;
        return dispatch_cfg<3>(g);
    }
    // ---- TF32x3 with the lo-part pre-pass into the context workspace, batch processed in chunks
    int64_t chunk = g.batch;
    const int64_t budget = gemm_ws_budget();
    if (g.batch > 1) {
        int64_t per = 4 * ((g.sA ? round4(g.sA) : 0) + (g.sB ? round4(g.sB) : 0));
        if (per > 0 && per * chunk > budget) chunk = budget / per;
        if (chunk < 1) chunk = 1;
    }
    { const int rcf = gemm_reset_nonfinite(); if (rcf != NB200_OK) return rcf; }   // once per call: shared operands are split once
    for (int64_t b0 = 0; b0 < g.batch; b0 += chunk) {
        const int64_t nb = g.batch - b0 < chunk ? g.batch - b0 : chunk;
        const int64_t sa = span(g.sA ? nb : 1, g.sA, g.M, g.lda, g.K), sb = span(g.sB ? nb : 1, g.sB, g.K, g.ldb, g.N);
        // workspace offsets follow the FULL chunk size (a shared operand split with the first chunk must not move)
        const int64_t na = round4(span(g.sA ? chunk : 1, g.sA, g.M, g.lda, g.K)), nbb = round4(span(g.sB ? chunk : 1, g.sB, g.K, g.ldb, g.N));
        int rc = ensure_gemm_ws((na + nbb) * 4 + 256);
        if (rc != NB200_OK) return rc;
        float *ws = static_cast<float *>(ctx().gemm_ws);
        float *a_lo = ws, *b_lo = ws + na;
        const float *a_src = g.A + (g.sA ? b0 * g.sA : 0), *b_src = g.B + (g.sB ? b0 * g.sB : 0);
        // a shared (stride-0) operand is split once, with the first chunk
        const bool do_a = (b0 == 0 || g.sA), do_b = (b0 == 0 || g.sB);
        // the vector path needs 16-byte aligned sources (tensor_path_ok guarantees it for the bases and strides)
        if ((rc = launch_split(a_src, a_lo, do_a ? sa : 0, b_src, b_lo, do_b ? sb : 0)) != NB200_OK) return rc;
        GemmArgs c = g;
        c.batch = nb;
        c.A = a_src; c.A_lo = a_lo; c.B = b_src; c.B_lo = b_lo;
        c.C = g.C + b0 * g.sC;
        if ((rc = dispatch_cfg<3>(c)) != NB200_OK) return rc;
    }
    return NB200_OK;
}

// ---- internal hooks for the host-buffer pipeline (host_pipeline.cu): split one operand / run with given lo parts
int gemm_split_operand(const float *in, float *lo, int64_t n) { return launch_split(in, lo, n, nullptr, nullptr, 0); }
// BF16x3 flavour of the two hooks: packed bf16 operand pairs (leading dimension = cols rounded up to 8)
int gemm_bf16_split(const float *in, void *hi, void *lo, int64_t rows, int64_t cols) {
    SplitSpan s = make_span(in, static_cast<__nv_bfloat16 *>(hi), static_cast<__nv_bfloat16 *>(lo), 1, rows, cols, cols, 0);
    SplitSpan none = make_span(nullptr, nullptr, nullptr, 0, 0, 0, 0, 0);
    return launch_split_bf16(s, none);
}
int gemm_bf16_presplit(float *C, const void *a_hi, const void *a_lo, const void *b_hi, const void *b_lo, int64_t M, int64_t N,
                       int64_t K, int64_t ldc) {
    GemmArgs g{C, static_cast<const float *>(a_hi), static_cast<const float *>(b_hi), static_cast<const float *>(a_lo),
               static_cast<const float *>(b_lo), 1, M, N, K, round8(K), round8(N), ldc, 0, 0, 0};
    const int v = gemm_variant();
    const int cg = v ? (v >> 8) : (M > 128 ? 2 : 1);
    const int bn = v ? (v & 0xFF) * 2 : (cg == 2 ? bf16_pair_bn(1, M, N) : 128);
    if (cg == 2 && bn == 256) return launch_gemm<GemmCfg<2, 256, 3, false, true, true>>(g);
    return cg == 2 ? launch_gemm<GemmCfg<2, 128, 3, false, true>>(g) : launch_gemm<GemmCfg<1, 128, 3, false, true>>(g);
}
int gemm_presplit(float *C, const float *A, const float *A_lo, const float *B, const float *B_lo, int64_t M, int64_t N,
                  int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int precision) {
    GemmArgs g{C, A, B, A_lo, B_lo, 1, M, N, K, lda, ldb, ldc, 0, 0, 0};
    if (!tensor_path_ok(g)) return set_error(NB200_EINVAL, "gemm_presplit: shape not served by the tensor path");
    return precision == NB200_GEMM_TF32X1 ? dispatch_cfg<1>(g) : dispatch_cfg<3>(g);
}

}  // namespace nb200

using namespace nb200;

extern "C" int nb200_sgemm_batched(float *C, const float *A, const float *B, int64_t batch, int64_t M, int64_t N,
                                   int64_t K, int64_t strideA, int64_t strideB, int64_t strideC, int precision) {
    NB_READY();
    if (!C || !A || !B || batch < 0 || M < 0 || N < 0 || K < 0 || strideA < 0 || strideB < 0 || strideC < 0)
        return set_error(NB200_EINVAL, "nb200_sgemm_batched: bad argument");
    if (precision < NB200_GEMM_TF32X3 || precision > NB200_GEMM_FP16X3U)
        return set_error(NB200_EINVAL, "nb200_sgemm: unknown precision %d", precision);
    GemmArgs g{C, A, B, nullptr, nullptr, batch, M, N, K, K, N, N, strideA, strideB, strideC};
    return gemm_impl(g, precision);
}

extern "C" int nb200_sgemm(float *C, const float *A, const float *B, int64_t M, int64_t N, int64_t K, int64_t lda,
                           int64_t ldb, int64_t ldc, int precision) {
    NB_READY();
    if (!C || !A || !B || M < 0 || N < 0 || K < 0 || lda < K || ldb < N || ldc < N)
        return set_error(NB200_EINVAL, "Shape mismatch for matmul (M=%lld N=%lld K=%lld lda=%lld ldb=%lld ldc=%lld)",
                         (long long)M, (long long)N, (long long)K, (long long)lda, (long long)ldb, (long long)ldc);
    if (precision < NB200_GEMM_TF32X3 || precision > NB200_GEMM_FP16X3U)
        return set_error(NB200_EINVAL, "nb200_sgemm: unknown precision %d", precision);
    GemmArgs g{C, A, B, nullptr, nullptr, 1, M, N, K, lda, ldb, ldc, 0, 0, 0};
    return gemm_impl(g, precision);
}

extern "C" int nb200_gemm_resolve_precision(int precision, int64_t K) { return gemm_resolve_precision(precision, K); }

extern "C" int nb200_sgemm_workspace_bytes(int64_t batch, int64_t M, int64_t N, int64_t K, int precision, int64_t *bytes) {
    if (!bytes) return set_error(NB200_EINVAL, "null argument");
    if (precision == NB200_GEMM_TF32X1) { *bytes = 0; return NB200_OK; }
    precision = gemm_resolve_precision(precision, K);
    int64_t per = precision == NB200_GEMM_BF16X3 ? 4 * (M * round8(K) + K * round8(N)) + 1024
                : (precision == NB200_GEMM_FP16X3 || precision == NB200_GEMM_FP16X3U) ? 4 * (M * round8(K) + K * round8(N)) + 4 * round4(M) + 2 * ((16 + 4 * round4(N) + 255) & ~int64_t(255)) + 4 * (round4(M * K) + round4(K * N)) + 2 * (int64_t)FIX_CAP * 16 + 1024
                                                 : 4 * (round4(M * K) + round4(K * N));
    int64_t total = per * batch;
    const int64_t budget = gemm_ws_budget();
    *bytes = (total > budget && batch > 1) ? (budget / per > 0 ? (budget / per) * per : per) : total;
    return NB200_OK;
}
