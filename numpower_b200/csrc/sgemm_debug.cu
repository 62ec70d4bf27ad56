// Bring-up probe for the tcgen05 path (test infrastructure compiled into libnb200.so, not on any
// product path): one CTA, one 128 x 128 x 32 tile.  Dumps (1) the shared-memory image after the TMA
// loads, (2) a tcgen05.st -> tcgen05.ld round trip, (3) the TMEM accumulator after 4 K=8 MMAs.
// flags bit0: fill smem by hand (swizzled) instead of TMA; bit1: B operand K-major (expects B^T in bt).
#include "common.cuh"
#include <cuda.h>

namespace nb200 {
namespace dbg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const float *A, const float *B, const float *Bt, int lda, int ldb, int ldbt, int flags,
             int b_layout, int b_lbo, int b_sbo, int b_kstep,
             float *smem_dump /*[2][4096]*/, float *st_ld_dump /*[128][32]*/, float *acc_dump /*[128][128]*/,
             unsigned int *info) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    float *sm = reinterpret_cast<float *>(raw + (base - smem_u32(raw)));
    const uint32_t a_s = base, b_s = base + 16384, bar = base + 32768, bar2 = bar + 8, tptr = bar + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool manual = flags & 1, bk = flags & 2;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
    if (tid == 0) { info[0] = tmem; info[1] = base; }

    // ---- (1) operands into shared memory
    if (!manual) {
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32768));
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(a_s), "l"(&tmA), "r"(bar), "r"(0), "r"(0), "r"(0) : "memory");
            if (!bk) {
                for (int j = 0; j < 4; j++)
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(b_s + j * 4096), "l"(&tmB), "r"(bar), "r"(j * 32), "r"(0), "r"(0) : "memory");
            } else {
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(b_s), "l"(&tmB), "r"(bar), "r"(0), "r"(0), "r"(0) : "memory");
            }
        }
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 24) && !ok; spin++)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
        if (tid == 0) info[2] = ok;
    } else {
        // hand-written 128B swizzle: 16-byte chunk index (bits 4-6) ^= row-in-atom (bits 7-9)
        for (int i = tid; i < 128 * 32; i += 128) {      // A: [m][k] rows of 128 B
            int m = i / 32, k = i % 32;
            uint32_t off = m * 128 + k * 4;
            off ^= ((off >> 7) & 7) << 4;
            sm[off / 4] = A[m * lda + k];
        }
        for (int i = tid; i < 128 * 32; i += 128) {
            uint32_t off;
            float v;
            if (!bk) {                                     // B MN-major: chunk j, k-row, 32 n
                int k = i / 128, n = i % 128;
                off = (n / 32) * 4096 + k * 128 + (n % 32) * 4;
                v = B[k * ldb + n];
            } else {                                       // B K-major: [n][k]
                int n = i / 32, k = i % 32;
                off = n * 128 + k * 4;
                v = Bt[n * ldbt + k];
            }
            if (!bk && b_layout == 1) off ^= ((off >> 7) & 3) << 5;   // 128B span, 32B atoms
            else off ^= ((off >> 7) & 7) << 4;
            sm[4096 + off / 4] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    for (int i = tid; i < 8192; i += 128) smem_dump[i] = sm[i];

    // ---- (2) tcgen05.st / tcgen05.ld round trip on columns [128, 160)
    {
        uint32_t v[32];
        for (int q = 0; q < 32; q++) v[q] = __float_as_uint((float)(tid * 100 + q));
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                     "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                       "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
                       "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
                       "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 32; q++) st_ld_dump[tid * 32 + q] = __uint_as_float(r[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- (3) four K=8 MMAs into columns [0,128), then read back
    if (warp == 1) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((bk ? 0u : 1u) << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            info[3] = idesc;
            for (int k = 0; k < 4; k++) {
                auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
                    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
                    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
                    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
                    d |= (uint64_t)1 << 46;
                    d |= (uint64_t)layout << 61;
                    return d;
                };
                uint64_t da = desc(a_s + k * 32, 16, 1024, 2);
                uint64_t db = bk ? desc(b_s + k * 32, 16, 1024, 2) : desc(b_s + k * b_kstep, b_lbo, b_sbo, b_layout);
                if (k == 0) { info[4] = (uint32_t)da; info[5] = (uint32_t)(da >> 32); info[6] = (uint32_t)db; info[7] = (uint32_t)(db >> 32); }
                uint32_t acc = k != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
        }
        __syncwarp();
    }
    {
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 24) && !ok; spin++)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar2), "r"(0) : "memory");
        if (tid == 0) info[8] = ok;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < 4; c++) {
        uint32_t r[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 32; q++) acc_dump[tid * 128 + c * 32 + q] = __uint_as_float(r[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

}  // namespace dbg
}  // namespace nb200

using namespace nb200;

typedef CUresult (*EncodeTiledFn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// A[128,K>=32] lda, B[K>=32,128] ldb, Bt[128,K] ldbt (B transposed).  Returns CUDA status info in `info`.
extern "C" int nb200_debug_tcgen05_probe(const float *A, const float *B, const float *Bt, int lda, int ldb, int ldbt,
                                         int flags, int b_layout, int b_lbo, int b_sbo, int b_kstep, float *smem_dump, float *st_ld_dump, float *acc_dump,
                                         unsigned int *info) {
    NB_READY();
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    NB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn2 enc = reinterpret_cast<EncodeTiledFn2>(fp);
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[3] = {(cuuint64_t)lda, 128, 1};
        cuuint64_t strides[2] = {(cuuint64_t)lda * 4, (cuuint64_t)lda * 4 * 128};
        cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(A), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(NB200_ECUDA, "encode A failed %d", (int)r);
    }
    if (!(flags & 2)) {
        cuuint64_t dims[3] = {(cuuint64_t)ldb, 32, 1};
        cuuint64_t strides[2] = {(cuuint64_t)ldb * 4, (cuuint64_t)ldb * 4 * 32};
        cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(B), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, b_layout == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(NB200_ECUDA, "encode B failed %d", (int)r);
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)ldbt, 128, 1};
        cuuint64_t strides[2] = {(cuuint64_t)ldbt * 4, (cuuint64_t)ldbt * 4 * 128};
        cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(Bt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(NB200_ECUDA, "encode Bt failed %d", (int)r);
    }
    NB_CUDA(cudaFuncSetAttribute(dbg::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
    dbg::probe_kernel<<<1, 128, 40000, ctx().stream>>>(ma, mb, A, B, Bt, lda, ldb, ldbt, flags, b_layout, b_lbo, b_sbo, b_kstep, smem_dump, st_ld_dump, acc_dump, info);
    NB_LAUNCH_CHECK();
    NB_CUDA(cudaStreamSynchronize(ctx().stream));
    return NB200_OK;
}
