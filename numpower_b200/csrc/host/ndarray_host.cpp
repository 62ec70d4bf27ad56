// Host-side mirror of NumPower's NDArray operator layer (include/nb200_host.h) on the nb200 C-ABI.
// Shape logic, broadcasting rules, result allocation and error strings follow the reference's
// src/ndmath/arithmetics.c, src/ndarray.c, src/ndmath/calculation.c and src/ndmath/linalg.c
// (cited per function); all arithmetic happens in libnb200's CUDA kernels.
#include "../../../include/nb200.h"
#include "../../../include/nb200_host.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
thread_local char g_err[512] = "";
void *fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return nullptr;
}
void *fail_backend(const char *what) { return fail("%s: %s", what, nb200_last_error()); }

NB_NDArray *make(int ndim, const int64_t *shape, int device, bool alloc) {
    if (ndim < 0 || ndim > 8) return (NB_NDArray *)fail("ndim %d unsupported (max 8)", ndim);
    NB_NDArray *a = (NB_NDArray *)calloc(1, sizeof(NB_NDArray));
    a->ndim = ndim;
    a->numel = 1;
    for (int i = 0; i < ndim; i++) {
        if (shape[i] < 0) { free(a); return (NB_NDArray *)fail("negative dimension"); }
        a->shape[i] = shape[i];
        a->numel *= shape[i];
    }
    a->device = device;
    a->refcount = 1;
    if (alloc) {
        if (device == NB_DEVICE_GPU) {
            if (nb200_alloc((void **)&a->data, a->numel * 4) != NB200_OK) { free(a); return (NB_NDArray *)fail_backend("vmalloc"); }
        } else {
            a->data = (float *)malloc(a->numel > 0 ? a->numel * 4 : 4);
        }
    }
    return a;
}

// NumPy-style broadcast of two shapes (superset of NDArray_IsBroadcastable, ndarray.c:1124-1162):
// element strides, 0 where the operand is broadcast.
bool broadcast2(const NB_NDArray *a, const NB_NDArray *b, int *ndim, int64_t *shape, int64_t *sa, int64_t *sb) {
    int n = a->ndim > b->ndim ? a->ndim : b->ndim;
    int64_t stra = 1, strb = 1;
    for (int i = n - 1; i >= 0; i--) {
        int ia = i - (n - a->ndim), ib = i - (n - b->ndim);
        int64_t da = ia >= 0 ? a->shape[ia] : 1, db = ib >= 0 ? b->shape[ib] : 1;
        if (da != db && da != 1 && db != 1) return false;
        shape[i] = da == 1 ? db : da;
        sa[i] = (da == 1 && shape[i] != 1) ? 0 : stra;
        sb[i] = (db == 1 && shape[i] != 1) ? 0 : strb;
        if (da == 1) sa[i] = 0;
        if (db == 1) sb[i] = 0;
        stra *= da;
        strb *= db;
    }
    *ndim = n;
    return true;
}

bool gpu_pair(const NB_NDArray *a, const NB_NDArray *b) {
    // arithmetics.c:163-166: 0-dim scalars are exempt from the device check
    if (a->device != b->device && a->ndim != 0 && b->ndim != 0) {
        fail("Device mismatch, both NDArray MUST be in the same device.");
        return false;
    }
    const NB_NDArray *big = a->ndim == 0 ? b : a;
    if (big->device != NB_DEVICE_GPU) {
        fail("NDArray is on the CPU: this backend computes on the GPU only (call gpu() first; the CPU path is the reference's)");
        return false;
    }
    return true;
}
float scalar_value(const NB_NDArray *s, bool *ok) {
    float v = 0.f;
    *ok = true;
    if (s->device == NB_DEVICE_CPU) v = s->data[0];
    else if (nb200_copy_d2h(&v, s->data, 4) != NB200_OK) *ok = false;
    return v;
}
}  // namespace

extern "C" {

const char *NB_last_error(void) { return g_err; }

NB_NDArray *NB_NDArray_FromHost(const float *data, int ndim, const int64_t *shape) {
    NB_NDArray *a = make(ndim, shape, NB_DEVICE_CPU, true);
    if (a && a->numel > 0) memcpy(a->data, data, a->numel * 4);
    return a;
}
NB_NDArray *NB_NDArray_Empty(int ndim, const int64_t *shape, int device) { return make(ndim, shape, device, true); }

// NDArray_ToGPU ndarray.c:1037-1068 (the reference stages through pageable memory + cudaMemcpy)
NB_NDArray *NB_NDArray_ToGPU(NB_NDArray *a) {
    NB_NDArray *r = make(a->ndim, a->shape, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    int rc = a->device == NB_DEVICE_GPU ? nb200_copy_d2d(r->data, a->data, a->numel * 4) : nb200_copy_h2d(r->data, a->data, a->numel * 4);
    if (rc != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("NDArray_ToGPU"); }
    return r;
}
// NDArray_ToCPU ndarray.c:1075-1093
NB_NDArray *NB_NDArray_ToCPU(NB_NDArray *a) {
    NB_NDArray *r = make(a->ndim, a->shape, NB_DEVICE_CPU, true);
    if (!r) return nullptr;
    if (a->device == NB_DEVICE_CPU) memcpy(r->data, a->data, a->numel * 4);
    else if (nb200_copy_d2h(r->data, a->data, a->numel * 4) != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("NDArray_ToCPU"); }
    return r;
}
int NB_NDArray_CopyToHost(NB_NDArray *a, float *dst) {
    if (a->device == NB_DEVICE_CPU) { memcpy(dst, a->data, a->numel * 4); return 0; }
    if (nb200_copy_d2h(dst, a->data, a->numel * 4) != NB200_OK) { fail_backend("toArray"); return -1; }
    return 0;
}
// $a[i] — NDArrayIterator_GET iterators.c:94-111: data + i*strides[0], shape without dim 0, base = a
NB_NDArray *NB_NDArray_Slice0(NB_NDArray *a, int64_t index) {
    if (a->ndim < 1 || index < 0 || index >= a->shape[0]) return (NB_NDArray *)fail("Index out of bounds");
    NB_NDArray *r = make(a->ndim - 1, a->shape + 1, a->device, false);
    if (!r) return nullptr;
    r->data = a->data + index * r->numel;
    r->base = a;
    a->refcount++;
    return r;
}
NB_NDArray *NB_NDArray_Reshape(NB_NDArray *a, int ndim, const int64_t *shape) {
    NB_NDArray *r = make(ndim, shape, a->device, false);
    if (!r) return nullptr;
    if (r->numel != a->numel) { free(r); return (NB_NDArray *)fail("NDArray Reshape: Incompatible shape"); }
    r->data = a->data;
    r->base = a;
    a->refcount++;
    return r;
}
void NB_NDArray_FREE(NB_NDArray *a) {
    if (!a) return;
    if (--a->refcount > 0) return;
    if (a->base) NB_NDArray_FREE(a->base);
    else if (a->data) {
        if (a->device == NB_DEVICE_GPU) nb200_free(a->data);
        else free(a->data);
    }
    free(a);
}

// Shared skeleton of NDArray_{Add,...}_Float (arithmetics.c:160-278): device check, scalar operand,
// broadcast, fresh contiguous result of the larger operand's shape.  Differences from the
// reference: the scalar is passed by value (no NDArray_Fill temp, :169-181) and the broadcast
// operand is never materialised (no NDArray_Broadcast copy, :186-197) — stride-0 views instead.
NB_NDArray *NB_NDArray_Binary(int op, NB_NDArray *a, NB_NDArray *b) {
    if (!a || !b) return (NB_NDArray *)fail("null operand");
    if (!gpu_pair(a, b)) return nullptr;
    if ((a->ndim == 0) != (b->ndim == 0)) {
        NB_NDArray *arr = a->ndim == 0 ? b : a, *sc = a->ndim == 0 ? a : b;
        bool ok;
        float s = scalar_value(sc, &ok);
        if (!ok) return (NB_NDArray *)fail_backend("scalar read");
        NB_NDArray *r = make(arr->ndim, arr->shape, NB_DEVICE_GPU, true);
        if (!r) return nullptr;
        if (nb200_ew_binary_scalar(op, r->data, arr->data, s, a->ndim == 0, arr->numel) != NB200_OK) {
            NB_NDArray_FREE(r);
            return (NB_NDArray *)fail_backend("elementwise");
        }
        return r;
    }
    int ndim;
    int64_t shape[8], sa[8], sb[8];
    if (!broadcast2(a, b, &ndim, shape, sa, sb)) return (NB_NDArray *)fail("Can't broadcast arrays.");  // arithmetics.c:199-202
    NB_NDArray *r = make(ndim, shape, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    const float *pa = a->data, *pb = b->data;
    float *tmp = nullptr;
    if (a->device == NB_DEVICE_CPU || b->device == NB_DEVICE_CPU) {  // a 0-dim CPU scalar next to a 0-dim GPU scalar
        NB_NDArray *c = a->device == NB_DEVICE_CPU ? a : b;
        if (nb200_alloc((void **)&tmp, 4) != NB200_OK || nb200_copy_h2d(tmp, c->data, 4) != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("scalar upload"); }
        if (c == a) pa = tmp; else pb = tmp;
    }
    int rc = nb200_ew_binary(op, r->data, pa, pb, ndim, shape, sa, sb);
    if (tmp) { nb200_synchronize(); nb200_free(tmp); }
    if (rc != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("elementwise"); }
    return r;
}
NB_NDArray *NB_NDArray_Add_Float(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_ADD, a, b); }
NB_NDArray *NB_NDArray_Subtract_Float(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_SUB, a, b); }
NB_NDArray *NB_NDArray_Multiply_Float(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_MUL, a, b); }
NB_NDArray *NB_NDArray_Divide_Float(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_DIV, a, b); }
NB_NDArray *NB_NDArray_Mod_Float(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_MOD, a, b); }
NB_NDArray *NB_NDArray_Pow_Float(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_POW, a, b); }
NB_NDArray *NB_NDArray_Maximum(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_MAXIMUM, a, b); }
NB_NDArray *NB_NDArray_Minimum(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_MINIMUM, a, b); }
NB_NDArray *NB_NDArray_Arctan2(NB_NDArray *a, NB_NDArray *b) { return NB_NDArray_Binary(NB200_ARCTAN2, a, b); }

NB_NDArray *NB_NDArray_MulAdd(NB_NDArray *a, NB_NDArray *b, NB_NDArray *c) {
    if (!a || !b || !c) return (NB_NDArray *)fail("null operand");
    if (a->device != NB_DEVICE_GPU || b->device != NB_DEVICE_GPU || c->device != NB_DEVICE_GPU)
        return (NB_NDArray *)fail("Device mismatch, both NDArray MUST be in the same device.");
    int nd1, nd2;
    int64_t s1[8], sa[8], sb[8], s2[8], sab[8], sc[8];
    if (!broadcast2(a, b, &nd1, s1, sa, sb)) return (NB_NDArray *)fail("Can't broadcast arrays.");
    NB_NDArray ab;
    memset(&ab, 0, sizeof(ab));
    ab.ndim = nd1;
    memcpy(ab.shape, s1, sizeof(s1));
    if (!broadcast2(&ab, c, &nd2, s2, sab, sc)) return (NB_NDArray *)fail("Can't broadcast arrays.");
    // re-express a and b strides in the final (nd2) shape
    int64_t fa[8], fb[8];
    NB_NDArray fin;
    memset(&fin, 0, sizeof(fin));
    fin.ndim = nd2;
    memcpy(fin.shape, s2, sizeof(s2));
    int64_t dummy[8], dshape[8];
    int dn;
    if (!broadcast2(a, &fin, &dn, dshape, fa, dummy) || !broadcast2(b, &fin, &dn, dshape, fb, dummy))
        return (NB_NDArray *)fail("Can't broadcast arrays.");
    NB_NDArray *r = make(nd2, s2, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    if (nb200_ew_mul_add(r->data, a->data, b->data, c->data, nd2, s2, fa, fb, sc) != NB200_OK) {
        NB_NDArray_FREE(r);
        return (NB_NDArray *)fail_backend("mul_add");
    }
    return r;
}

// NDArray_Map / Map1F / Map2F (ndarray.c:682-744): out = zeros-like ; out[i] = op(in[i]) flat
NB_NDArray *NB_NDArray_Map(NB_NDArray *a, int op, float p0, float p1) {
    if (!a) return (NB_NDArray *)fail("null operand");
    if (a->device != NB_DEVICE_GPU) return (NB_NDArray *)fail("NDArray is on the CPU: this backend computes on the GPU only");
    NB_NDArray *r = make(a->ndim, a->shape, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    if (nb200_ew_unary(op, r->data, a->data, a->numel, p0, p1) != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("map"); }
    if (op == NB200_UN_ARCCOS || op == NB200_UN_ARCCOSH || op == NB200_UN_ARCTANH) {
        int flag = 0;
        if (nb200_poll_domain_error(&flag) == NB200_OK && flag) {
            // double_math.c:145-148, 181-184, 189-198: the reference prints this and exit(1)s
            NB_NDArray_FREE(r);
            return (NB_NDArray *)fail("RuntimeError: Invalid argument provided for %s",
                                      op == NB200_UN_ARCCOS ? "arccos" : op == NB200_UN_ARCCOSH ? "arccosh" : "arctanh");
        }
    }
    return r;
}

static int full_reduce(NB_NDArray *a, int op, float *out) {
    if (!a || !out) { fail("null argument"); return -1; }
    if (a->device != NB_DEVICE_GPU) { fail("NDArray is on the CPU: this backend computes on the GPU only"); return -1; }
    if (a->numel == 0) { *out = op == NB200_PROD ? 1.f : 0.f; return 0; }  // empty loops of arithmetics.c:44,66
    if (nb200_reduce_full_host(op, out, a->data, a->numel) != NB200_OK) { fail_backend("reduce"); return -1; }
    return 0;
}
int NB_NDArray_Sum_Float(NB_NDArray *a, float *out) { return full_reduce(a, NB200_SUM, out); }
int NB_NDArray_Float_Prod(NB_NDArray *a, float *out) { return full_reduce(a, NB200_PROD, out); }
int NB_NDArray_Min(NB_NDArray *a, float *out) { return full_reduce(a, NB200_MIN, out); }
int NB_NDArray_Max(NB_NDArray *a, float *out) { return full_reduce(a, NB200_MAX, out); }

// reduce() ndarray.c:523-578: output shape = input shape minus `axis` (no keepdims)
NB_NDArray *NB_reduce(NB_NDArray *a, int axis, int op, int order) {
    if (!a) return (NB_NDArray *)fail("null operand");
    if (a->device != NB_DEVICE_GPU) return (NB_NDArray *)fail("NDArray is on the CPU: this backend computes on the GPU only");
    if (axis < 0 || axis >= a->ndim)
        return (NB_NDArray *)fail("axis %d is out of bounds for array of dimension %d", axis, a->ndim);  // ndarray.c:534-538
    int64_t oshape[8], outer = 1, inner = 1;
    int j = 0;
    for (int i = 0; i < a->ndim; i++) {
        if (i < axis) outer *= a->shape[i];
        if (i > axis) inner *= a->shape[i];
        if (i != axis) oshape[j++] = a->shape[i];
    }
    NB_NDArray *r = make(a->ndim - 1, oshape, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    if (a->shape[axis] == 0) {   // no slices: the reference returns its NDArray_Zeros result untouched (ndarray.c:568-569)
        if (r->numel > 0 && nb200_memset_zero(r->data, r->numel * 4) != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("reduce"); }
        return r;
    }
    if (nb200_reduce_axis(op, r->data, a->data, outer, a->shape[axis], inner, order) != NB200_OK) {
        NB_NDArray_FREE(r);
        return (NB_NDArray *)fail_backend("reduce");
    }
    return r;
}

// NDArray_ArgMinMaxCommon calculation.c:73-194.  No transpose/flatten copy is needed: the kernel
// indexes (outer, m, inner) directly.  Result: float32 indices; keepdims as in :141-160.
NB_NDArray *NB_NDArray_ArgMinMaxCommon(NB_NDArray *a, int axis, int keepdims, int is_argmax) {
    if (!a) return (NB_NDArray *)fail("null operand");
    if (a->device != NB_DEVICE_GPU) return (NB_NDArray *)fail("NDArray is on the CPU: this backend computes on the GPU only");
    int64_t outer = 1, m, inner = 1, oshape[8];
    int ondim;
    if (axis == NB_MAX_DIMS_AXIS || a->ndim == 0) {
        m = a->numel;
        ondim = keepdims ? a->ndim : 0;
        for (int i = 0; i < ondim; i++) oshape[i] = 1;
    } else {
        if (axis < 0) axis += a->ndim;
        if (axis < 0 || axis >= a->ndim) return (NB_NDArray *)fail("Invalid axis parameter");  // calculation.c:95-99
        m = a->shape[axis];
        int j = 0;
        for (int i = 0; i < a->ndim; i++) {
            if (i < axis) outer *= a->shape[i];
            if (i > axis) inner *= a->shape[i];
            if (i != axis) oshape[j++] = a->shape[i];
            else if (keepdims) oshape[j++] = 1;
        }
        ondim = j;
    }
    if (m == 0) return (NB_NDArray *)fail("attempt to get %s of an empty sequence", is_argmax ? "argmax" : "argmin");  // :169-172
    NB_NDArray *r = make(ondim, oshape, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    if (nb200_argminmax(is_argmax, r->data, a->data, outer, m, inner) != NB200_OK) {
        NB_NDArray_FREE(r);
        return (NB_NDArray *)fail_backend("argminmax");
    }
    return r;
}

// NDArray_Matmul linalg.c:216-245 -> NDArray_FMatmul :44-82
NB_NDArray *NB_NDArray_Matmul(NB_NDArray *a, NB_NDArray *b, int precision) {
    if (!a || !b) return (NB_NDArray *)fail("null operand");
    if (a->device != b->device) return (NB_NDArray *)fail("Device mismatch, both NDArray MUST be in the same device.");
    if (a->ndim != b->ndim) return (NB_NDArray *)fail("Arrays must have the same shape. Broadcasting not implemented.");
    if (a->ndim == 0) return NB_NDArray_Multiply_Float(a, b);
    if (a->ndim == 1) return NB_NDArray_Dot(a, b);
    if (a->device != NB_DEVICE_GPU) return (NB_NDArray *)fail("NDArray is on the CPU: this backend computes on the GPU only");
    if (a->shape[a->ndim - 1] != b->shape[b->ndim - 2]) return (NB_NDArray *)fail("Shape mismatch for matmul. cols(a) != rows(b)");
    // The reference rejects ndim > 2 ("Stack of matrices not allowed", linalg.c:240-243); here a stack of
    // equal leading dims runs as ONE batched launch (SURVEY.md §8 f, N1).
    int64_t batch = 1, oshape[8];
    for (int i = 0; i < a->ndim - 2; i++) {
        if (a->shape[i] != b->shape[i]) return (NB_NDArray *)fail("Stack of matrices not allowed (leading dimensions differ)");
        batch *= a->shape[i];
        oshape[i] = a->shape[i];
    }
    const int64_t M = a->shape[a->ndim - 2], K = a->shape[a->ndim - 1], N = b->shape[b->ndim - 1];
    oshape[a->ndim - 2] = M;
    oshape[a->ndim - 1] = N;
    NB_NDArray *r = make(a->ndim, oshape, NB_DEVICE_GPU, true);
    if (!r) return nullptr;
    int rc = batch == 1 ? nb200_sgemm(r->data, a->data, b->data, M, N, K, K, N, N, precision)
                        : nb200_sgemm_batched(r->data, a->data, b->data, batch, M, N, K, M * K, K * N, M * N, precision);
    if (rc != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("matmul"); }
    return r;
}

// NDArray_Dot linalg.c:354-393 ; NDArray_Inner :310-345 (= Sum_Float(Multiply(a, b)))
NB_NDArray *NB_NDArray_Dot(NB_NDArray *a, NB_NDArray *b) {
    if (!a || !b) return (NB_NDArray *)fail("null operand");
    if (a->device != b->device) return (NB_NDArray *)fail("Device mismatch, both NDArray MUST be in the same device.");
    if (a->ndim == 1 && b->ndim == 1) {
        if (a->shape[0] != b->shape[0]) return (NB_NDArray *)fail("Shape is not aligned to perform the inner product.");
        NB_NDArray *mul = NB_NDArray_Multiply_Float(a, b);
        if (!mul) return nullptr;
        NB_NDArray *r = make(0, nullptr, NB_DEVICE_GPU, true);
        int rc = r ? nb200_reduce_full(NB200_SUM, r->data, mul->data, mul->numel) : NB200_ENOMEM;
        nb200_synchronize();
        NB_NDArray_FREE(mul);
        if (rc != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("inner"); }
        return r;
    }
    if (a->ndim == 2 && b->ndim == 2) return NB_NDArray_Matmul(a, b, NB200_GEMM_AUTO);
    if (a->ndim == 0 || b->ndim == 0) return NB_NDArray_Multiply_Float(a, b);
    if (a->ndim > 0 && b->ndim == 1) {
        if (a->device != NB_DEVICE_GPU) return (NB_NDArray *)fail("NDArray is on the CPU: this backend computes on the GPU only");
        const int64_t cols = a->shape[a->ndim - 1], rows = a->numel / (cols ? cols : 1);
        if (cols != b->shape[0]) return (NB_NDArray *)fail("Shape is not aligned to perform the dot product.");
        NB_NDArray *r = make(a->ndim - 1, a->shape, NB_DEVICE_GPU, true);
        if (!r) return nullptr;
        if (nb200_gemv(r->data, a->data, b->data, rows, cols) != NB200_OK) { NB_NDArray_FREE(r); return (NB_NDArray *)fail_backend("dot"); }
        return r;
    }
    return (NB_NDArray *)fail("Not implemented");  // linalg.c:387-390
}

}  // extern "C"
