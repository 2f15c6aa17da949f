// Shared tcgen05 / TMA / mbarrier helpers of the tensor-core kernels (tc_gemm.cu, conv_tc.cu).
#pragma once
#include "mms_common.cuh"
#include <cuda.h>

namespace mms {

constexpr int TC_BM = 128;          // rows per CTA tile = UMMA M
constexpr int TC_BK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 192;
constexpr int TC_UMMA_K = 8;        // tf32: 32 bytes per MMA k-step

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// contiguous global -> shared bulk copy (cp.async.bulk; 16-byte aligned, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}

// kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                  // c_format = F32
    d |= 2u << 7;                  // a_format = TF32
    d |= 2u << 10;                 // b_format = TF32
    d |= (uint32_t)(n >> 3) << 17; // n_dim
    d |= (uint32_t)(TC_BM >> 4) << 24;  // m_dim
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MN-major 32-bit operands have ONE legal shared-memory layout: SWIZZLE_128B with a 32-byte base (Swizzle<2,5,2> on the
// byte address): rows of 128 bytes (32 elements along M/N), the 32-byte chunks of a row XOR-ed with (row % 4); atoms of
// 4 k-rows.  TMA produces it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t block_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
    d |= (uint64_t)(block_bytes >> 4) << 16;            // leading byte offset: next 32-element block along M/N
    d |= (uint64_t)(512 >> 4) << 32;                    // stride byte offset: next atom of 4 k-rows
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int n) {
    return umma_idesc_tf32(n) | (1u << 15) | (1u << 16);   // A and B both MN-major
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (floats); box = [box_rows, 32 cols], 128-byte swizzle
static inline int make_map(CUtensorMap* map, const float* g, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int box_cols = TC_BK) {
    EncodeTiledFn fn = encode_tiled_fn();
    MMS_REQUIRE(fn, "tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
    MMS_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (ld * 4) % 16 == 0, "tc_gemm: operand must be 16-byte aligned with ld %% 4 == 0");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)g, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tc_gemm: cuTensorMapEncodeTiled failed with %d (rows %lld cols %lld ld %lld box %d)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld, box_rows);
        return MMS_E_CUDA;
    }
    return MMS_OK;
}

}  // namespace mms
