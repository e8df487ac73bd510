// conv_tcgen05_rows.cu -- fused conv2d forward for short reductions (K = kh*kw*c <= 32: the first-layer
// shapes 3x3x3, 5x5x1, ...) as an implicit GEMM on tcgen05 with the A operand in TENSOR MEMORY.
//
//   out[b,oy,ox,f] = sum_{ky,kx,c} images[b, oy*sh+ky*dh-ph, ox*sw+kx*dw-pw, c] * W[(ky,kx,c), f] (+ bias[f])
//   == gemm(im2col2d(images) viewed [B*oh*ow, K], W)   (PhpMath::im2col2d :2158-2314 + PhpBlas::gemm :927-947)
//
// Why a kernel of its own: the op moves 846 MB for BASELINE config C4 (0.13 ms at HBM speed), and the
// shared-memory-staged variant in gemm_tcgen05_skinny.cu is bound by shared-memory bandwidth and by a
// serial epilogue -> MMA -> epilogue chain (measured with the clock-stamp trace below).  Here:
//   * the cols rows never touch shared memory: a producer thread owns one output cell, reads its K inputs
//     from a TMA-loaded image slab (LDS.32, lane stride str_w*chan floats), splits them into tf32 hi / lo
//     in registers and writes them with tcgen05.st into a TMEM A stage (lane = row, column = k);
//   * tcgen05.mma takes A from TMEM (`[d], [a], bdesc`): A_hi x [W_hi | W_lo] fills the big (hi*hi) and the
//     small (hi*lo) accumulator in one instruction, A_lo x W_hi adds the other cross term;
//   * up to four accumulator stages let the MMA warp run ahead of the epilogue;
//   * the epilogue groups double-buffer their 128-byte swizzled staging tiles, so a tile's TMA store drains
//     while the next tile is converted; TMA clips ragged tiles.
//   * TMA instructions cost the issuing warp ~300 cycles each (measured), so they live in warps of their own:
//     two slab loaders (one per producer group) and one store issuer, hand-shaking through mbarriers.
// Persistent CTA per SM, 640 threads: warps 0-7 producers (two groups of four, group g builds the tiles of
// parity g into A stage g), warp 8 MMA issuer, warps 9-16 epilogue (two groups of four, alternating tiles),
// warps 17-18 slab loaders, warp 19 store issuer.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int CR_BM = 128;
constexpr int CR_BK = 32;
constexpr int CR_THREADS = 20 * 32;
constexpr int CR_MMA_WARP = 8;
constexpr int CR_EPI_WARP0 = 9;
constexpr int CR_LOAD_WARP0 = 17;                    // warps 17, 18: slab loader of producer group 0 / 1
constexpr int CR_STORE_WARP = 19;                    // issues the TMA stores of both epilogue groups
constexpr int CR_SLOTS = 3;                          // slab slots per producer group
constexpr int CR_PREFETCH = 4;                       // L2 prefetch distance of the slab loaders (tiles of their group)
constexpr int CR_MAX_BOXES = 8;
constexpr int CR_MAX_ACC = 4;
constexpr int CR_SMEM_LIMIT = 226 * 1024;

__device__ __forceinline__ uint32_t cr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cr_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cr_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cr_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void cr_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cr_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// every lane executes these; one elected lane issues
__device__ __forceinline__ void cr_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// one k-step of 3xTF32 into one accumulator, cross terms first: D (+)= A_hi W_lo; D += A_lo W_hi; D += A_hi W_hi
__device__ __forceinline__ void cr_mma3_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                           uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, t, e;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %5, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %5, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %5, t;\n\t}"
        :: "r"(tmem_d), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cr_commit2(uint32_t bar0, uint32_t bar1)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}"
        :: "r"(bar0), "r"(bar1) : "memory");
}
__device__ __forceinline__ void cr_expect_tx_elect(uint32_t bar, uint32_t bytes)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cr_arrive_elect(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}"
        :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cr_tma_load3_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n\t}"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void cr_tma_store4_elect(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n\t}"
        :: "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}
__device__ __forceinline__ void cr_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void cr_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cr_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void cr_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled tile: byte offset of element (row r, column kk < 32)
__device__ __forceinline__ uint32_t cr_swz(int r, int kk)
{
    return (uint32_t)(r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2)));
}
__device__ __forceinline__ uint64_t cr_desc(uint32_t smem_addr)          // K-major SW128: SBO = 1024
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// round to tf32, nearest with ties away from zero (what cvt.rna.tf32.f32 expands to, minus its Inf/NaN test:
// Inf stays Inf, a NaN stays NaN in at least one of hi / lo)
__device__ __forceinline__ uint32_t cr_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ float cr_lds(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

struct CrParams {
    int N, K, NP;                  // filters, kh*kw*c (<= 32), N rounded up to 32
    const float* W; int64_t ldW;   // [K][N]
    const float* bias;             // optional [N]
    int num_tiles;
    uint32_t tmem_cols;
    int acc_stages;                // accumulator stages in TMEM (2..4)
    int stg_bufs;                  // staging tiles per epilogue group (1 or 2)
    int chan, str_h, str_w, pad_h, pad_w;
    const float* images; int im_h, im_w; int row_elems;   // for the L2 prefetch of future slabs (row_elems: floats per slab row)
    // tile = cv_rows out rows x cv_tw out columns of one image (<= 128 cells, cv_rows > 1 only when cv_tw == out_w)
    int cv_tw, cv_rows, cv_col_blocks, cv_row_blocks, cv_slab_rows;
    // slab = cv_nbox TMA boxes of cv_bw floats x cv_slab_rows rows; box i starts cv_box_start[i] floats into the
    // slab row (multiples of 4: TMA wants 16-byte aligned origins, the slab itself starts cv_xshift floats left
    // of the tile's first input column); column k of a cols row sits koff[k] bytes after the cell's run origin
    int cv_nbox, cv_bw, cv_box_elems, cv_slot_elems, cv_xshift;
    int cv_box_start[CR_MAX_BOXES];
    int koff[CR_BK];
    uint32_t off_stg, off_ring, off_misc;   // shared-memory layout (bytes from the 1024-aligned base)
    int trace_t0;                  // first traced tile of CTA 0 (RB200_SKT_TRACE_T0)
    long long* trace;              // diagnostic (RB200_SKT_TRACE=1): [64 tiles][16 events] SM clock stamps of CTA 0
};

__device__ __forceinline__ void cr_trace(const CrParams& p, int it, int e)
{
    it -= p.trace_t0;
    if (p.trace && blockIdx.x == 0 && it >= 0 && it < 64) p.trace[it * 16 + e] = clock64();
}

template <int PLANES, bool TRACE>
__global__ void __launch_bounds__(CR_THREADS, 1)
conv_rows_kernel(const __grid_constant__ CrParams p, const __grid_constant__ CUtensorMap tmapI,
                 const __grid_constant__ CUtensorMap tmapO)
{
    extern __shared__ uint8_t cr_raw[];
    const uint32_t base = (cr_smem_u32(cr_raw) + 1023u) & ~1023u;
    uint8_t* gen = cr_raw + (base - cr_smem_u32(cr_raw));
    // layout: W planes [plane][n][32 k] | staging tiles | slab ring | bias | barriers
    const uint32_t bar = base + p.off_misc + 512u;
    auto a_full   = [&](int s) { return bar + 8u * s; };
    auto a_empty  = [&](int s) { return bar + 8u * (2 + s); };
    auto t_full   = [&](int a) { return bar + 8u * (4 + a); };
    auto t_empty  = [&](int a) { return bar + 8u * (4 + CR_MAX_ACC + a); };
    auto slab_full  = [&](int g, int s) { return bar + 8u * (uint32_t)(4 + 2 * CR_MAX_ACC + g * CR_SLOTS + s); };
    auto slab_empty = [&](int g, int s) { return bar + 8u * (uint32_t)(4 + 2 * CR_MAX_ACC + 2 * CR_SLOTS + g * CR_SLOTS + s); };
    auto stg_full  = [&](int ge, int sb) { return bar + 8u * (uint32_t)(4 + 2 * CR_MAX_ACC + 4 * CR_SLOTS + ge * 2 + sb); };
    auto stg_empty = [&](int ge, int sb) { return bar + 8u * (uint32_t)(8 + 2 * CR_MAX_ACC + 4 * CR_SLOTS + ge * 2 + sb); };
    const uint32_t tmem_slot = bar + 8u * (uint32_t)(12 + 2 * CR_MAX_ACC + 4 * CR_SLOTS);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
    float* bias_s = reinterpret_cast<float*>(gen + p.off_misc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ACC_COLS = p.NP;                                         // one fp32 accumulator per tile (see the MMA warp)
    const uint32_t a_col0 = (uint32_t)(p.acc_stages * ACC_COLS);      // A stage s: columns a_col0 + s*PLANES*32 (hi | lo)

    if (warp == CR_MMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapI)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapO)) : "memory");
        for (int s = 0; s < 2; s++) { cr_mbar_init(a_full(s), 4); cr_mbar_init(a_empty(s), 1); }
        for (int a = 0; a < CR_MAX_ACC; a++) { cr_mbar_init(t_full(a), 1); cr_mbar_init(t_empty(a), 4); }
        for (int g = 0; g < 2; g++) {
            for (int s = 0; s < CR_SLOTS; s++) { cr_mbar_init(slab_full(g, s), 1); cr_mbar_init(slab_empty(g, s), 4); }
            for (int s = 0; s < 2; s++) { cr_mbar_init(stg_full(g, s), 4); cr_mbar_init(stg_empty(g, s), 1); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CR_EPI_WARP0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tmem_slot), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // stage W once: global [K][N] -> K-major [plane][n][32 k] swizzled; the lo plane directly follows the hi plane,
    // so one UMMA descriptor with N = 2*NP reads W_hi | W_lo as a single operand
    for (int e = threadIdx.x; e < CR_BK * p.NP; e += CR_THREADS) {
        const int k = e / p.NP, n = e - k * p.NP;
        float v = 0.f;
        if (k < p.K && n < p.N) v = __ldg(p.W + (int64_t)k * p.ldW + n);
        const uint32_t hi = cr_rna(v);
        const uint32_t o = cr_swz(n, k);
        *reinterpret_cast<uint32_t*>(gen + o) = hi;
        if (PLANES == 2) *reinterpret_cast<float*>(gen + o + (uint32_t)(p.NP * 128)) = v - __uint_as_float(hi);
    }
    for (int n = threadIdx.x; n < p.NP; n += CR_THREADS) bias_s[n] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cr_fence_before();
    __syncthreads();
    cr_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    const int step = gridDim.x;

    if (warp < CR_MMA_WARP) {
        // ===== producers: group g = warp / 4 builds the tiles of parity g into A stage g =====
        const int g = warp >> 2;
        const int r = threadIdx.x & (CR_BM - 1);                       // cols row of the tile == TMEM lane
        int rr = p.cv_rows > 1 ? r / p.cv_tw : 0;
        int cc = r - rr * p.cv_tw;
        if (r >= p.cv_rows * p.cv_tw) { rr = 0; cc = 0; }              // never-stored rows: stay inside the slab
        const int x0 = cc * p.str_w * p.chan + p.cv_xshift;
        int bi = 0;
        for (int i = 1; i < p.cv_nbox; i++) { if (p.cv_box_start[i] <= x0) bi = i; }
        const int roff = bi * p.cv_box_elems + rr * p.str_h * p.cv_bw + (x0 - p.cv_box_start[bi]);
        const uint32_t ring_g = base + p.off_ring + (uint32_t)(g * CR_SLOTS * p.cv_slot_elems) * 4u;
        const bool issuer = (warp & 3) == 0 && lane == 0;                // trace stamps only
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + a_col0 + (uint32_t)(g * PLANES * CR_BK);
        const int kq_last = (p.K - 1) >> 2;
        for (int j = 0; (int64_t)blockIdx.x + (int64_t)(2 * j + g) * step < p.num_tiles; j++) {
            const int slot = j % CR_SLOTS, n = j / CR_SLOTS;
            cr_mbar_wait(slab_full(g, slot), (uint32_t)(n & 1));
            if (TRACE && issuer) cr_trace(p, 2 * j + g, 0);
            const uint32_t sl = ring_g + (uint32_t)(slot * p.cv_slot_elems + roff) * 4u;
            float v[CR_BK];
#pragma unroll
            for (int q = 0; q < CR_BK / 4; q++) {
                if (4 * q < p.K) {                                       // koff[k >= K] repeats koff[0]: a valid address
#pragma unroll
                    for (int i = 0; i < 4; i++) v[4 * q + i] = cr_lds(sl + (uint32_t)p.koff[4 * q + i]);
                    if (q == kq_last) {                                  // the chunk K ends in: zero its tail
                        if (4 * q + 1 >= p.K) v[4 * q + 1] = 0.f;
                        if (4 * q + 2 >= p.K) v[4 * q + 2] = 0.f;
                        if (4 * q + 3 >= p.K) v[4 * q + 3] = 0.f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) v[4 * q + i] = 0.f;
                }
            }
            __syncwarp();
            if (lane == 0) cr_mbar_arrive(slab_empty(g, slot));
            if (TRACE && issuer) cr_trace(p, 2 * j + g, 1);
            cr_mbar_wait(a_empty(g), (uint32_t)((j & 1) ^ 1));           // the MMAs that read this A stage are done
            cr_fence_after();
            if (TRACE && issuer) cr_trace(p, 2 * j + g, 2);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    hi[e] = cr_rna(v[16 * h + e]);
                    lo[e] = __float_as_uint(v[16 * h + e] - __uint_as_float(hi[e]));
                }
                cr_st16(ta + 16 * h, hi);
                if (PLANES == 2) cr_st16(ta + CR_BK + 16 * h, lo);
            }
            cr_wait_st();
            cr_fence_before();
            __syncwarp();
            if (lane == 0) cr_mbar_arrive(a_full(g));
            if (TRACE && issuer) cr_trace(p, 2 * j + g, 3);
        }
    } else if (warp == CR_MMA_WARP) {
        // ===== MMA issuer: whole warp in the loop (uniform control flow), one elected lane issues =====
        const uint32_t idesc_n  = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(CR_BM >> 4) << 24);
        const int ksteps = (p.K + 7) >> 3;
        // W descriptors are tile invariant: keep them in registers
        uint64_t bh[CR_BK / 8], bl[CR_BK / 8];
#pragma unroll
        for (int kk = 0; kk < CR_BK / 8; kk++) {
            bh[kk] = cr_desc(base + kk * 32);
            bl[kk] = cr_desc(base + (uint32_t)(p.NP * 128) + kk * 32);
        }
        int acc = 0; uint32_t acc_phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += step, it++) {
            const int stage = it & 1;
            cr_mbar_wait(t_empty(acc), acc_phase ^ 1);
            if (TRACE && lane == 0) cr_trace(p, it, 4);
            cr_mbar_wait(a_full(stage), (uint32_t)((it >> 1) & 1));
            cr_fence_after();
            if (TRACE && lane == 0) cr_trace(p, it, 5);
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
            const uint32_t tmem_a = tmem_base + a_col0 + (uint32_t)(stage * PLANES * CR_BK);
#pragma unroll
            for (int kk = 0; kk < CR_BK / 8; kk++) {
                if (kk < ksteps) {
                    // 3xTF32 goes into ONE accumulator: with K <= 32 a tile sees at most 12 accumulation steps, so the
                    // truncating TMEM adds cost ~3e-7 relative -- not worth a second accumulator (half the TMEM loads
                    // and no big + small add in the epilogue)
                    if (PLANES == 1) cr_mma_ts(tmem_d, tmem_a + kk * 8, bh[kk], idesc_n, kk > 0 ? 1u : 0u);
                    else cr_mma3_ts(tmem_d, tmem_a + kk * 8, tmem_a + CR_BK + kk * 8, bh[kk], bl[kk], idesc_n, kk > 0 ? 1u : 0u);
                }
            }
            cr_commit2(a_empty(stage), t_full(acc));
            if (TRACE && lane == 0) cr_trace(p, it, 6);
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp < CR_LOAD_WARP0) {
        // ===== epilogue: group ge drains the tiles of parity ge; TMEM -> registers (+ bias) -> swizzled staging
        // tile -> stg_full; the store warp issues the TMA stores and hands the staging tile back (stg_empty) =====
        const int ge = (warp - CR_EPI_WARP0) >> 2;
        const int q = warp & 3;                                        // TMEM lane quarter of this warp
        const int r = q * 32 + lane;
        const uint32_t sw = (uint32_t)(r & 7);
        const bool leader = q == 0 && lane == 0;                       // trace stamps only
        const uint32_t stg_tile = (uint32_t)(p.NP >> 5) * 16384u;
        uint32_t soff[8];                                              // swizzled 16-byte chunk offsets of this thread's row
#pragma unroll
        for (int v = 0; v < 8; v++) soff[v] = ((uint32_t)v ^ sw) << 4;
        int acc = 0; uint32_t acc_phase = 0;
        int it = 0, jt = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += step, it++) {
            if ((it & 1) == ge) {
                const int sb = p.stg_bufs == 2 ? (jt & 1) : 0;
                const int use = p.stg_bufs == 2 ? (jt >> 1) : jt;      // how often this staging tile was used before
                if (TRACE && leader) cr_trace(p, it, 13);
                cr_mbar_wait(t_full(acc), acc_phase);
                cr_fence_after();
                if (TRACE && leader) cr_trace(p, it, 8);
                cr_mbar_wait(stg_empty(ge, sb), (uint32_t)((use & 1) ^ 1));   // its previous TMA store has read it
                if (TRACE && leader) cr_trace(p, it, 9);
                const uint32_t stg = base + p.off_stg + (uint32_t)(ge * p.stg_bufs + sb) * stg_tile;
                const uint32_t row_addr = stg + (uint32_t)r * 128u;
                const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);
                // 64 columns at a time: both TMEM loads in flight, one wait, then sixteen 128-bit stores whose
                // swizzled offsets are per-thread constants
#pragma unroll 1
                for (int c64 = 0; c64 < p.NP; c64 += 64) {
                    const bool two = c64 + 32 < p.NP;
                    uint32_t ra[32], rb[32];
                    cr_ld32(tacc + (uint32_t)c64, ra);
                    if (two) cr_ld32(tacc + (uint32_t)(c64 + 32), rb);
                    cr_wait_ld();
                    if (p.bias) {
                        const float4* bv = reinterpret_cast<const float4*>(bias_s + c64);
#pragma unroll
                        for (int v = 0; v < 8; v++) {
                            const float4 b4 = bv[v];
                            ra[4 * v]     = __float_as_uint(__uint_as_float(ra[4 * v]) + b4.x);
                            ra[4 * v + 1] = __float_as_uint(__uint_as_float(ra[4 * v + 1]) + b4.y);
                            ra[4 * v + 2] = __float_as_uint(__uint_as_float(ra[4 * v + 2]) + b4.z);
                            ra[4 * v + 3] = __float_as_uint(__uint_as_float(ra[4 * v + 3]) + b4.w);
                        }
                        if (two) {
#pragma unroll
                            for (int v = 0; v < 8; v++) {
                                const float4 b4 = bv[8 + v];
                                rb[4 * v]     = __float_as_uint(__uint_as_float(rb[4 * v]) + b4.x);
                                rb[4 * v + 1] = __float_as_uint(__uint_as_float(rb[4 * v + 1]) + b4.y);
                                rb[4 * v + 2] = __float_as_uint(__uint_as_float(rb[4 * v + 2]) + b4.z);
                                rb[4 * v + 3] = __float_as_uint(__uint_as_float(rb[4 * v + 3]) + b4.w);
                            }
                        }
                    }
                    const uint32_t box0 = row_addr + (uint32_t)(c64 >> 5) * 16384u;
#pragma unroll
                    for (int v = 0; v < 8; v++) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     :: "r"(box0 + soff[v]), "r"(ra[4 * v]), "r"(ra[4 * v + 1]), "r"(ra[4 * v + 2]), "r"(ra[4 * v + 3]) : "memory");
                    }
                    if (two) {
#pragma unroll
                        for (int v = 0; v < 8; v++) {
                            asm volatile("st.shared.v4.b32 [%0 + 16384], {%1, %2, %3, %4};"
                                         :: "r"(box0 + soff[v]), "r"(rb[4 * v]), "r"(rb[4 * v + 1]), "r"(rb[4 * v + 2]), "r"(rb[4 * v + 3]) : "memory");
                        }
                    }
                }
                if (TRACE && leader) cr_trace(p, it, 14);
                cr_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (TRACE && leader) cr_trace(p, it, 15);
                __syncwarp();
                if (lane == 0) { cr_mbar_arrive(t_empty(acc)); cr_mbar_arrive(stg_full(ge, sb)); }
                if (TRACE && leader) cr_trace(p, it, 10);
                jt++;
            }
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp < CR_STORE_WARP) {
        // ===== slab loader of producer group g: whole warp in the loop, one elected lane issues =====
        const int g = warp - CR_LOAD_WARP0;
        const int per_img = p.cv_row_blocks * p.cv_col_blocks;
        const uint32_t ring_g = base + p.off_ring + (uint32_t)(g * CR_SLOTS * p.cv_slot_elems) * 4u;
        const uint32_t slot_tx = (uint32_t)(p.cv_nbox * p.cv_bw * p.cv_slab_rows) * 4u;
        for (int jj = 0; (int64_t)blockIdx.x + (int64_t)(2 * jj + g) * step < p.num_tiles; jj++) {
            const int tile = blockIdx.x + (2 * jj + g) * step;
            const int slot = jj % CR_SLOTS, n = jj / CR_SLOTS;
            {
                // pull the slab of a tile CR_PREFETCH steps ahead into L2 with plain prefetches: a TMA load that waits
                // on DRAM holds up the TMA stores queued behind it (measured: 0.122 -> 0.155 ms for config C4)
                const int64_t pt = (int64_t)tile + (int64_t)(2 * CR_PREFETCH) * step;
                if (pt < p.num_tiles) {
                    const int ptile = (int)pt;
                    const int pb = ptile / per_img, pq = ptile - pb * per_img;
                    const int prb = pq / p.cv_col_blocks, pcb = pq - prb * p.cv_col_blocks;
                    const int px = max(0, (pcb * p.cv_tw * p.str_w - p.pad_w) * p.chan);
                    const int py = prb * p.cv_rows * p.str_h - p.pad_h;
                    const int row_floats = p.im_w * p.chan;
                    const int nfl = min(p.row_elems, row_floats - px);                 // floats of each slab row inside the image
                    const int lines = (nfl * 4 + 127) >> 7;
                    const float* img_b = p.images + (int64_t)pb * p.im_h * row_floats;
                    for (int i = lane; i < lines * p.cv_slab_rows; i += 32) {
                        const int sr = i / lines, ln = i - sr * lines;
                        const int iy = py + sr;
                        if (iy >= 0 && iy < p.im_h) {
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(img_b + (int64_t)iy * row_floats + px + ln * 32) : "memory");
                        }
                    }
                }
            }
            cr_mbar_wait(slab_empty(g, slot), (uint32_t)((n & 1) ^ 1));       // passes at once for the first CR_SLOTS tiles
            const int b = tile / per_img, q = tile - b * per_img;
            const int rb = q / p.cv_col_blocks, cb = q - rb * p.cv_col_blocks;
            const int x = (cb * p.cv_tw * p.str_w - p.pad_w) * p.chan - p.cv_xshift;      // multiple of 4 floats
            const int y = rb * p.cv_rows * p.str_h - p.pad_h;
            const uint32_t bar_f = slab_full(g, slot);
            cr_expect_tx_elect(bar_f, slot_tx);
            const uint32_t dst = ring_g + (uint32_t)(slot * p.cv_slot_elems) * 4u;
#pragma unroll 1
            for (int i = 0; i < p.cv_nbox; i++) {
                cr_tma_load3_elect(dst + (uint32_t)(i * p.cv_box_elems) * 4u, &tmapI, bar_f, x + p.cv_box_start[i], y, b);
            }
        }
    } else {
        // ===== store issuer: tiles in order; a staging tile goes back to its epilogue group once the store issued
        // after it has been committed and the older ones have finished reading shared memory =====
        const int per_img = p.cv_row_blocks * p.cv_col_blocks;
        const uint32_t stg_tile = (uint32_t)(p.NP >> 5) * 16384u;
        const int nbx = (p.N + 31) >> 5;
        int it = 0;
        int prev_ge = -1, prev_sb = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += step, it++) {
            const int ge = it & 1, jt = it >> 1;
            const int sb = p.stg_bufs == 2 ? (jt & 1) : 0;
            const int use = p.stg_bufs == 2 ? (jt >> 1) : jt;
            const int b = tile / per_img, qq = tile - b * per_img;
            const int rbk = qq / p.cv_col_blocks, cbk = qq - rbk * p.cv_col_blocks;
            cr_mbar_wait(stg_full(ge, sb), (uint32_t)(use & 1));
            const uint32_t stg = base + p.off_stg + (uint32_t)(ge * p.stg_bufs + sb) * stg_tile;
#pragma unroll 1
            for (int bx = 0; bx < nbx; bx++) {
                cr_tma_store4_elect(&tmapO, stg + (uint32_t)bx * 16384u, bx * 32, cbk * p.cv_tw, rbk * p.cv_rows, b);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (TRACE && lane == 0) cr_trace(p, it, 11);
            if (prev_ge >= 0) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // everything but the newest store has been read
                __syncwarp();
                cr_arrive_elect(stg_empty(prev_ge, prev_sb));
            }
            prev_ge = ge; prev_sb = sb;
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (prev_ge >= 0) { __syncwarp(); cr_arrive_elect(stg_empty(prev_ge, prev_sb)); }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    cr_fence_before();
    __syncthreads();
    if (warp == CR_EPI_WARP0) {
        cr_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace

namespace rb200 {

int encode_tensor_map_f32(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                          const uint32_t* box, bool swizzle128);

// Returns RB200_E_UNSUPPORTED (without setting an error the caller has to report) when the shape is outside this
// kernel's envelope; conv2d_tcgen05 (gemm_tcgen05_skinny.cu) then takes its shared-memory-staged kernels.
int conv2d_tcgen05_rows(int mode, const float* images, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                        int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                        int64_t dilation_h, int64_t dilation_w, int64_t pad_h, int64_t pad_w, int64_t out_h, int64_t out_w,
                        const float* W, int64_t filters, const float* bias, float* out)
{
    Context& c = ctx();
    static int env = -1;
    if (env < 0) { const char* e = getenv("RB200_CONV_ROWS_TMEM"); env = (e && e[0] == '0') ? 0 : 1; }
    if (!env) return RB200_E_UNSUPPORTED;
    if (mode != RB200_GEMM_TF32 && mode != RB200_GEMM_TF32X3) return RB200_E_UNSUPPORTED;
    const int planes = mode == RB200_GEMM_TF32X3 ? 2 : 1;
    const int64_t K = filter_h * filter_w * channels;
    const int64_t row_floats = im_w * channels;
    const int64_t run = (filter_w - 1) * dilation_w * channels + channels;
    auto gcd32 = [](int64_t v) { int g = 1; while (g < 32 && (v % (2 * g)) == 0) g *= 2; return g; };
    if (K < 1 || K > CR_BK || filters < 32 || filters > 128 || (filters % 4) != 0) return RB200_E_UNSUPPORTED;
    if ((row_floats % 4) != 0 || (reinterpret_cast<uintptr_t>(images) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) {
        return RB200_E_UNSUPPORTED;
    }
    if (run + 3 > 256 || gcd32(stride_w * channels) > 2 || im_h > 0x7fffffffLL || batches > 0x7fffffffLL) return RB200_E_UNSUPPORTED;

    CrParams p;
    memset(&p, 0, sizeof(p));
    p.N = (int)filters; p.K = (int)K; p.NP = (p.N + 31) / 32 * 32;
    p.W = W; p.ldW = filters; p.bias = bias;
    p.images = images; p.im_h = (int)im_h; p.im_w = (int)im_w;
    p.chan = (int)channels; p.str_h = (int)stride_h; p.str_w = (int)stride_w; p.pad_h = (int)pad_h; p.pad_w = (int)pad_w;
    // TMEM: acc_stages x NP accumulator columns + two A stages of planes*32 columns
    const int acc_cols = p.NP, a_cols = 2 * planes * CR_BK;
    p.acc_stages = (512 - a_cols) / acc_cols;
    if (p.acc_stages > CR_MAX_ACC) p.acc_stages = CR_MAX_ACC;
    if (p.acc_stages < 2) return RB200_E_UNSUPPORTED;
    uint32_t cols = (uint32_t)(p.acc_stages * acc_cols + a_cols), pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    p.tmem_cols = pow2;
    // tiling: <= 128 cells per tile, a block of out columns of one out row (wide images) or several whole out rows
    if (out_w >= CR_BM) {
        // balanced column blocks (222 columns -> 112 + 110, not 128 + 94): TMA spends its time on a box's clipped
        // rows too.  A multiple of 4 keeps every slab origin 16-byte aligned with one shift for all blocks.
        const int64_t nblk = rb_ceil_div(out_w, CR_BM);
        p.cv_tw = (int)((rb_ceil_div(out_w, nblk) + 3) / 4 * 4);
        p.cv_rows = 1;
    }
    else { p.cv_tw = (int)out_w; p.cv_rows = (int)(CR_BM / out_w); if (p.cv_rows > out_h) p.cv_rows = (int)out_h; }
    p.cv_slab_rows = (p.cv_rows - 1) * p.str_h + (int)((filter_h - 1) * dilation_h) + 1;
    if (p.cv_slab_rows > 256) return RB200_E_UNSUPPORTED;
    p.cv_xshift = (int)((((-pad_w * channels) % 4) + 4) % 4);
    const int row_elems = ((p.cv_tw - 1) * p.str_w + (int)((filter_w - 1) * dilation_w) + 1) * p.chan + p.cv_xshift;
    p.row_elems = row_elems;
    p.cv_bw = row_elems >= 256 ? 256 : (row_elems + 3) / 4 * 4;
    if (run + 3 > p.cv_bw) return RB200_E_UNSUPPORTED;
    int nbox = 1;
    p.cv_box_start[0] = 0;
    for (int cc = 0; cc < p.cv_tw; cc++) {
        const int x0 = cc * p.str_w * p.chan + p.cv_xshift;
        if (x0 + run > p.cv_box_start[nbox - 1] + p.cv_bw) {
            if (nbox == CR_MAX_BOXES) return RB200_E_UNSUPPORTED;
            p.cv_box_start[nbox++] = x0 / 4 * 4;
        }
    }
    p.cv_nbox = nbox;
    p.cv_box_elems = (p.cv_bw * p.cv_slab_rows + 31) / 32 * 32;
    p.cv_slot_elems = nbox * p.cv_box_elems;
    const int kwc = (int)(filter_w * channels);
    for (int k = 0; k < p.K; k++) {
        const int ky = k / kwc, rem = k - ky * kwc, kx = rem / p.chan, ch = rem - kx * p.chan;
        p.koff[k] = (int)((ky * dilation_h * p.cv_bw + kx * dilation_w * p.chan + ch) * 4);      // bytes
    }
    for (int k = p.K; k < CR_BK; k++) p.koff[k] = p.koff[0];
    p.cv_col_blocks = (int)rb_ceil_div(out_w, p.cv_tw);
    p.cv_row_blocks = (int)rb_ceil_div(out_h, p.cv_rows);
    const int64_t tiles = batches * p.cv_row_blocks * p.cv_col_blocks;
    if (tiles > 0x3fffffffLL) return RB200_E_UNSUPPORTED;
    p.num_tiles = (int)tiles;
    // shared memory: W planes | staging | ring | bias + barriers
    const size_t w_bytes = (size_t)planes * p.NP * 128;
    const size_t ring = (size_t)2 * CR_SLOTS * p.cv_slot_elems * 4;
    const size_t stg_tile = (size_t)(p.NP / 32) * 16384;
    size_t total = 0;
    for (p.stg_bufs = 2; p.stg_bufs >= 1; p.stg_bufs--) {
        p.off_stg = (uint32_t)w_bytes;                                  // multiple of 4096
        p.off_ring = (uint32_t)(p.off_stg + 2 * p.stg_bufs * stg_tile);
        p.off_misc = (uint32_t)((p.off_ring + ring + 127) / 128 * 128);
        total = p.off_misc + 512 + 512 + 1024;                          // bias, barriers, base alignment slack
        if (total <= (size_t)CR_SMEM_LIMIT) break;
    }
    if (p.stg_bufs < 1) return RB200_E_UNSUPPORTED;

    CUtensorMap mapI, mapO;
    {
        const uint64_t dims[3] = {(uint64_t)row_floats, (uint64_t)im_h, (uint64_t)batches};
        const uint64_t strides[2] = {(uint64_t)row_floats * 4, (uint64_t)row_floats * 4 * (uint64_t)im_h};
        const uint32_t box[3] = {(uint32_t)p.cv_bw, (uint32_t)p.cv_slab_rows, 1};
        if (encode_tensor_map_f32(&mapI, images, 3, dims, strides, box, false) != RB200_OK) return RB200_E_UNSUPPORTED;
    }
    {
        const uint64_t row = (uint64_t)filters * 4;
        const uint64_t dims[4] = {(uint64_t)filters, (uint64_t)out_w, (uint64_t)out_h, (uint64_t)batches};
        const uint64_t strides[3] = {row, row * (uint64_t)out_w, row * (uint64_t)out_w * (uint64_t)out_h};
        const uint32_t box[4] = {32, (uint32_t)p.cv_tw, (uint32_t)p.cv_rows, 1};
        if (encode_tensor_map_f32(&mapO, out, 4, dims, strides, box, true) != RB200_OK) return RB200_E_UNSUPPORTED;
    }
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_LIMIT));
        RB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_LIMIT));
        RB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_LIMIT));
        RB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_LIMIT));
        configured = true;
    }
    static int trace_env = -1;
    if (trace_env < 0) { const char* e = getenv("RB200_SKT_TRACE"); trace_env = (e && e[0] == '1') ? 1 : 0; }
    if (trace_env) {
        { const char* e = getenv("RB200_SKT_TRACE_T0"); p.trace_t0 = e ? atoi(e) : 0; }
        RB_CUDA(cudaMalloc(&p.trace, 64 * 16 * sizeof(long long)));
        RB_CUDA(cudaMemsetAsync(p.trace, 0, 64 * 16 * sizeof(long long), c.stream));
    }
    const int grid = p.num_tiles < c.sm_count ? p.num_tiles : c.sm_count;
    if (p.trace) {
        if (planes == 1) conv_rows_kernel<1, true><<<grid, CR_THREADS, total, c.stream>>>(p, mapI, mapO);
        else             conv_rows_kernel<2, true><<<grid, CR_THREADS, total, c.stream>>>(p, mapI, mapO);
    } else {
        if (planes == 1) conv_rows_kernel<1, false><<<grid, CR_THREADS, total, c.stream>>>(p, mapI, mapO);
        else             conv_rows_kernel<2, false><<<grid, CR_THREADS, total, c.stream>>>(p, mapI, mapO);
    }
    RB_LAUNCH_CHECK("conv_rows_kernel");
    c.last_gemm_path = planes == 1 ? "tcgen05_tf32_tmem_conv" : "tcgen05_tf32x3_tmem_conv";
    if (p.trace) {
        // diagnostic dump: per tile of CTA 0, event stamps relative to the first one (SM clocks)
        static long long host[64 * 16];
        RB_CUDA(cudaStreamSynchronize(c.stream));
        RB_CUDA(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
        RB_CUDA(cudaFree(p.trace));
        long long t0 = 0;
        for (int i = 0; i < 64 * 16; i++) if (host[i] && (!t0 || host[i] < t0)) t0 = host[i];
        fprintf(stderr, "conv rows trace (CTA 0): tile | producers: slab_ready lds_done a_empty_ok a_full_arrive | mma: t_empty_ok a_full_ok issued | "
                        "epilogue: top t_full_ok stg_free staged proxy_fence arrive | store_committed\n");
        static const int order[14] = {0, 1, 2, 3, -1, 4, 5, 6, -1, 13, 8, 9, 14, 15};
        for (int it = 0; it < 64; it++) {
            fprintf(stderr, "%3d |", it + p.trace_t0);
            for (int i = 0; i < 14; i++) {
                if (order[i] < 0) { fprintf(stderr, " |"); continue; }
                fprintf(stderr, " %7lld", host[it * 16 + order[i]] ? host[it * 16 + order[i]] - t0 : -1LL);
            }
            fprintf(stderr, " %7lld | %7lld\n", host[it * 16 + 10] ? host[it * 16 + 10] - t0 : -1LL, host[it * 16 + 11] ? host[it * 16 + 11] - t0 : -1LL);
        }
    }
    return RB200_OK;
}

}  // namespace rb200
