// gemm_tcgen05.cu -- float32 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators in TMEM, operands staged in shared memory by TMA).  sm_100a only.
//
// Replaces the inner loop of PhpBlas::gemm (src/Drivers/MatlibPHP/PhpBlas.php:927-947) for
// float32 when the operands are TMA-addressable (16-byte aligned bases, leading dimensions
// that are multiples of 4 elements); everything else goes to gemm_simt.cu.
//
// Two arithmetic modes (include/rindow_b200.h):
//   RB200_GEMM_TF32   : one tcgen05.mma per k-step on the raw float32 tiles (the tensor core
//                       reads the top 19 bits).  Fastest; max error ~1e-3 of max|C|.
//   RB200_GEMM_TF32X3 : A and B are first split into tf32 "hi" (cvt.rna) and "lo" (a - hi)
//                       planes by a streaming pre-pass (split_tf32_kernel); the GEMM then
//                       issues hi*hi into one TMEM accumulator and hi*lo + lo*hi into a second one
//                       (summed in the epilogue).
//                       ~fp32 accuracy (meets the reference's isclose bound, rtol 1e-4).
//
// Kernel anatomy (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.3d into a STAGES-deep smem ring,
//              mbarrier expect_tx / complete_tx
//   warp 1   : MMA issuer    -- one elected lane issues tcgen05.mma (M=128, N=BN, K=8 per
//              instruction, 4 per 32-float k-block), tcgen05.commit frees the smem slot and,
//              after the last k-block, publishes the accumulator
//   warp 2   : TMEM allocator (2 accumulators of BN columns: the epilogue of tile i overlaps
//              the mainloop of tile i+1)
//   warps 4-7: epilogue      -- tcgen05.ld 32x32b -> registers -> alpha/beta -> global
// Tiles are visited in a grouped (L2-friendly) order: 16 row-blocks share each B column-block.
//
// Shared-memory operand layouts (128-byte swizzle, 1024-byte aligned tiles):
//   K-major  (A NoTrans / B Trans): rows of 32 floats; TMA box {32, rows}; UMMA descriptor
//            SBO = 1024 (8 rows); advancing K by 8 floats = +32 bytes on the start address.
//   MN-major (A Trans / B NoTrans): boxes {32 mn, 32 k} of 4 KB, one per 32 columns, in the
//            128-byte swizzle with 32-BYTE atoms (TMA SWIZZLE_128B_ATOM_32B, UMMA layout
//            SWIZZLE_128B_BASE32B -- the only layout the tensor core takes for MN-major 32-bit
//            operands); descriptor LBO = 4096 (next 32 columns), SBO = 512 (next 4 k);
//            advancing K by 8 = +1024 bytes.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // floats per k-block = one 128-byte swizzle row
constexpr int TC_UMMA_K = 8;              // tf32: 32 bytes per instruction
constexpr int TC_THREADS = 256;
constexpr int TC_GROUP_M = 16;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        :: "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SW128).
// layout: 2 = SWIZZLE_128B (16-byte atoms; K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte atoms:
// the only layout the tensor core accepts for MN-major 32-bit operands, cf. CUTLASS
// sm100_smem_selector "for mn-major tf32 operands, SW128_32B is the only available smem layout").
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

// instruction descriptor for kind::tf32, fp32 accumulate (InstrDescriptor in the same header)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major)
{
    return (1u << 4)            // c_format = F32
         | (2u << 7)            // a_format = TF32
         | (2u << 10)           // b_format = TF32
         | ((uint32_t)a_mn_major << 15)
         | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
    int64_t M, N, K;
    float alpha, beta;
    float* C;
    int64_t ldC;
    int num_m_blocks, num_n_blocks, num_k_blocks;
    // UMMA descriptor fields of an MN-major operand tile (boxes of 32 mn x 32 k floats):
    uint32_t mn_lbo, mn_sbo, mn_layout;
    int group_m;                   // row-blocks per rasterisation group (CTA-pair kernel)
    // rb200_sgemm_allgather: peer copies of C (same row / column indexing, same ldC) that the CTA-pair kernel's
    // epilogue also stores every finished value to -- NVLink peer stores overlapping the remaining tiles
    int n_peers;
    float* peers[RB200_MAX_PEERS];
    // rb200_sgemm_allgather_mc: the multicast alias of C (NVLS: one multimem.st leaves this GPU once and the NVSwitch
    // replicates it into every rank's copy, this rank's included); nullptr = unicast peers above
    float* mc;
    // mc_pusher: the epilogue only stores its tile locally; warps 2 and 3 of each CTA read finished tiles back through L2
    // and multicast them as whole contiguous row segments (512 bytes per instruction), up to 4 tiles behind the epilogue
    int mc_pusher;                 // 0 = epilogue stores, 1 = warp 3 pushes, 2 = warps 2 and 3 push half a tile each
    int mc_push_rows;              // rows in flight per pusher warp (4 or 8)
    // strided batch (1-CTA kernels): tile index = (matrix, row block, column block); the third tensor-map coordinate
    // of an operand is matrix * bmul * PLANES + plane (bmul = 0 shares the operand across the batch)
    int batch, a_bmul, b_bmul;
    int64_t strideC;
    // CTA-pair kernel: every `sync_kb` k-blocks the producers of all pairs meet at a device-wide progress counter, so
    // the tiles of a wave walk K within one chunk of each other and the A / B panel slices they share are fetched
    // from DRAM once (0 = free running).  sync_cnt[0] = arrivals, sync_cnt[1] = finished CTAs (reset by the last one).
    // The spin needs every CTA of the wave resident at once: the grid is at most one CTA per SM, and the library
    // issues all of its kernels on ONE stream (rb200_set_stream orders a new stream behind the old one), so no other
    // kernel of this context can hold SMs while a wave waits (a kernel of another stream that does not itself wait
    // on this one only delays the wave).  Two processes on one GPU time-slice whole contexts; under MPS, where two
    // such GEMMs could interleave their CTAs, run with RB200_GEMM_SYNC_KB=0.
    int sync_kb;
    unsigned int* sync_cnt;
};

template <int BN, int PLANES, int STAGES>
struct TcSmem {
    static constexpr int A_PLANE = TC_BM * TC_BK * 4;         // 16 KB
    static constexpr int B_PLANE = BN * TC_BK * 4;            // 16 / 32 KB
    static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
    static constexpr int BARRIER_OFF = STAGES * STAGE;
    static constexpr int STG_LD = 36;                         // floats per staged epilogue row (32 + 4: conflict-free float4 rows)
    static constexpr int STG_OFF = BARRIER_OFF + 256;         // 4 epilogue warps x [32][STG_LD] floats
    static constexpr int TOTAL = STG_OFF + 4 * 32 * STG_LD * 4 + 1024;   // + alignment slack
};

// grouped tile order: TC_GROUP_M row-blocks share each column-block before moving on
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& mb, int& nb, int group_m = TC_GROUP_M)
{
    const int per_group = group_m * num_n;
    const int group = tile / per_group;
    const int first_m = group * group_m;
    const int gsz = min(num_m - first_m, group_m);
    const int in_group = tile - group * per_group;
    mb = first_m + in_group % gsz;
    nb = in_group / gsz;
}

template <int BN, bool A_MN, bool B_MN, int PLANES, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, TcParams p)
{
    using L = TcSmem<BN, PLANES, STAGES>;
    // two accumulator stages.  3xTF32: each stage holds a "big" accumulator (hi*hi) and a separate
    // "small" one (hi*lo + lo*hi), summed in the epilogue -- the TMEM adder truncates, and adding the
    // 2^-12-sized cross terms into the big accumulator would triple the number of truncating adds.
    constexpr int ACC_COLS = PLANES * BN;                     // columns per accumulator stage (256)
    constexpr int TMEM_COLS = 2 * ACC_COLS;                   // 512
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BARRIER_OFF;
    // barrier layout: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
    auto full_bar  = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar= [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_ptr_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tmem_ptr_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + L::BARRIER_OFF + 8 * (2 * STAGES + 4));

    const int tiles_per = p.num_m_blocks * p.num_n_blocks;
    const int num_tiles = tiles_per * p.batch;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int mb, nb;
                const int bt = tile / tiles_per;
                tile_coords(tile - bt * tiles_per, p.num_m_blocks, p.num_n_blocks, mb, nb);
                const int m0 = mb * TC_BM, n0 = nb * BN;
                const int a3 = bt * p.a_bmul * PLANES, b3 = bt * p.b_bmul * PLANES;
                for (int kb = 0; kb < p.num_k_blocks; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), (uint32_t)L::STAGE);
                    const uint32_t sA = smem_base + stage * L::STAGE;
                    const uint32_t sB = sA + PLANES * L::A_PLANE;
                    const int k0 = kb * TC_BK;
#pragma unroll
                    for (int pl = 0; pl < PLANES; pl++) {
                        if (!A_MN) {
                            tma_load_3d(sA + pl * L::A_PLANE, &tmapA, full_bar(stage), k0, m0, a3 + pl);
                        } else {
#pragma unroll
                            for (int j = 0; j < TC_BM / 32; j++)
                                tma_load_3d(sA + pl * L::A_PLANE + j * 4096, &tmapA, full_bar(stage), m0 + 32 * j, k0, a3 + pl);
                        }
                        if (!B_MN) {
                            tma_load_3d(sB + pl * L::B_PLANE, &tmapB, full_bar(stage), k0, n0, b3 + pl);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / 32; j++)
                                tma_load_3d(sB + pl * L::B_PLANE + j * 4096, &tmapB, full_bar(stage), n0 + 32 * j, k0, b3 + pl);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
                for (int kb = 0; kb < p.num_k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sA = smem_base + stage * L::STAGE;
                    const uint32_t sB = sA + PLANES * L::A_PLANE;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / TC_UMMA_K; kk++) {
                        // K-major: +32 bytes per UMMA_K inside the swizzle row; MN-major: +1024 (8 k-rows)
                        const uint32_t aoff = A_MN ? kk * 1024 : kk * 32;
                        const uint32_t boff = B_MN ? kk * 1024 : kk * 32;
                        const uint64_t a_hi = A_MN ? make_smem_desc(sA + aoff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                   : make_smem_desc(sA + aoff, 16, 1024, 2);
                        const uint64_t b_hi = B_MN ? make_smem_desc(sB + boff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                   : make_smem_desc(sB + boff, 16, 1024, 2);
                        const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                        if (PLANES == 1) {
                            tc_mma_tf32(tmem_d, a_hi, b_hi, idesc, first);
                        } else {
                            const uint64_t a_lo = A_MN ? make_smem_desc(sA + L::A_PLANE + aoff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                       : make_smem_desc(sA + L::A_PLANE + aoff, 16, 1024, 2);
                            const uint64_t b_lo = B_MN ? make_smem_desc(sB + L::B_PLANE + boff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                       : make_smem_desc(sB + L::B_PLANE + boff, 16, 1024, 2);
                            tc_mma_tf32(tmem_d + BN, a_hi, b_lo, idesc, first);   // cross terms -> small accumulator
                            tc_mma_tf32(tmem_d + BN, a_lo, b_hi, idesc, 1u);
                            tc_mma_tf32(tmem_d, a_hi, b_hi, idesc, first);        // hi*hi -> big accumulator
                        }
                    }
                    tc_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
                    if (kb == p.num_k_blocks - 1) tc_commit(tfull_bar(acc));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int mb, nb;
            const int bt = tile / tiles_per;
            tile_coords(tile - bt * tiles_per, p.num_m_blocks, p.num_n_blocks, mb, nb);
            const int64_t col0 = (int64_t)nb * BN;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ch++) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * 32);
                tc_ld_32x32b_x32(taddr, r);
                if (PLANES == 2) {
                    uint32_t rs[32];
                    tc_ld_32x32b_x32(taddr + BN, rs);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(rs[e]));
                } else {
                    tc_wait_ld();
                }
                const int64_t c0 = col0 + ch * 32;
                if (c0 < p.N) {                                   // warp-uniform
                    // lane = row holds 32 consecutive columns: transpose the warp's 32 x 32 block through a padded
                    // shared-memory tile so that every global access covers whole 128-byte row segments (4 rows x
                    // 128 bytes per instruction instead of 32 rows x 16 bytes) -- output-heavy short-K products
                    // (attention scores, conv) are bound by exactly these stores
                    float* stg = reinterpret_cast<float*>(smem_gen + L::STG_OFF) + (warp - 4) * (32 * L::STG_LD);
                    __syncwarp();
#pragma unroll
                    for (int v = 0; v < 8; v++) {
                        *reinterpret_cast<float4*>(stg + lane * L::STG_LD + 4 * v) =
                            make_float4(p.alpha * __uint_as_float(r[4 * v + 0]), p.alpha * __uint_as_float(r[4 * v + 1]),
                                        p.alpha * __uint_as_float(r[4 * v + 2]), p.alpha * __uint_as_float(r[4 * v + 3]));
                    }
                    __syncwarp();
                    const int64_t row_base = (int64_t)mb * TC_BM + q * 32;
                    const int cc = (lane & 7) * 4;
                    float* cmat = p.C + (int64_t)bt * p.strideC;
#pragma unroll
                    for (int pass = 0; pass < 8; pass++) {
                        const int rr = pass * 4 + (lane >> 3);
                        const int64_t grow = row_base + rr;
                        if (grow < p.M && c0 + cc < p.N) {
                            float4 o = *reinterpret_cast<const float4*>(stg + rr * L::STG_LD + cc);
                            float* dstp = cmat + grow * p.ldC + c0 + cc;
                            if (c0 + cc + 4 <= p.N) {
                                if (p.beta != 0.f) {
                                    const float4 old = *reinterpret_cast<const float4*>(dstp);
                                    o.x = fmaf(p.beta, old.x, o.x); o.y = fmaf(p.beta, old.y, o.y);
                                    o.z = fmaf(p.beta, old.z, o.z); o.w = fmaf(p.beta, old.w, o.w);
                                }
                                *reinterpret_cast<float4*>(dstp) = o;
                            } else {
                                const float oe[4] = {o.x, o.y, o.z, o.w};
                                for (int e = 0; e < 4; e++) {
                                    if (c0 + cc + e < p.N) dstp[e] = (p.beta != 0.f) ? fmaf(p.beta, dstp[e], oe[e]) : oe[e];
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// 2-CTA variant (1xTF32): a CTA pair (cluster of 2 on one TPC) computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2 (M = 256).  CTA r stages only ITS 128 rows of A and ITS 128 columns of B; the
// tensor cores of both SMs read both halves of B, so the L2 -> shared-memory traffic per MAC drops by a
// third against the 1-CTA kernel (64 KB instead of 96 KB per 256x256x32 block) and the 32 KB stages
// allow a 6-deep ring.  Each CTA keeps two 128 x 256 accumulators in its own TMEM and drains its own
// 128 rows.  Synchronisation: both CTAs' TMA loads complete_tx on the LEADER's full barrier (the leader
// posts the expect_tx for both), the leader's tcgen05.commit multicasts to the empty / tmem_full barriers
// of both CTAs, and the epilogue warps of both CTAs arrive on the leader's tmem_empty barrier.
// =====================================================================================================
// PLANES = 1 (TF32): pair tile 256 x 256, 32 KB stages, 6-deep ring, two 128 x 256 accumulators per CTA.
// PLANES = 2 (3xTF32): pair tile 256 x 128 of hi/lo planes, 48 KB stages, 4-deep ring; each accumulator
// stage holds a big (hi*hi) and a small (hi*lo + lo*hi) 128 x 128 accumulator, summed in the epilogue.
template <int PLANES>
struct Tc2Cfg {
    static constexpr int BN = PLANES == 1 ? 256 : 128;        // pair-tile columns (instruction N)
    static constexpr int HALF_N = BN / 2;                      // columns of B this CTA stages
    static constexpr int A_PLANE = TC_BM * TC_BK * 4;          // 16 KB: this CTA's 128 rows of A
    static constexpr int B_PLANE = HALF_N * TC_BK * 4;         // 16 / 8 KB
    static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
    static constexpr int STAGES = PLANES == 1 ? 6 : 4;
    static constexpr int BAR_OFF = STAGES * STAGE;
    static constexpr int STG_LD = 36;                          // floats per staged row (32 + 4: conflict-free float4 rows)
    static constexpr int STG_OFF = BAR_OFF + 256;              // 4 epilogue warps x [32][STG_LD] floats (peer stores)
    static constexpr int SMEM = STG_OFF + 4 * 32 * STG_LD * 4 + 1024;
    static constexpr int ACC_COLS = PLANES * BN;               // 256 TMEM columns per accumulator stage
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint32_t leader_bar,
                                                int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        :: "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar) : "memory");
}
// one 16-byte store to a multicast address: the switch writes it to the same offset of every bound allocation
__device__ __forceinline__ void multimem_st_f4(float* mc_addr, const float4& v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_f1(float* mc_addr, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool A_MN, bool B_MN, int PLANES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, TcParams p)
{
    using L = Tc2Cfg<PLANES>;
    constexpr int STAGES = L::STAGES;
    constexpr int BN = L::BN;                                 // pair tile: 256 x BN
    constexpr int ACC_COLS = L::ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFF;
    auto full_bar  = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar= [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_ptr_slot = bar_base + 8u * (2 * STAGES + 4);
    auto stored_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 5 + i); };       // epilogue -> pusher, 4 slots
    auto pushed_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 9 + i); };       // pusher -> epilogue, 4 slots
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader (issues the MMAs)
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
        for (int i = 0; i < 4; i++) { mbar_init(stored_bar(i), 4); mbar_init(pushed_bar(i), p.mc_pusher == 2 ? 2 : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tmem_ptr_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                        // barriers of BOTH CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFF + 8 * (2 * STAGES + 4));

    const int num_tiles = p.num_m_blocks * p.num_n_blocks;     // here: pair tiles (256 rows each)

    if (warp == 0) {
        // ===== TMA producer (one per CTA: its half of A and of B) =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            unsigned int sync_target = 0;                          // arrivals expected at this CTA's next meeting point
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int mb, nb;
                tile_coords(tile, p.num_m_blocks, p.num_n_blocks, mb, nb, p.group_m);
                const int m0 = mb * 256 + (int)rank * 128, n0 = nb * BN + (int)rank * L::HALF_N;
                // CTAs that have a tile in this wave (all of them except in the last, partial wave)
                const int wave_first = tile - pair;
                const unsigned int wave_ctas = 2u * (unsigned int)min(num_pairs, num_tiles - wave_first);
                for (int kb = 0; kb < p.num_k_blocks; kb++) {
                    if (p.sync_kb > 0 && kb > 0 && kb % p.sync_kb == 0) {
                        // bound the drift between the tiles of a wave: nobody starts chunk c + 1 of K before everybody
                        // has issued the loads of chunk c (a throttle, not a data dependency: relaxed atomics)
                        sync_target += wave_ctas;
                        atomicAdd(p.sync_cnt, 1u);
                        while (*reinterpret_cast<volatile unsigned int*>(p.sync_cnt) < sync_target) { }
                    }
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t leader_full = mapa_rank(full_bar(stage), 0);
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * L::STAGE);     // both CTAs' bytes
                    const uint32_t sA = smem_base + stage * L::STAGE;
                    const uint32_t sB = sA + PLANES * L::A_PLANE;
                    const int k0 = kb * TC_BK;
#pragma unroll
                    for (int pl = 0; pl < PLANES; pl++) {
                        if (!A_MN) {
                            tma_load_3d_2sm(sA + pl * L::A_PLANE, &tmapA, leader_full, k0, m0, pl);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                tma_load_3d_2sm(sA + pl * L::A_PLANE + j * 4096, &tmapA, leader_full, m0 + 32 * j, k0, pl);
                        }
                        if (!B_MN) {
                            tma_load_3d_2sm(sB + pl * L::B_PLANE, &tmapB, leader_full, k0, n0, pl);
                        } else {
#pragma unroll
                            for (int j = 0; j < L::HALF_N / 32; j++)
                                tma_load_3d_2sm(sB + pl * L::B_PLANE + j * 4096, &tmapB, leader_full, n0 + 32 * j, k0, pl);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only =====
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_tf32(256, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);     // both CTAs' epilogues drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
                for (int kb = 0; kb < p.num_k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);         // both halves of A and B have landed
                    tc_fence_after();
                    const uint32_t sA = smem_base + stage * L::STAGE;
                    const uint32_t sB = sA + PLANES * L::A_PLANE;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / TC_UMMA_K; kk++) {
                        const uint32_t aoff = A_MN ? kk * 1024 : kk * 32;
                        const uint32_t boff = B_MN ? kk * 1024 : kk * 32;
                        const uint64_t ad = A_MN ? make_smem_desc(sA + aoff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                 : make_smem_desc(sA + aoff, 16, 1024, 2);
                        const uint64_t bd = B_MN ? make_smem_desc(sB + boff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                 : make_smem_desc(sB + boff, 16, 1024, 2);
                        const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                        if (PLANES == 1) {
                            tc_mma_tf32_2sm(tmem_d, ad, bd, idesc, first);
                        } else {
                            const uint64_t a_lo = A_MN ? make_smem_desc(sA + L::A_PLANE + aoff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                       : make_smem_desc(sA + L::A_PLANE + aoff, 16, 1024, 2);
                            const uint64_t b_lo = B_MN ? make_smem_desc(sB + L::B_PLANE + boff, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                                       : make_smem_desc(sB + L::B_PLANE + boff, 16, 1024, 2);
                            tc_mma_tf32_2sm(tmem_d + BN, ad, b_lo, idesc, first);  // cross terms -> small accumulator
                            tc_mma_tf32_2sm(tmem_d + BN, a_lo, bd, idesc, 1u);
                            tc_mma_tf32_2sm(tmem_d, ad, bd, idesc, first);         // hi*hi -> big accumulator
                        }
                    }
                    tc_commit_2sm(empty_bar(stage));           // frees this stage in BOTH CTAs
                    if (kb == p.num_k_blocks - 1) tc_commit_2sm(tfull_bar(acc));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: this CTA's 128 rows of the pair tile =====
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            int mb, nb;
            tile_coords(tile, p.num_m_blocks, p.num_n_blocks, mb, nb, p.group_m);
            const int64_t row = (int64_t)mb * 256 + (int64_t)rank * 128 + q * 32 + lane;
            const int64_t col0 = (int64_t)nb * BN;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            float* crow = p.C + row * p.ldC;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ch++) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * 32);
                tc_ld_32x32b_x32(taddr, r);
                if (PLANES == 2) {
                    uint32_t rs[32];
                    tc_ld_32x32b_x32(taddr + BN, rs);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(rs[e]));
                } else {
                    tc_wait_ld();
                }
                const int64_t c0 = col0 + ch * 32;
                if (c0 < p.N) {                                   // warp-uniform
                    if (row_ok) {
                        if (c0 + 32 <= p.N) {
#pragma unroll
                            for (int v = 0; v < 8; v++) {
                                float4 o;
                                o.x = p.alpha * __uint_as_float(r[4 * v + 0]);
                                o.y = p.alpha * __uint_as_float(r[4 * v + 1]);
                                o.z = p.alpha * __uint_as_float(r[4 * v + 2]);
                                o.w = p.alpha * __uint_as_float(r[4 * v + 3]);
                                float4* dst = reinterpret_cast<float4*>(crow + c0) + v;
                                if (p.beta != 0.f) {
                                    const float4 old = *dst;
                                    o.x = fmaf(p.beta, old.x, o.x); o.y = fmaf(p.beta, old.y, o.y);
                                    o.z = fmaf(p.beta, old.z, o.z); o.w = fmaf(p.beta, old.w, o.w);
                                }
                                *dst = o;
                                r[4 * v + 0] = __float_as_uint(o.x); r[4 * v + 1] = __float_as_uint(o.y);
                                r[4 * v + 2] = __float_as_uint(o.z); r[4 * v + 3] = __float_as_uint(o.w);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; e++) {
                                if (c0 + e < p.N) {
                                    float o = p.alpha * __uint_as_float(r[e]);
                                    if (p.beta != 0.f) o = fmaf(p.beta, crow[c0 + e], o);
                                    crow[c0 + e] = o;
                                    r[e] = __float_as_uint(o);
                                }
                            }
                        }
                    }
                    if ((p.n_peers > 0 || p.mc != nullptr) && !p.mc_pusher) {
                        // all-gather leg: the warp's 32 x 32 block of final values goes through a padded smem tile so
                        // that every peer store instruction writes whole 128-byte row segments (NVLink-friendly),
                        // instead of 32 scattered 16-byte pieces
                        float* stg = reinterpret_cast<float*>(smem_gen + L::STG_OFF) + (warp - 4) * (32 * L::STG_LD);
                        __syncwarp();
#pragma unroll
                        for (int v = 0; v < 8; v++) {
                            *reinterpret_cast<float4*>(stg + lane * L::STG_LD + 4 * v) =
                                make_float4(__uint_as_float(r[4 * v + 0]), __uint_as_float(r[4 * v + 1]),
                                            __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                        }
                        __syncwarp();
                        const int64_t row_base = (int64_t)mb * 256 + (int64_t)rank * 128 + q * 32;
                        const int cc = (lane & 7) * 4;
#pragma unroll
                        for (int pass = 0; pass < 8; pass++) {
                            const int rr = pass * 4 + (lane >> 3);
                            const int64_t grow = row_base + rr;
                            if (grow < p.M && c0 + cc < p.N) {
                                const float4 val = *reinterpret_cast<const float4*>(stg + rr * L::STG_LD + cc);
                                const int64_t at = grow * p.ldC + c0 + cc;
                                if (p.mc != nullptr) {
                                    if (c0 + cc + 4 <= p.N) multimem_st_f4(p.mc + at, val);
                                    else {
                                        const float ve[4] = {val.x, val.y, val.z, val.w};
                                        for (int e = 0; e < 4; e++)
                                            if (c0 + cc + e < p.N) multimem_st_f1(p.mc + at + e, ve[e]);
                                    }
                                } else if (c0 + cc + 4 <= p.N) {
                                    for (int pj = 0; pj < p.n_peers; pj++)
                                        st_stream_f4(reinterpret_cast<float4*>(p.peers[pj] + at), val);
                                } else {
                                    const float ve[4] = {val.x, val.y, val.z, val.w};
                                    for (int pj = 0; pj < p.n_peers; pj++)
                                        for (int e = 0; e < 4; e++)
                                            if (c0 + cc + e < p.N) p.peers[pj][at + e] = ve[e];
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));   // leader's barrier, 8 arrivals
            if (p.mc_pusher) {
                // hand the tile to the pusher warp: its rows are in global memory (L2) once the fence has passed; slot
                // it % 4 may be reused when the pusher has finished the tile of four iterations ago
                __threadfence();
                __syncwarp();
                if (it >= 4) mbar_wait(pushed_bar(it & 3), (uint32_t)(((it >> 2) - 1) & 1));
                if (lane == 0) mbar_arrive(stored_bar(it & 3));
            }
            it++;
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (p.mc_pusher == 2 || (p.mc_pusher == 1 && warp == 3)) {
        // ===== pusher(s): warp 3 (and, with mc_pusher == 2, warp 2, which is idle between TMEM alloc and dealloc; then
        // 64 rows of the tile each): finished tiles of this CTA (128 rows x BN columns, 1 KB / 512 B contiguous per row)
        // are read back through L2 and multicast, 512 contiguous bytes per instruction, mc_push_rows rows in flight =====
        const int r_begin = p.mc_pusher == 2 ? (warp - 2) * 64 : 0;
        const int r_end = p.mc_pusher == 2 ? r_begin + 64 : 128;
        const int pr = p.mc_push_rows;
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, it++) {
            int mb, nb;
            tile_coords(tile, p.num_m_blocks, p.num_n_blocks, mb, nb, p.group_m);
            mbar_wait(stored_bar(it & 3), (uint32_t)((it >> 2) & 1));
            const int64_t row0 = (int64_t)mb * 256 + (int64_t)rank * 128;
            const int64_t col0 = (int64_t)nb * BN;
            int64_t ncols = p.N - col0;
            if (ncols > BN) ncols = BN;
            const bool vec = (ncols % 4 == 0);
            constexpr int PR = 8;                                 // upper bound of rows in flight per iteration
            for (int r = r_begin; r < r_end; r += pr) {
                float4 v[PR][BN / 128];
#pragma unroll
                for (int u = 0; u < PR; u++) {
                    const int64_t row = row0 + r + u;
#pragma unroll
                    for (int h = 0; h < BN / 128; h++) {
                        const int64_t cc = col0 + h * 128 + lane * 4;
                        if (u < pr && vec && row < p.M && cc - col0 < ncols) {
                            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                                         : "=f"(v[u][h].x), "=f"(v[u][h].y), "=f"(v[u][h].z), "=f"(v[u][h].w)
                                         : "l"(p.C + row * p.ldC + cc) : "memory");
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PR; u++) {
                    const int64_t row = row0 + r + u;
#pragma unroll
                    for (int h = 0; h < BN / 128; h++) {
                        const int64_t cc = col0 + h * 128 + lane * 4;
                        if (u < pr && row < p.M && cc - col0 < ncols) {
                            if (vec) {
                                multimem_st_f4(p.mc + row * p.ldC + cc, v[u][h]);
                            } else {
                                for (int e = 0; e < 4; e++)
                                    if (cc + e - col0 < ncols) {
                                        float x;
                                        asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(x) : "l"(p.C + row * p.ldC + cc + e) : "memory");
                                        multimem_st_f1(p.mc + row * p.ldC + cc + e, x);
                                    }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pushed_bar(it & 3));
        }
    }

    tc_fence_before();
    cluster_sync_all();                                        // nobody leaves while the peer still uses its smem / TMEM
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
    if (p.sync_kb > 0 && threadIdx.x == 0) {
        // the last CTA to finish puts the progress counter back to zero for the next launch on this stream
        if (atomicAdd(p.sync_cnt + 1, 1u) == gridDim.x - 1) {
            p.sync_cnt[0] = 0u;
            p.sync_cnt[1] = 0u;
            __threadfence();
        }
    }
}

// ---- hi/lo split pre-pass for the 3xTF32 mode ------------------------------------------------------
// planes[0] = tf32(x) (round to nearest, low 13 bits zero), planes[1] = x - planes[0] (exact in fp32).
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src,
                  float* __restrict__ dst, int64_t ld_dst, int64_t plane_stride, int64_t src_batch_stride)
{
    // blockIdx.z = matrix of a strided batch: its hi / lo planes sit at dst + z * 2 * plane_stride
    src += (int64_t)blockIdx.z * src_batch_stride;
    dst += (int64_t)blockIdx.z * 2 * plane_stride;
    const int64_t cv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 column
    if (cv * 4 >= cols) return;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(src + r * ld_src) + cv);
        float4 hi, lo;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); hi.x = __uint_as_float(t); lo.x = v.x - hi.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); hi.y = __uint_as_float(t); lo.y = v.y - hi.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); hi.z = __uint_as_float(t); lo.z = v.z - hi.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); hi.w = __uint_as_float(t); lo.w = v.w - hi.w;
        *(reinterpret_cast<float4*>(dst + r * ld_dst) + cv) = hi;
        *(reinterpret_cast<float4*>(dst + plane_stride + r * ld_dst) + cv) = lo;
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        }
    }
    return fn;
}

// 3-D map over a row-major [outer, inner] float matrix with `planes` copies `plane_stride` elements apart.
int make_map(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld, int planes,
             int64_t plane_stride, int box_inner, int box_outer, CUtensorMapSwizzle swizzle)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { rb200::set_error("cuTensorMapEncodeTiled is not available from the driver"); return RB200_E_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(planes > 1 ? plane_stride * 4 : ld * 4 * outer)};
    if (strides[1] % 16 != 0) strides[1] = (strides[1] + 15) / 16 * 16;
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rb200::set_error("cuTensorMapEncodeTiled failed (CUresult %d): inner=%lld outer=%lld ld=%lld planes=%d",
                         (int)r, (long long)inner, (long long)outer, (long long)ld, planes);
        return RB200_E_CUDA;
    }
    return RB200_OK;
}

template <int BN, bool A_MN, bool B_MN, int PLANES, int STAGES>
int launch_tf32(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p)
{
    rb200::Context& c = rb200::ctx();
    using L = TcSmem<BN, PLANES, STAGES>;
    auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, PLANES, STAGES>;
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    const int64_t tiles = (int64_t)p.num_m_blocks * p.num_n_blocks * p.batch;
    const int grid = tiles < c.sm_count ? (int)tiles : c.sm_count;
    kern<<<grid, TC_THREADS, L::TOTAL, c.stream>>>(mA, mB, p);
    RB_LAUNCH_CHECK("gemm_tf32_kernel");
    return RB200_OK;
}

template <int BN, int PLANES, int STAGES>
int launch_by_major(bool a_mn, bool b_mn, const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p)
{
    if (!a_mn && !b_mn) return launch_tf32<BN, false, false, PLANES, STAGES>(mA, mB, p);
    if (!a_mn &&  b_mn) return launch_tf32<BN, false, true,  PLANES, STAGES>(mA, mB, p);
    if ( a_mn && !b_mn) return launch_tf32<BN, true,  false, PLANES, STAGES>(mA, mB, p);
    return launch_tf32<BN, true, true, PLANES, STAGES>(mA, mB, p);
}

template <bool A_MN, bool B_MN, int PLANES>
int launch_2cta_t(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p)
{
    rb200::Context& c = rb200::ctx();
    auto kern = gemm_tf32_2cta_kernel<A_MN, B_MN, PLANES>;
    constexpr int SMEM = Tc2Cfg<PLANES>::SMEM;
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    const int max_pairs = c.sm_count / 2;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    kern<<<2 * pairs, TC_THREADS, SMEM, c.stream>>>(mA, mB, p);
    RB_LAUNCH_CHECK("gemm_tf32_2cta_kernel");
    return RB200_OK;
}

template <int PLANES>
int launch_2cta(bool a_mn, bool b_mn, const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p)
{
    if (!a_mn && !b_mn) return launch_2cta_t<false, false, PLANES>(mA, mB, p);
    if (!a_mn &&  b_mn) return launch_2cta_t<false, true, PLANES>(mA, mB, p);
    if ( a_mn && !b_mn) return launch_2cta_t<true, false, PLANES>(mA, mB, p);
    return launch_2cta_t<true, true, PLANES>(mA, mB, p);
}

}  // namespace

namespace {
// fallback of rb200_sgemm_allgather_mc for the kernel families whose epilogue does not store to the multicast alias:
// the finished stripe is read once and multicast with 16-byte multimem stores (rows of `cols` floats, stride ld)
__global__ void __launch_bounds__(256)
mc_push_rows_kernel(const float* __restrict__ src, float* __restrict__ mc, int64_t rows, int64_t cols, int64_t ld, int vec)
{
    const int64_t per_row = vec ? cols / 4 : cols;
    const int64_t total = rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row, c = i - r * per_row;
        if (vec) {
            const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(src + r * ld) + c);
            multimem_st_f4(mc + r * ld + 4 * c, v);
        } else {
            multimem_st_f1(mc + r * ld + c, src[r * ld + c]);
        }
    }
}
}  // namespace

namespace rb200 {

int mc_push_rows(const float* src, float* mc, int64_t rows, int64_t cols, int64_t ld)
{
    Context& c = ctx();
    const int vec = (cols % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(mc) & 15) == 0) ? 1 : 0;
    const int64_t total = rows * (vec ? cols / 4 : cols);
    int64_t grid = rb_ceil_div(total, 256);
    if (grid > (int64_t)c.sm_count * 8) grid = (int64_t)c.sm_count * 8;
    if (grid < 1) grid = 1;
    mc_push_rows_kernel<<<(unsigned)grid, 256, 0, c.stream>>>(src, mc, rows, cols, ld, vec);
    RB_LAUNCH_CHECK("mc_push_rows_kernel");
    return RB200_OK;
}

// General float32 tensor map (rank <= 5) for the skinny / implicit-GEMM kernels: image slabs (unswizzled
// loads) and output tiles (128-byte swizzled stores).  dims innermost first, strides in bytes for
// dimensions 1..rank-1, zero fill / clipping outside the tensor.
int encode_tensor_map_f32(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                          const uint32_t* box, bool swizzle128)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return RB200_E_CUDA; }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], estr[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) s[i] = strides[i];
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                    const_cast<float*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims[0..1]=%llu,%llu box[0..1]=%u,%u", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return RB200_E_CUDA;
    }
    return RB200_OK;
}

// exported for conv_tcgen05_tma.cu: the operand tensor map of the GEMM kernels (K-major: box {box_inner k, box_outer rows},
// 128-byte swizzle; MN-major: boxes {32 mn, 32 k}, 32-byte atoms), the hi / lo split pre-pass, and a general encoder
// with per-dimension element strides (strided boxes)
int make_gemm_map(void* map, const float* base, int64_t inner, int64_t outer, int64_t ld, int planes,
                  int64_t plane_stride, int box_inner, int box_outer, int mn_major)
{
    return make_map(reinterpret_cast<CUtensorMap*>(map), base, inner, outer, ld, planes, plane_stride, box_inner, box_outer,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int split_tf32_planes(const float* src, int64_t rows, int64_t cols, int64_t ld, float* dst, int64_t ld_dst,
                      int64_t plane_stride)
{
    Context& c = ctx();
    if (cols % 4 || ld % 4 || ld_dst % 4) {
        // the pre-pass moves float4 columns: pad-free callers only (conv: channels % 32 == 0; W: ld_dst padded to 4)
        if (ld % 4 || ld_dst % 4) { set_error("split_tf32_planes: leading dimensions must be multiples of 4"); return RB200_E_INVALID; }
    }
    const int64_t c4 = rb_ceil_div(cols, 4);
    dim3 grid((unsigned)rb_ceil_div(c4, 256), (unsigned)(rows < 65535 ? rows : 65535));
    split_tf32_kernel<<<grid, 256, 0, c.stream>>>(src, rows, cols, ld, dst, ld_dst, plane_stride, 0);
    RB_LAUNCH_CHECK("split_tf32_kernel");
    return RB200_OK;
}

int encode_tensor_map_f32_strided(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                                  const uint32_t* box, const uint32_t* elem_strides, bool swizzle128)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return RB200_E_CUDA; }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], estr[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; estr[i] = elem_strides[i]; }
    for (int i = 0; i + 1 < rank; i++) s[i] = strides[i];
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                    const_cast<float*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=%llu,%llu,%llu box=%u,%u,%u estr=%u,%u,%u", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  box[0], box[1], rank > 2 ? box[2] : 0, elem_strides[0], elem_strides[1], rank > 2 ? elem_strides[2] : 0);
        return RB200_E_CUDA;
    }
    return RB200_OK;
}

// Whether the tensor-core path can take this problem (TMA alignment rules).
bool gemm_tcgen05_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t ldA,
                           const float* B, int64_t ldB, const float* C, int64_t ldC)
{
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!al16(A) || !al16(B) || !al16(C)) return false;
    if (ldA % 4 || ldB % 4 || ldC % 4) return false;
    if (M < 1 || N < 1 || K < 1) return false;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) return false;
    if (rb_ceil_div(M, TC_BM) * rb_ceil_div(N, 128) > 0x3fffffffLL) return false;
    return true;
}

// A is rowsA x colsA row-major with ldA (transA: stored [K,M]); B likewise (transB: stored [N,K]).
int gemm_tcgen05(int mode, bool transA, bool transB, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* A, int64_t ldA, const float* B, int64_t ldB, float beta, float* C, int64_t ldC)
{
    Context& c = ctx();
    const bool a_mn = transA;          // stored [K, M]: M contiguous
    const bool b_mn = !transB;         // stored [K, N]: N contiguous
    const int planes = (mode == RB200_GEMM_TF32X3) ? 2 : 1;
    const int BN = planes == 1 ? 256 : 128;

    const float* srcA = A; const float* srcB = B;
    int64_t sldA = ldA, sldB = ldB, plsA = 0, plsB = 0;
    const int64_t rowsA = transA ? K : M, colsA = transA ? M : K;
    const int64_t rowsB = transB ? N : K, colsB = transB ? K : N;
    if (planes == 2) {
        // split pre-pass into the workspace: [2][rows][cols4] per operand
        const int64_t cA = rb_ceil_div(colsA, 4) * 4, cB = rb_ceil_div(colsB, 4) * 4;
        plsA = rowsA * cA; plsB = rowsB * cB;
        const size_t need = (size_t)(2 * plsA + 2 * plsB) * sizeof(float);
        int rc = ensure_workspace(need);
        if (rc != RB200_OK) return rc;
        float* wA = reinterpret_cast<float*>(c.workspace);
        float* wB = wA + 2 * plsA;
        {
            dim3 grid((unsigned)rb_ceil_div(cA / 4, 256), (unsigned)(rowsA < 65535 ? rowsA : 65535));
            split_tf32_kernel<<<grid, 256, 0, c.stream>>>(A, rowsA, colsA, ldA, wA, cA, plsA, 0);
            RB_LAUNCH_CHECK("split_tf32_kernel");
        }
        {
            dim3 grid((unsigned)rb_ceil_div(cB / 4, 256), (unsigned)(rowsB < 65535 ? rowsB : 65535));
            split_tf32_kernel<<<grid, 256, 0, c.stream>>>(B, rowsB, colsB, ldB, wB, cB, plsB, 0);
            RB_LAUNCH_CHECK("split_tf32_kernel");
        }
        srcA = wA; srcB = wB; sldA = cA; sldB = cB;
    }

    // MN-major tiles use the 128-byte swizzle with 32-byte atoms on both sides (TMA and UMMA descriptor):
    // k-rows of 128 bytes, 32-byte chunks XORed with (k mod 4); descriptor SBO = 512 (4 k-rows),
    // LBO = 4096 (next box of 32 columns).  Debug overrides: RB200_DBG_MN_{SWIZZLE,LBO,SBO,LAYOUT}.
    static const int dbg_sw = getenv("RB200_DBG_MN_SWIZZLE") ? atoi(getenv("RB200_DBG_MN_SWIZZLE")) : -1;
    static const int dbg_lbo = getenv("RB200_DBG_MN_LBO") ? atoi(getenv("RB200_DBG_MN_LBO")) : -1;
    static const int dbg_sbo = getenv("RB200_DBG_MN_SBO") ? atoi(getenv("RB200_DBG_MN_SBO")) : -1;
    static const int dbg_layout = getenv("RB200_DBG_MN_LAYOUT") ? atoi(getenv("RB200_DBG_MN_LAYOUT")) : -1;
    const CUtensorMapSwizzle k_sw = CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapSwizzle mn_sw = dbg_sw >= 0 ? (CUtensorMapSwizzle)dbg_sw : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;

    // The TMEM fp32 accumulator TRUNCATES on every accumulation (measured: profiles/, accumulation
    // probe -- all-positive inputs drift by -1.4e-5 at K=4096 and -9.3e-5 at K=16384, linearly in K).
    // That is below the 1xTF32 tolerance, but the 3xTF32 parity mode promises the reference's 1e-4
    // bound, so there K is processed in chunks of <= 2048 and the chunk results are added in the
    // epilogue with round-to-nearest FMAs (C = alpha*acc + 1*C): drift <= ~5e-6 per chunk, and the
    // chunk sums no longer accumulate it coherently.
    const int64_t KC = (planes == 2) ? 2048 : K;
    c.last_gemm_path = planes == 1 ? "tcgen05_tf32" : "tcgen05_tf32x3";
    for (int64_t kc = 0; kc < K; kc += KC) {
        const int64_t kl = (K - kc) < KC ? (K - kc) : KC;
        // K-major: inner = K, outer = M|N, box {32, rows}.  MN-major: inner = M|N, outer = K, box {32, 32}.
        const float* pa = a_mn ? srcA + kc * sldA : srcA + kc;
        const float* pb = b_mn ? srcB + kc * sldB : srcB + kc;
        CUtensorMap mA, mB;
        int rc;
        if (!a_mn) rc = make_map(&mA, pa, kl, M, sldA, planes, plsA, TC_BK, TC_BM, k_sw);
        else       rc = make_map(&mA, pa, M, kl, sldA, planes, plsA, 32, TC_BK, mn_sw);
        if (rc != RB200_OK) return rc;
        if (!b_mn) rc = make_map(&mB, pb, kl, N, sldB, planes, plsB, TC_BK, BN, k_sw);
        else       rc = make_map(&mB, pb, N, kl, sldB, planes, plsB, 32, TC_BK, mn_sw);
        if (rc != RB200_OK) return rc;

        TcParams p;
        p.M = M; p.N = N; p.K = kl; p.alpha = alpha; p.beta = (kc == 0) ? beta : 1.0f; p.C = C; p.ldC = ldC;
        p.num_m_blocks = (int)rb_ceil_div(M, TC_BM);
        p.num_n_blocks = (int)rb_ceil_div(N, BN);
        p.num_k_blocks = (int)rb_ceil_div(kl, TC_BK);
        p.mn_lbo = dbg_lbo >= 0 ? (uint32_t)dbg_lbo : 4096u;
        p.mn_sbo = dbg_sbo >= 0 ? (uint32_t)dbg_sbo : 512u;
        p.mn_layout = dbg_layout >= 0 ? (uint32_t)dbg_layout : 1u;
        p.group_m = TC_GROUP_M;
        p.n_peers = 0; p.mc = nullptr; p.mc_pusher = 0; p.mc_push_rows = 4; p.sync_kb = 0; p.sync_cnt = nullptr;
        p.batch = 1; p.a_bmul = 0; p.b_bmul = 0; p.strideC = 0;
        static const int use_2cta = getenv("RB200_GEMM_2CTA") ? atoi(getenv("RB200_GEMM_2CTA")) : 1;
        static const int use_2cta_x3 = getenv("RB200_GEMM_2CTA_X3") ? atoi(getenv("RB200_GEMM_2CTA_X3")) : 1;
        static const int use_small = getenv("RB200_GEMM_SMALL_TILES") ? atoi(getenv("RB200_GEMM_SMALL_TILES")) : 1;
        // Tile shape by estimated makespan (tiles are strided over the persistent CTAs, so a launch takes
        // ceil(tiles / units) tile times): the CTA pair wins whenever it fills the machine; problems with fewer
        // pair tiles than half the pairs (1024^3: 16 of 74) run 128 x 128 tiles on single CTAs instead -- four
        // times the tiles, a quarter of the work each.
        const bool pair_ok = (planes == 1 || use_2cta_x3) && use_2cta && M >= 256 && c.sm_count >= 2;
        const int64_t pairs = c.sm_count / 2;
        const int64_t tiles_pair = rb_ceil_div(M, 256) * rb_ceil_div(N, BN);
        const int64_t tiles_small = rb_ceil_div(M, 128) * rb_ceil_div(N, 128);
        const double t_pair = (double)rb_ceil_div(tiles_pair, pairs) * (256.0 * BN / 2.0);
        const double t_small = (double)rb_ceil_div(tiles_small, (int64_t)c.sm_count) * (128.0 * 128.0) * 1.15;
        const bool small_tiles = planes == 1 && use_small && c.n_gemm_peers == 0 && !c.gemm_mc && (!pair_ok || t_small < t_pair);
        // ... and 128 x 64 tiles when even those leave more than a third of the SMs idle (1024^3: 64 tiles of 128 x 128
        // on 148 SMs -> 128 tiles of 128 x 64; the launch is latency-bound, not tensor-bound, at that size)
        const int64_t tiles_64 = rb_ceil_div(M, 128) * rb_ceil_div(N, 64);
        const double t_64 = (double)rb_ceil_div(tiles_64, (int64_t)c.sm_count) * (128.0 * 64.0) * 1.3;
        const bool tiny_tiles = small_tiles && t_64 < t_small && (!pair_ok || t_64 < t_pair);
        // 3xTF32: the pair tile is 256 x 128 and the single-CTA tile 128 x 128 (the same work per SM); small
        // problems take 128 x 64 tiles
        const int64_t tiles_x3s = rb_ceil_div(M, 128) * rb_ceil_div(N, 64);
        const double t_x3s = (double)rb_ceil_div(tiles_x3s, (int64_t)c.sm_count) * (128.0 * 64.0) * 1.15;
        const bool small_x3 = planes == 2 && use_small && c.n_gemm_peers == 0 && !c.gemm_mc && t_x3s < t_pair;
        if (small_x3) {
            CUtensorMap mBs;
            if (!b_mn) rc = make_map(&mBs, pb, kl, N, sldB, planes, plsB, TC_BK, 64, k_sw);
            else       rc = make_map(&mBs, pb, N, kl, sldB, planes, plsB, 32, TC_BK, mn_sw);
            if (rc != RB200_OK) return rc;
            p.num_n_blocks = (int)rb_ceil_div(N, 64);
            rc = launch_by_major<64, 2, 4>(a_mn, b_mn, mA, mBs, p);
            c.last_gemm_path = "tcgen05_tf32x3_64";
        } else if (tiny_tiles) {
            CUtensorMap mBs;
            if (!b_mn) rc = make_map(&mBs, pb, kl, N, sldB, planes, plsB, TC_BK, 64, k_sw);
            else       rc = make_map(&mBs, pb, N, kl, sldB, planes, plsB, 32, TC_BK, mn_sw);
            if (rc != RB200_OK) return rc;
            p.num_n_blocks = (int)rb_ceil_div(N, 64);
            rc = launch_by_major<64, 1, 8>(a_mn, b_mn, mA, mBs, p);
            c.last_gemm_path = "tcgen05_tf32_64";
        } else if (small_tiles) {
            CUtensorMap mBs;
            if (!b_mn) rc = make_map(&mBs, pb, kl, N, sldB, planes, plsB, TC_BK, 128, k_sw);
            else       rc = make_map(&mBs, pb, N, kl, sldB, planes, plsB, 32, TC_BK, mn_sw);
            if (rc != RB200_OK) return rc;
            p.num_n_blocks = (int)rb_ceil_div(N, 128);
            rc = launch_by_major<128, 1, 6>(a_mn, b_mn, mA, mBs, p);
            c.last_gemm_path = "tcgen05_tf32_128";
        } else if (pair_ok) {
            // CTA-pair kernel: 256-row pair tiles; each CTA stages half of the pair tile's B columns
            CUtensorMap mB2;
            if (!b_mn) rc = make_map(&mB2, pb, kl, N, sldB, planes, plsB, TC_BK, BN / 2, k_sw);
            else       rc = make_map(&mB2, pb, N, kl, sldB, planes, plsB, 32, TC_BK, mn_sw);
            if (rc != RB200_OK) return rc;
            p.num_m_blocks = (int)rb_ceil_div(M, 256);
            // 74 pair tiles in flight as 8 pair-rows x ~9 column blocks: smallest A+B footprint per k-step,
            // and an A panel (8 x 256 rows) that stays L2-resident across the waves of its group
            static const int dbg_group = getenv("RB200_GEMM_GROUP") ? atoi(getenv("RB200_GEMM_GROUP")) : 8;
            p.group_m = dbg_group > 0 ? dbg_group : 8;
            if ((c.n_gemm_peers > 0 || c.gemm_mc) && kc + KC >= K) {     // the chunk that produces the final values
                p.n_peers = c.n_gemm_peers;
                for (int pj = 0; pj < c.n_gemm_peers; pj++) p.peers[pj] = c.gemm_peers[pj];
                p.mc = c.gemm_mc;
                // default: the pusher warp (N = 8, 16384^3: 2.03 ms with epilogue stores of 128-byte row segments, 1.92 ms
                // with the pusher's 512-byte contiguous stores; the contiguous multimem copy kernel alone moves a rank's
                // stripe into all eight copies at 772 GB/s of ingress per GPU -- profiles/nvls_gather_variants_r02.txt)
                // One pusher warp with four rows in flight is the measured optimum: two warps x eight rows read 2.06 ms
                // (the extra L2 read-back and store traffic competes with the operand loads).
                static const int mc_pusher = getenv("RB200_MC_PUSHER") ? atoi(getenv("RB200_MC_PUSHER")) : 1;
                static const int mc_push_rows = getenv("RB200_MC_PUSH_ROWS") ? atoi(getenv("RB200_MC_PUSH_ROWS")) : 4;
                p.mc_pusher = (mc_pusher > 0 && p.mc != nullptr && p.beta == 0.f) ? (mc_pusher >= 2 ? 2 : 1) : 0;
                p.mc_push_rows = mc_push_rows >= 8 ? 8 : 4;
                c.gemm_peers_written = true;
            }
            {
                // long reductions: keep the tiles of a wave within one K chunk (64 k-blocks = 2048) of each other (see
                // TcParams::sync_kb).  Free running, the 74 pair tiles of a wave drift apart over a long K loop until the
                // A / B panel slices they share have left L2 before the slower tiles ask for them: ncu at 16384^3 showed
                // 47.7 GB of DRAM reads per launch (L2 hit rate 62 %, SM clock 1.20 GHz under the power cap) against
                // 17.7 GB with the meeting points (75 %, 1.34 GHz) -- 631-662 -> 704-718 TFLOP/s; 8192^3 +1 %.
                // RB200_GEMM_SYNC_KB=0 switches it off.
                static const int sync_kb = getenv("RB200_GEMM_SYNC_KB") ? atoi(getenv("RB200_GEMM_SYNC_KB")) : 64;
                if (sync_kb > 0 && !c.gemm_sync) {               // per context: freed by rb200_shutdown with the device
                    RB_CUDA(cudaMalloc(&c.gemm_sync, 2 * sizeof(unsigned int)));
                    RB_CUDA(cudaMemsetAsync(c.gemm_sync, 0, 2 * sizeof(unsigned int), c.stream));
                }
                const int64_t kmin = getenv("RB200_GEMM_SYNC_MINK") ? atoll(getenv("RB200_GEMM_SYNC_MINK")) : 0;
                if (sync_kb > 0 && kl >= kmin) { p.sync_kb = sync_kb; p.sync_cnt = c.gemm_sync; }
            }
            rc = planes == 1 ? launch_2cta<1>(a_mn, b_mn, mA, mB2, p) : launch_2cta<2>(a_mn, b_mn, mA, mB2, p);
            c.last_gemm_path = planes == 1 ? "tcgen05_tf32_2cta" : "tcgen05_tf32x3_2cta";
        } else if (planes == 1) rc = launch_by_major<256, 1, 4>(a_mn, b_mn, mA, mB, p);
        else                    rc = launch_by_major<128, 2, 3>(a_mn, b_mn, mA, mB, p);
        if (rc != RB200_OK) return rc;
    }
    return RB200_OK;
}

// Strided batch on the single-CTA kernels: the batch is the third tensor-map dimension (times the hi / lo planes in
// 3xTF32 mode) and an extra digit of the tile index, so ONE persistent launch covers every matrix.  Operands with
// stride 0 are shared (their third dimension is just the planes).  Strides must keep 16-byte alignment.
bool gemm_tcgen05_batched_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t ldA, int64_t strideA,
                                   const float* B, int64_t ldB, int64_t strideB, const float* C, int64_t ldC,
                                   int64_t strideC, int64_t batch)
{
    if (!gemm_tcgen05_eligible(M, N, K, A, ldA, B, ldB, C, ldC)) return false;
    if (strideA % 4 || strideB % 4 || strideC % 4) return false;
    if (batch > 32768) return false;
    if (rb_ceil_div(M, TC_BM) * rb_ceil_div(N, 64) * batch > 0x3fffffffLL) return false;
    return M >= 64 && N >= 64 && K >= 32;
}

int gemm_tcgen05_batched(int mode, bool transA, bool transB, int64_t M, int64_t N, int64_t K, float alpha,
                         const float* A, int64_t ldA, int64_t strideA, const float* B, int64_t ldB, int64_t strideB,
                         float beta, float* C, int64_t ldC, int64_t strideC, int64_t batch)
{
    Context& c = ctx();
    const bool a_mn = transA, b_mn = !transB;
    const int planes = (mode == RB200_GEMM_TF32X3) ? 2 : 1;
    const int64_t rowsA = transA ? K : M, colsA = transA ? M : K;
    const int64_t rowsB = transB ? N : K, colsB = transB ? K : N;
    const int64_t nA = strideA ? batch : 1, nB = strideB ? batch : 1;       // distinct matrices per operand
    const float* srcA = A; const float* srcB = B;
    int64_t sldA = ldA, sldB = ldB, thirdA = strideA, thirdB = strideB;     // third-dimension stride (elements)
    if (planes == 2) {
        const int64_t cA = rb_ceil_div(colsA, 4) * 4, cB = rb_ceil_div(colsB, 4) * 4;
        const int64_t plsA = rowsA * cA, plsB = rowsB * cB;
        int rc = ensure_workspace((size_t)(2 * plsA * nA + 2 * plsB * nB) * sizeof(float));
        if (rc != RB200_OK) return rc;
        float* wA = reinterpret_cast<float*>(c.workspace);
        float* wB = wA + 2 * plsA * nA;
        {
            dim3 grid((unsigned)rb_ceil_div(cA / 4, 256), (unsigned)(rowsA < 1024 ? rowsA : 1024), (unsigned)nA);
            split_tf32_kernel<<<grid, 256, 0, c.stream>>>(A, rowsA, colsA, ldA, wA, cA, plsA, strideA);
            RB_LAUNCH_CHECK("split_tf32_kernel");
        }
        {
            dim3 grid((unsigned)rb_ceil_div(cB / 4, 256), (unsigned)(rowsB < 1024 ? rowsB : 1024), (unsigned)nB);
            split_tf32_kernel<<<grid, 256, 0, c.stream>>>(B, rowsB, colsB, ldB, wB, cB, plsB, strideB);
            RB_LAUNCH_CHECK("split_tf32_kernel");
        }
        srcA = wA; srcB = wB; sldA = cA; sldB = cB; thirdA = plsA; thirdB = plsB;
    }
    const CUtensorMapSwizzle k_sw = CU_TENSOR_MAP_SWIZZLE_128B, mn_sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    // tile width: the wider tile only when it still gives every SM work
    const int64_t mblocks = rb_ceil_div(M, TC_BM);
    int BN;
    if (planes == 1) BN = (N > 128 && mblocks * rb_ceil_div(N, 256) * batch >= c.sm_count) ? 256 : 128;
    else             BN = (N > 64 && mblocks * rb_ceil_div(N, 128) * batch >= c.sm_count) ? 128 : 64;
    const int64_t KC = (planes == 2) ? 2048 : K;
    for (int64_t kc = 0; kc < K; kc += KC) {
        const int64_t kl = (K - kc) < KC ? (K - kc) : KC;
        const float* pa = a_mn ? srcA + kc * sldA : srcA + kc;
        const float* pb = b_mn ? srcB + kc * sldB : srcB + kc;
        CUtensorMap mA, mB;
        int rc;
        // third dimension: planes x matrices; a count of 1 makes make_map ignore the stride
        const int cntA = (int)(planes * nA), cntB = (int)(planes * nB);
        if (!a_mn) rc = make_map(&mA, pa, kl, M, sldA, cntA, thirdA, TC_BK, TC_BM, k_sw);
        else       rc = make_map(&mA, pa, M, kl, sldA, cntA, thirdA, 32, TC_BK, mn_sw);
        if (rc != RB200_OK) return rc;
        if (!b_mn) rc = make_map(&mB, pb, kl, N, sldB, cntB, thirdB, TC_BK, BN, k_sw);
        else       rc = make_map(&mB, pb, N, kl, sldB, cntB, thirdB, 32, TC_BK, mn_sw);
        if (rc != RB200_OK) return rc;
        TcParams p;
        p.M = M; p.N = N; p.K = kl; p.alpha = alpha; p.beta = (kc == 0) ? beta : 1.0f; p.C = C; p.ldC = ldC;
        p.num_m_blocks = (int)mblocks;
        p.num_n_blocks = (int)rb_ceil_div(N, BN);
        p.num_k_blocks = (int)rb_ceil_div(kl, TC_BK);
        p.mn_lbo = 4096u; p.mn_sbo = 512u; p.mn_layout = 1u;
        p.group_m = TC_GROUP_M;
        p.n_peers = 0; p.mc = nullptr; p.mc_pusher = 0; p.mc_push_rows = 4; p.sync_kb = 0; p.sync_cnt = nullptr;
        p.batch = (int)batch; p.a_bmul = strideA ? 1 : 0; p.b_bmul = strideB ? 1 : 0; p.strideC = strideC;
        if (planes == 1) rc = BN == 256 ? launch_by_major<256, 1, 4>(a_mn, b_mn, mA, mB, p) : launch_by_major<128, 1, 6>(a_mn, b_mn, mA, mB, p);
        else             rc = BN == 128 ? launch_by_major<128, 2, 3>(a_mn, b_mn, mA, mB, p) : launch_by_major<64, 2, 4>(a_mn, b_mn, mA, mB, p);
        if (rc != RB200_OK) return rc;
    }
    c.last_gemm_path = planes == 1 ? "tcgen05_tf32_batched" : "tcgen05_tf32x3_batched";
    return RB200_OK;
}

}  // namespace rb200
