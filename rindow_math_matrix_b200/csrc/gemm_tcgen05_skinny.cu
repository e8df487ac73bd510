// gemm_tcgen05_skinny.cu -- tall-and-skinny float32 GEMM on tcgen05 for operands TMA cannot address:
// C[M,N] = alpha * A[M,K] * B[K,N] (+ beta*C), NoTrans/NoTrans, M large, N <= 128, K <= 128, ANY leading
// dimension / alignment.  This is the conv-as-GEMM shape of BASELINE config C4 -- cols [B*oh*ow, 27]
// x W [27, 64] -- whose row pitch (27 floats) rules out TMA.  The op is HBM-bound (AI ~ 9.5 flop/B),
// so the point of the tensor core here is to take the arithmetic off the FFMA pipe, not peak flops.
//
// Same reference semantics as gemm_tcgen05.cu (PhpBlas::gemm, PhpBlas.php:927-947), same two modes:
//   1 plane : operands rounded to tf32 with cvt.rna while staging (unbiased, unlike the raw-fp32
//             truncation of the TMA path)
//   2 planes: hi = rna(x), lo = x - hi computed in registers while staging (no pre-pass); hi*hi goes to
//             a "big" TMEM accumulator, hi*lo + lo*hi to a "small" one, summed in the epilogue.
//
// Persistent CTA per SM, 544 threads:
//   warps 0-7 : producers -- coalesced global loads of the A tile (128 rows x 32 k), software 128-byte
//               swizzle into the K-major UMMA layout, fence.proxy.async, mbarrier arrive (2 stages)
//   warp  8   : MMA issuer -- the whole warp runs the loop, one elected lane issues tcgen05.mma M=128, K=8
//               (3xTF32: A_hi x [B_hi | B_lo] with N = 2*NP, then A_lo x B_hi); commits free A stages / publish TMEM
//   warps 9-16: epilogue, two groups of four (group g drains accumulator stage g, tiles alternate) --
//               tcgen05.ld -> registers -> 128-byte swizzled staging tile -> TMA store (whole 128-byte lines,
//               ragged tiles clipped by the tensor map); unaligned C / beta != 0 / N < 32 take a shared-memory
//               transpose with plain 128-bit stores instead
// B (K x N, tiny) is staged once per CTA, transposed to K-major, all k-blocks resident ([k-block][plane][n]).
// The fused conv2d entry lives here for K in (32, 128] (cp.async slab ring + per-k gather); K <= 32 goes to
// conv_tcgen05_rows.cu, which keeps A in tensor memory.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int SKT_BM = 128;
constexpr int SKT_BK = 32;
constexpr int SKT_PROD_WARPS = 8;                    // two producer warps per scheduler: a lone warp issues at ~0.3 IPC
constexpr int SKT_PRODUCERS = SKT_PROD_WARPS * 32;
constexpr int SKT_MMA_WARP = SKT_PROD_WARPS;
constexpr int SKT_EPI_WARP0 = SKT_PROD_WARPS + 1;    // first of 2 x 4 epilogue warps; (warp & 3) covers the 4 TMEM lane quarters
constexpr int SKT_THREADS = (SKT_PROD_WARPS + 1 + 8) * 32;   // 544
constexpr int SKT_A_TILE = SKT_BM * SKT_BK * 4;      // 16 KB per plane per stage
constexpr int SKT_STAGES = 2;
constexpr int SKT_PREFETCH = 3;                      // raw A tiles in flight ahead of the one being staged
constexpr int SKT_RAW_SLOTS = SKT_PREFETCH + 1;
constexpr int SKT_RAW_STAGE = SKT_BM * SKT_BK * 4;   // 16 KB
constexpr int SKT_SMEM_LIMIT = 226 * 1024;

__device__ __forceinline__ uint32_t sk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sk_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void sk_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sk_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// warp-converged variants: every lane executes them, one elected lane issues
__device__ __forceinline__ void sk_mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void sk_commit_elect(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        :: "r"(bar) : "memory");
}
__device__ __forceinline__ void sk_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sk_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled tile: byte offset of element (row r, column kk < 32)
__device__ __forceinline__ uint32_t sk_swz(int r, int kk)
{
    return (uint32_t)(r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2)));
}
__device__ __forceinline__ uint64_t sk_desc(uint32_t smem_addr)          // K-major SW128: SBO = 1024
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// round to tf32, nearest with ties away from zero == cvt.rna.tf32.f32, which sm_100a expands to this add + mask
// plus an Inf/NaN test the split does not need: Inf stays Inf, a NaN stays NaN in at least one of hi / lo
__device__ __forceinline__ float sk_rna(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
struct SktParams {
    int64_t M;
    int N, K, NP, KB;              // NP = N rounded up to 32, KB = ceil(K / 32)
    float alpha, beta;
    const float* A; int64_t ldA;
    const float* B; int64_t ldB;
    float* C; int64_t ldC;
    int num_tiles;
    int a_contig;                  // ldA == K and A 16-byte aligned: the 128-row tile is one contiguous block
    unsigned k_magic;              // floor(2^32 / K) + 1
    uint32_t tmem_cols;
    int epi_groups;                // 1 or 2 groups of four epilogue warps (2 when the staging fits)
    int raw_slots;                 // slots of the raw cp.async ring (0 when the tile is not contiguous)
    // implicit-GEMM convolution (conv != 0): A is never materialised; row m = (b, oy, ox) and column
    // k = (ky, kx, c) of the virtual cols matrix are gathered straight from the channels-last images
    int conv;
    int im_h, im_w, chan, filt_w, str_h, str_w, dil_h, dil_w, pad_h, pad_w, out_h, out_w;
    const float* bias;             // optional [N] added in the epilogue (conv only)
    // conv tiling: an MMA tile = cv_rows out rows x cv_tw out columns of one image (cv_rows*cv_tw <= 128,
    // always one contiguous run of rows of the [B*oh*ow, N] output); its input slab is staged in smem
    int cv_tw, cv_rows, cv_col_blocks, cv_row_blocks;
    int cv_slab_rows, cv_width_x, cv_row_elems, cv_slab_elems;
    uint32_t ring_bytes;           // bytes of the raw / slab ring
    int tma_store;                 // epilogue stores C tiles with TMA (tmapC: 4-D [.., .., rows, N], 128-byte swizzle)
};

struct SktTile { int64_t m_start; int nvalid; int b, oy0, ox0; };

__device__ __forceinline__ SktTile skt_tile(const SktParams& p, int tile)
{
    SktTile t;
    if (!p.conv) {
        t.m_start = (int64_t)tile * SKT_BM;
        t.nvalid = (int)min((int64_t)SKT_BM, p.M - t.m_start);
        t.b = t.oy0 = t.ox0 = 0;
        return t;
    }
    const int per_img = p.cv_row_blocks * p.cv_col_blocks;
    t.b = tile / per_img;
    const int q = tile - t.b * per_img;
    const int rb = q / p.cv_col_blocks, cb = q - rb * p.cv_col_blocks;
    t.oy0 = rb * p.cv_rows;
    t.ox0 = cb * p.cv_tw;
    t.m_start = ((int64_t)t.b * p.out_h + t.oy0) * p.out_w + t.ox0;
    const int rows = min(p.cv_rows, p.out_h - t.oy0);
    // cv_rows > 1 only when cv_tw == out_w, so the valid cells are always one contiguous run
    t.nvalid = p.cv_rows > 1 ? rows * p.out_w : min(p.cv_tw, p.out_w - t.ox0);
    return t;
}

template <int PLANES>
__global__ void __launch_bounds__(SKT_THREADS, 1)
gemm_tf32_skinny_kernel(const __grid_constant__ SktParams p, const __grid_constant__ CUtensorMap tmapC)
{
    extern __shared__ uint8_t skt_raw[];
    const uint32_t base = (sk_smem_u32(skt_raw) + 1023u) & ~1023u;
    uint8_t* gen = skt_raw + (base - sk_smem_u32(skt_raw));
    // layout: A stages | B planes | epilogue staging | barriers
    const uint32_t a_stage_bytes = PLANES * SKT_A_TILE;
    const uint32_t b_plane_bytes = (uint32_t)p.KB * p.NP * 128;
    const uint32_t off_b = SKT_STAGES * a_stage_bytes;
    const uint32_t off_epi = off_b + PLANES * b_plane_bytes;
    const int epi_pitch = p.NP + 4;                                  // floats; keeps rows 16-byte aligned
    const uint32_t off_raw = (off_epi + (uint32_t)p.epi_groups * 4u * 32u * epi_pitch * 4u + 127u) & ~127u;
    const uint32_t off_bar = off_raw + p.ring_bytes;
    const uint32_t bar = base + ((off_bar + 7u) & ~7u);
    const uint32_t off_rinfo = ((off_bar + 7u) & ~7u) + 128u;            // [128][3] ints, conv only
    auto a_full  = [&](int s) { return bar + 8u * s; };
    auto a_empty = [&](int s) { return bar + 8u * (SKT_STAGES + s); };
    auto t_full  = [&](int a) { return bar + 8u * (2 * SKT_STAGES + a); };
    auto t_empty = [&](int a) { return bar + 8u * (2 * SKT_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar + 8u * (2 * SKT_STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ACC_COLS = PLANES * p.NP;

    if (warp == SKT_MMA_WARP && lane == 0) {
        for (int s = 0; s < SKT_STAGES; s++) { sk_mbar_init(a_full(s), SKT_PROD_WARPS); sk_mbar_init(a_empty(s), 1); }
        for (int a = 0; a < 2; a++) { sk_mbar_init(t_full(a), 1); sk_mbar_init(t_empty(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == SKT_EPI_WARP0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tmem_slot), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the A stages once: columns K..32*KB-1 of the last k-block and rows past M stay zero forever
    for (uint32_t o = threadIdx.x * 16u; o < SKT_STAGES * a_stage_bytes; o += SKT_THREADS * 16u) {
        *reinterpret_cast<float4*>(gen + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // stage B once: global [K][N] (row pitch ldB) -> K-major [kb][n][32 k] swizzled, tf32 hi (and lo) planes
    for (int e = threadIdx.x; e < p.KB * 32 * p.NP; e += SKT_THREADS) {
        const int k = e / p.NP, n = e - k * p.NP;
        float v = 0.f;
        if (k < p.K && n < p.N) v = __ldg(p.B + (int64_t)k * p.ldB + n);
        const float hi = sk_rna(v);
        // [k-block][plane][n][32 k]: the lo plane of a k-block directly follows its hi plane, so one UMMA
        // descriptor with N = 2*NP reads B_hi | B_lo as a single operand
        const uint32_t o = off_b + (uint32_t)(k >> 5) * (uint32_t)(PLANES * p.NP * 128) + sk_swz(n, k & 31);
        *reinterpret_cast<float*>(gen + o) = hi;
        if (PLANES == 2) *reinterpret_cast<float*>(gen + o + (uint32_t)(p.NP * 128)) = v - hi;
    }
    float* bias_s = reinterpret_cast<float*>(gen + off_rinfo + 1024u);         // [NP] bias (or zeros); the cp.async conv producer owns the first 1 KB
    if (p.tma_store) {
        for (int n = threadIdx.x; n < p.NP; n += SKT_THREADS) bias_s[n] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    sk_fence_before();
    __syncthreads();
    sk_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp < SKT_PROD_WARPS) {
        // ===== producers =====
        const int t = threadIdx.x;                                    // 0..SKT_PRODUCERS-1
        int stage = 0; uint32_t phase = 0;
        if (p.a_contig && p.KB == 1) {
            // Each 128-row tile is rows*K contiguous floats.  They are brought into a raw shared-memory
            // ring with 16-byte cp.async copies issued SKT_PREFETCH tiles ahead (each thread later reads
            // back exactly the chunks it copied, so cp.async.wait_group is the only synchronisation), then
            // swizzled / split into the UMMA tiles.  Depth 3 keeps ~40 KB of reads in flight per SM.
            constexpr int MAXV = (SKT_BM * SKT_BK / 4) / SKT_PRODUCERS;        // 8 chunks of 16 bytes per thread
            uint8_t* raw = gen + off_raw;
            auto issue = [&](int tile, int slot) {
                if (tile < p.num_tiles) {
                    const int64_t m0 = (int64_t)tile * SKT_BM;
                    const int rows = (int)min((int64_t)SKT_BM, p.M - m0);
                    const int nvec = (rows * p.K) >> 2;
                    const float4* src = reinterpret_cast<const float4*>(p.A + m0 * p.ldA);
#pragma unroll
                    for (int i = 0; i < MAXV; i++) {
                        const int q = t + i * SKT_PRODUCERS;
                        if (q < nvec) {
                            const uint32_t dst = base + off_raw + (uint32_t)slot * SKT_RAW_STAGE + (uint32_t)q * 16u;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src + q) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            const int step = gridDim.x;
            int slot_in = 0;
#pragma unroll
            for (int d = 0; d < SKT_PREFETCH; d++) { issue(blockIdx.x + d * step, slot_in); slot_in = (slot_in + 1) % SKT_RAW_SLOTS; }
            int slot_out = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += step) {
                issue(tile + SKT_PREFETCH * step, slot_in);
                slot_in = (slot_in + 1) % SKT_RAW_SLOTS;
                asm volatile("cp.async.wait_group %0;" :: "n"(SKT_PREFETCH) : "memory");
                const int64_t m0 = (int64_t)tile * SKT_BM;
                const int rows = (int)min((int64_t)SKT_BM, p.M - m0);
                const int nelem = rows * p.K, nvec = nelem >> 2;
                sk_mbar_wait(a_empty(stage), phase ^ 1);
                uint8_t* sa = gen + stage * a_stage_bytes;
                const float4* rbuf = reinterpret_cast<const float4*>(raw + slot_out * SKT_RAW_STAGE);
#pragma unroll
                for (int i = 0; i < MAXV; i++) {
                    const int q = t + i * SKT_PRODUCERS;
                    if (q < nvec) {
                        const float4 v4 = rbuf[q];
                        const float vals[4] = {v4.x, v4.y, v4.z, v4.w};
                        const int e = q * 4;
                        int r = p.K == 1 ? e : (int)__umulhi((unsigned)e, p.k_magic);
                        int kk = e - r * p.K;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float hi = sk_rna(vals[j]);
                            const uint32_t o = sk_swz(r, kk);
                            *reinterpret_cast<float*>(sa + o) = hi;
                            if (PLANES == 2) *reinterpret_cast<float*>(sa + SKT_A_TILE + o) = vals[j] - hi;
                            if (++kk == p.K) { kk = 0; r++; }
                        }
                    }
                }
                for (int e = (nvec << 2) + t; e < nelem; e += SKT_PRODUCERS) {         // <= 3 leftover elements
                    const int r = p.K == 1 ? e : (int)__umulhi((unsigned)e, p.k_magic);
                    const int kk = e - r * p.K;
                    const float x = __ldcs(p.A + m0 * p.ldA + e);
                    const float hi = sk_rna(x);
                    *reinterpret_cast<float*>(sa + sk_swz(r, kk)) = hi;
                    if (PLANES == 2) *reinterpret_cast<float*>(sa + SKT_A_TILE + sk_swz(r, kk)) = x - hi;
                }
                // rows past M of a partial last tile: zero them (an earlier full tile left data there)
                for (int e = rows * 32 + t; e < SKT_BM * 32; e += SKT_PRODUCERS) {
                    const uint32_t o = sk_swz(e >> 5, e & 31);
                    *reinterpret_cast<float*>(sa + o) = 0.f;
                    if (PLANES == 2) *reinterpret_cast<float*>(sa + SKT_A_TILE + o) = 0.f;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) sk_mbar_arrive(a_full(stage));
                if (++stage == SKT_STAGES) { stage = 0; phase ^= 1; }
                slot_out = (slot_out + 1) % SKT_RAW_SLOTS;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if (p.conv) {
            // Implicit GEMM.  Per tile the input slab (the cv_slab_rows image rows x cv_width_x columns that
            // feed its cells; zero outside the image) is brought into a shared-memory ring with 4-byte
            // cp.async copies issued SKT_PREFETCH tiles ahead; the A stage is then built from the slab:
            // thread = fixed k = (ky,kx,c) (fixed slab offset), walking the rows.  No global latency and no
            // bounds test in the build loop.
            int* roff = reinterpret_cast<int*>(gen + off_rinfo);                   // [128] slab offset of row r
            {
                const int rr = p.cv_rows > 1 ? t / p.cv_tw : 0;
                const int cc = t - rr * p.cv_tw;
                roff[t] = rr * p.str_h * p.cv_row_elems + cc * p.str_w * p.chan;
                asm volatile("bar.sync 1, %0;" :: "n"(SKT_PRODUCERS) : "memory");
            }
            float* ring = reinterpret_cast<float*>(gen + off_raw);
            // tile -> (image, row block, column block) without divisions: tiles advance by gridDim.x, so the
            // mixed-radix digits advance by constant amounts with carries
            struct TileIter { int b, rb, cb; };
            const int step = gridDim.x;
            const int per_img = p.cv_row_blocks * p.cv_col_blocks;
            const int st_b = step / per_img, st_q = step - st_b * per_img;
            const int st_rb = st_q / p.cv_col_blocks, st_cb = st_q - st_rb * p.cv_col_blocks;
            auto iter_init = [&](int tile) {
                TileIter it;
                it.b = tile / per_img;
                const int q = tile - it.b * per_img;
                it.rb = q / p.cv_col_blocks;
                it.cb = q - it.rb * p.cv_col_blocks;
                return it;
            };
            auto iter_next = [&](TileIter& it) {
                it.cb += st_cb; it.rb += st_rb; it.b += st_b;
                if (it.cb >= p.cv_col_blocks) { it.cb -= p.cv_col_blocks; it.rb++; }
                if (it.rb >= p.cv_row_blocks) { it.rb -= p.cv_row_blocks; it.b++; }
            };
            auto issue_slab = [&](int tile, const TileIter& it, int slot) {
                if (tile < p.num_tiles) {
                    const int oy0 = it.rb * p.cv_rows, ox0 = it.cb * p.cv_tw;
                    const int x_start = ox0 * p.str_w - p.pad_w, y_start = oy0 * p.str_h - p.pad_h;
                    const float* img_b = p.A + (int64_t)it.b * p.im_h * p.im_w * p.chan;
                    const uint32_t dst0 = base + off_raw + (uint32_t)slot * SKT_RAW_STAGE;
                    float* dstg = ring + slot * (SKT_RAW_STAGE / 4);
                    // per slab row the in-image part is one index interval [lo, hi): no per-element division
                    const int x_lo = max(0, -x_start), x_hi = max(x_lo, min(p.cv_width_x, p.im_w - x_start));
                    const int lo = x_lo * p.chan, hi = x_hi * p.chan;
                    for (int sl = 0; sl < p.cv_slab_rows; sl++) {
                        const int iy = y_start + sl;
                        const bool row_ok = iy >= 0 && iy < p.im_h;
                        const float* src = img_b + ((int64_t)iy * p.im_w + x_start) * p.chan;
                        const uint32_t drow = dst0 + 4u * (uint32_t)(sl * p.cv_row_elems);
                        float* grow = dstg + sl * p.cv_row_elems;
                        const bool vec = row_ok && lo == 0 &&
                                         ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((drow & 15u) == 0);
                        if (vec) {
                            // 16-byte copies for the aligned in-image part, 4-byte copies / zeros for the rest
                            const int nv4 = hi >> 2;
                            for (int q4 = t; q4 < nv4; q4 += SKT_PRODUCERS) {
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(drow + 16u * q4), "l"(src + 4 * q4) : "memory");
                            }
                            for (int idx = (nv4 << 2) + t; idx < p.cv_row_elems; idx += SKT_PRODUCERS) {
                                if (idx < hi) {
                                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(drow + 4u * idx), "l"(src + idx) : "memory");
                                } else {
                                    grow[idx] = 0.f;
                                }
                            }
                        } else {
                            for (int idx = t; idx < p.cv_row_elems; idx += SKT_PRODUCERS) {
                                if (row_ok && idx >= lo && idx < hi) {
                                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(drow + 4u * idx), "l"(src + idx) : "memory");
                                } else {
                                    grow[idx] = 0.f;
                                }
                            }
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            int slot_in = 0;
            TileIter it_in = iter_init(blockIdx.x);
#pragma unroll
            for (int d = 0; d < SKT_PREFETCH; d++) {
                issue_slab(blockIdx.x + d * step, it_in, slot_in);
                iter_next(it_in);
                slot_in = (slot_in + 1) % SKT_RAW_SLOTS;
            }
            int slot_out = 0;
            const int kwc = p.filt_w * p.chan;
            // this thread's column k = (ky,kx,c) of each k-block: fixed slab offset
            const int kk = t & 31;
            const int r0 = t >> 5;
            const uint32_t swc = ((uint32_t)(((kk >> 2) ^ (r0 & 7)) << 4)) | (uint32_t)((kk & 3) << 2);
            const int rstride = p.str_w * p.chan;
            const bool linear = p.cv_rows == 1;                                    // roff[r] == r * rstride
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += step) {
                issue_slab(tile + SKT_PREFETCH * step, it_in, slot_in);
                iter_next(it_in);
                slot_in = (slot_in + 1) % SKT_RAW_SLOTS;
                asm volatile("cp.async.wait_group %0;" :: "n"(SKT_PREFETCH) : "memory");
                asm volatile("bar.sync 1, %0;" :: "n"(SKT_PRODUCERS) : "memory");  // every producer's copies of this slab landed
                const float* slab = ring + slot_out * (SKT_RAW_STAGE / 4);
                for (int kb = 0; kb < p.KB; kb++) {
                    sk_mbar_wait(a_empty(stage), phase ^ 1);
                    uint8_t* sa = gen + stage * a_stage_bytes;
                    const int k = kb * 32 + kk;
                    if (k < p.K) {
                        // columns k >= K keep the zeros of the one-time fill; rows past the tile's valid cells
                        // get whatever the slab holds -- their accumulator rows are never stored
                        const int ky = k / kwc, rem = k - ky * kwc;
                        const int kx = rem / p.chan, c = rem - kx * p.chan;
                        const float* sk = slab + ky * p.dil_h * p.cv_row_elems + kx * p.dil_w * p.chan + c;
#pragma unroll 8
                        for (int i = 0; i < SKT_BM / SKT_PROD_WARPS; i++) {
                            const int r = r0 + SKT_PROD_WARPS * i;
                            const float v = sk[linear ? r * rstride : roff[r]];
                            const float hi = sk_rna(v);
                            const uint32_t o = (uint32_t)(r << 7) + swc;
                            *reinterpret_cast<float*>(sa + o) = hi;
                            if (PLANES == 2) *reinterpret_cast<float*>(sa + SKT_A_TILE + o) = v - hi;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) sk_mbar_arrive(a_full(stage));
                    if (++stage == SKT_STAGES) { stage = 0; phase ^= 1; }
                }
                asm volatile("bar.sync 1, %0;" :: "n"(SKT_PRODUCERS) : "memory");  // slab slot may be refilled now
                slot_out = (slot_out + 1) % SKT_RAW_SLOTS;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            // generic: a warp reads one row's 32 k (coalesced 128 bytes), 8 rows in flight per thread
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int64_t m0 = (int64_t)tile * SKT_BM;
                const int rows = (int)min((int64_t)SKT_BM, p.M - m0);
                for (int kb = 0; kb < p.KB; kb++) {
                    sk_mbar_wait(a_empty(stage), phase ^ 1);
                    uint8_t* sa = gen + stage * a_stage_bytes;
                    const int kk = t & 31;
                    const int k = kb * 32 + kk;
                    const float* col = p.A + m0 * p.ldA + k;
                    for (int r0 = t >> 5; r0 < SKT_BM; r0 += 8 * SKT_PROD_WARPS) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int r = r0 + SKT_PROD_WARPS * i;
                            v[i] = (r < rows && k < p.K) ? __ldcs(col + (int64_t)r * p.ldA) : 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int r = r0 + SKT_PROD_WARPS * i;
                            const float hi = sk_rna(v[i]);
                            const uint32_t o = sk_swz(r, kk);
                            *reinterpret_cast<float*>(sa + o) = hi;
                            if (PLANES == 2) *reinterpret_cast<float*>(sa + SKT_A_TILE + o) = v[i] - hi;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) sk_mbar_arrive(a_full(stage));
                    if (++stage == SKT_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == SKT_MMA_WARP) {
        // ===== MMA issuer =====
        // The whole warp runs the loop (uniform control flow keeps descriptors in uniform registers); one
        // elected lane issues.  3xTF32 takes two instructions per k-step: A_hi x [B_hi | B_lo] fills the big
        // (hi*hi) and the small (hi*lo) accumulator at once (N = 2*NP, adjacent TMEM columns), then
        // A_lo x B_hi adds into the small one.
        {
            const uint32_t idesc_n  = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(SKT_BM >> 4) << 24);
            const uint32_t idesc_2n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NP >> 2) << 17) | ((uint32_t)(SKT_BM >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                sk_mbar_wait(t_empty(acc), acc_phase ^ 1);
                sk_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
                for (int kb = 0; kb < p.KB; kb++) {
                    sk_mbar_wait(a_full(stage), phase);
                    sk_fence_after();
                    const uint32_t sA = base + stage * a_stage_bytes;
                    const uint32_t sB = base + off_b + (uint32_t)kb * (uint32_t)(PLANES * p.NP * 128);
#pragma unroll
                    for (int kk = 0; kk < SKT_BK / 8; kk++) {
                        const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                        const uint64_t a_hi = sk_desc(sA + kk * 32), b_hi = sk_desc(sB + kk * 32);
                        if (PLANES == 1) {
                            sk_mma_elect(tmem_d, a_hi, b_hi, idesc_n, first);
                        } else {
                            const uint64_t a_lo = sk_desc(sA + SKT_A_TILE + kk * 32);
                            sk_mma_elect(tmem_d, a_hi, b_hi, idesc_2n, first);
                            sk_mma_elect(tmem_d + p.NP, a_lo, b_hi, idesc_n, 1u);
                        }
                    }
                    sk_commit_elect(a_empty(stage));
                    if (kb == p.KB - 1) sk_commit_elect(t_full(acc));
                    if (++stage == SKT_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: two groups of 4 warps, group g drains accumulator stage g (tiles alternate) =====
        const int g = (warp - SKT_EPI_WARP0) >> 2;
        if (g < p.epi_groups && p.tma_store) {
            // TMA-store epilogue: TMEM -> registers -> alpha (+ bias) -> 128-byte swizzled staging tile of the
            // group ([NP/32 boxes][128 rows][32 floats]) -> one lane issues cp.async.bulk.tensor stores, which
            // clip ragged tiles (out-of-range rows / columns are not written) and write whole 128-byte lines.
            const int q = warp & 3;                                   // TMEM lane quarter of this warp
            const int r = q * 32 + lane;                              // row of the tile this thread owns
            const uint32_t stg = base + off_epi + (uint32_t)g * (uint32_t)(p.NP >> 5) * 16384u;
            const uint32_t row_addr = stg + (uint32_t)r * 128u;
            const uint32_t sw = (uint32_t)(r & 7);
            const int nbx = (p.N + 31) >> 5;
            const bool leader = q == 0 && lane == 0;
            uint32_t acc_phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                if (p.epi_groups == 2 && acc != g) continue;
                sk_mbar_wait(t_full(acc), acc_phase);
                if (p.epi_groups == 2 || acc == 1) acc_phase ^= 1;
                sk_fence_after();
                if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile is free again
                asm volatile("bar.sync %0, 128;" :: "r"(2 + g) : "memory");
                for (int ch = 0; ch < p.NP / 16; ch++) {
                    uint32_t rr[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * 16);
                    sk_ld16(taddr, rr);
                    if (PLANES == 2) {
                        uint32_t rs[16];
                        sk_ld16(taddr + p.NP, rs);
                        sk_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; e++) rr[e] = __float_as_uint(__uint_as_float(rr[e]) + __uint_as_float(rs[e]));
                    } else {
                        sk_wait_ld();
                    }
                    const float4* bv = reinterpret_cast<const float4*>(bias_s + ch * 16);
                    const uint32_t box_addr = row_addr + (uint32_t)(ch >> 1) * 16384u;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const float4 b4 = bv[v];
                        float4 o;
                        o.x = fmaf(p.alpha, __uint_as_float(rr[4 * v]), b4.x);
                        o.y = fmaf(p.alpha, __uint_as_float(rr[4 * v + 1]), b4.y);
                        o.z = fmaf(p.alpha, __uint_as_float(rr[4 * v + 2]), b4.z);
                        o.w = fmaf(p.alpha, __uint_as_float(rr[4 * v + 3]), b4.w);
                        const uint32_t a = box_addr + ((((uint32_t)(ch & 1) * 4u + (uint32_t)v) ^ sw) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                    }
                }
                sk_fence_before();
                __syncwarp();
                if (lane == 0) sk_mbar_arrive(t_empty(acc));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" :: "r"(2 + g) : "memory");
                if (leader) {
                    const SktTile ti = skt_tile(p, tile);
                    int c1, c2, c3;
                    if (p.conv) { c1 = ti.ox0; c2 = ti.oy0; c3 = ti.b; }
                    else { c1 = (int)ti.m_start; c2 = 0; c3 = 0; }
#pragma unroll 1
                    for (int bx = 0; bx < nbx; bx++) {
                        asm volatile(
                            "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                            :: "l"(reinterpret_cast<uint64_t>(&tmapC)), "r"(bx * 32), "r"(c1), "r"(c2), "r"(c3),
                               "r"(stg + (uint32_t)bx * 16384u) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else if (g < p.epi_groups) {
            const int q = warp & 3;                                   // TMEM lane quarter of this warp
            float* ebuf = reinterpret_cast<float*>(gen + off_epi) + (warp - SKT_EPI_WARP0) * 32 * epi_pitch;
            const int nv = p.NP >> 2;                                 // float4 per row
            const bool c_vec = ((p.ldC & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
            // lane -> (row, 16-byte piece) map of the transposed store; loop invariant when nv divides 32
            const bool inv = (32 % nv) == 0;
            const int rstep = inv ? 32 / nv : 0;
            const int lr = inv ? lane / nv : 0, lc4 = inv ? lane - (lane / nv) * nv : 0;
            uint32_t acc_phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                if (p.epi_groups == 2 && acc != g) continue;
                const SktTile ti = skt_tile(p, tile);
                const int64_t row0 = ti.m_start + q * 32;
                const int vrows = ti.nvalid - q * 32;                 // valid rows among this warp's 32
                sk_mbar_wait(t_full(acc), acc_phase);
                // with one group both stages are drained by it (phase flips every second tile); with two
                // groups each group sees only its own stage (phase flips every visit)
                if (p.epi_groups == 2 || acc == 1) acc_phase ^= 1;
                sk_fence_after();
                for (int ch = 0; ch < p.NP / 16; ch++) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * 16);
                    sk_ld16(taddr, r);
                    if (PLANES == 2) {
                        uint32_t rs[16];
                        sk_ld16(taddr + p.NP, rs);
                        sk_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(rs[e]));
                    } else {
                        sk_wait_ld();
                    }
                    float4* dst = reinterpret_cast<float4*>(ebuf + lane * epi_pitch + ch * 16);
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        dst[v] = make_float4(p.alpha * __uint_as_float(r[4 * v]), p.alpha * __uint_as_float(r[4 * v + 1]),
                                             p.alpha * __uint_as_float(r[4 * v + 2]), p.alpha * __uint_as_float(r[4 * v + 3]));
                    }
                }
                // all TMEM reads of this accumulator are done: hand it back before the global stores
                sk_fence_before();
                __syncwarp();
                if (lane == 0) sk_mbar_arrive(t_empty(acc));
                // transpose out of shared memory: consecutive lanes -> consecutive 16-byte pieces of a C row
                if (vrows <= 0) {
                    // this warp's 32 rows are all past the tile's valid cells
                } else if (inv && c_vec && p.NP == p.N && p.beta == 0.f &&
                    (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
                    // fast path: full tile, no beta -- 4 independent LDS.128 / STG.128 pairs per step
                    float* cbase = p.C + (row0 + lr) * p.ldC + lc4 * 4;
                    const float* sbase = ebuf + lr * epi_pitch + lc4 * 4;
                    const int steps = nv;                             // 32 rows / rstep rows per step
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias) + lc4);
                    for (int s0 = 0; s0 < steps; s0 += 4) {
                        float4 o[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (s0 + u < steps) o[u] = *reinterpret_cast<const float4*>(sbase + (s0 + u) * rstep * epi_pitch);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (s0 + u < steps && lr + (s0 + u) * rstep < vrows) {
                                o[u].x += bv.x; o[u].y += bv.y; o[u].z += bv.z; o[u].w += bv.w;
                                __stcs(reinterpret_cast<float4*>(cbase + (int64_t)(s0 + u) * rstep * p.ldC), o[u]);
                            }
                        }
                    }
                } else {
                    for (int idx = lane; idx < 32 * nv; idx += 32) {
                        const int rr = idx / nv, c4 = idx - rr * nv;
                        const int64_t grow = row0 + rr;
                        const int col = c4 * 4;
                        if (rr < vrows && col < p.N) {
                            float4 o = *reinterpret_cast<const float4*>(ebuf + rr * epi_pitch + col);
                            if (p.bias) {
                                o.x += __ldg(p.bias + col);
                                if (col + 1 < p.N) o.y += __ldg(p.bias + col + 1);
                                if (col + 2 < p.N) o.z += __ldg(p.bias + col + 2);
                                if (col + 3 < p.N) o.w += __ldg(p.bias + col + 3);
                            }
                            float* cp = p.C + grow * p.ldC + col;
                            if (c_vec && col + 4 <= p.N) {
                                if (p.beta != 0.f) {
                                    const float4 old = *reinterpret_cast<const float4*>(cp);
                                    o.x = fmaf(p.beta, old.x, o.x); o.y = fmaf(p.beta, old.y, o.y);
                                    o.z = fmaf(p.beta, old.z, o.z); o.w = fmaf(p.beta, old.w, o.w);
                                }
                                __stcs(reinterpret_cast<float4*>(cp), o);
                            } else {
                                const float vals[4] = {o.x, o.y, o.z, o.w};
                                for (int j = 0; j < 4 && col + j < p.N; j++) {
                                    cp[j] = p.beta != 0.f ? fmaf(p.beta, cp[j], vals[j]) : vals[j];
                                }
                            }
                        }
                    }
                }
                __syncwarp();                                         // ebuf is reused by the next tile
            }
        }
    }

    sk_fence_before();
    __syncthreads();
    if (warp == SKT_EPI_WARP0) {
        sk_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

size_t skt_smem_bytes_ring(int planes, int NP, int KB, int epi_groups, size_t ring_bytes)
{
    size_t a = (size_t)SKT_STAGES * planes * SKT_A_TILE;
    size_t b = (size_t)planes * KB * NP * 128;
    size_t e = (size_t)epi_groups * 4 * 32 * (NP + 4) * 4;
    return a + b + e + 128 + ring_bytes + 256 + 128 * 3 * 4 + 1024;
}
size_t skt_smem_bytes(int planes, int NP, int KB, int epi_groups, int raw_slots)
{
    return skt_smem_bytes_ring(planes, NP, KB, epi_groups, (size_t)raw_slots * SKT_RAW_STAGE);
}

bool skt_tma_store_enabled()
{
    static int env = -1;
    if (env < 0) { const char* e = getenv("RB200_SKT_TMA_STORE"); env = (e && e[0] == '0') ? 0 : 1; }
    return env != 0;
}

// one launcher for the GEMM and the conv producers
int skt_launch(int planes, const SktParams& p, const CUtensorMap& tmapC, size_t smem, const char* name)
{
    rb200::Context& c = rb200::ctx();
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(gemm_tf32_skinny_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKT_SMEM_LIMIT));
        RB_CUDA(cudaFuncSetAttribute(gemm_tf32_skinny_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKT_SMEM_LIMIT));
        configured = true;
    }
    const int grid = p.num_tiles < c.sm_count ? p.num_tiles : c.sm_count;
    if (planes == 1) gemm_tf32_skinny_kernel<1><<<grid, SKT_THREADS, smem, c.stream>>>(p, tmapC);
    else             gemm_tf32_skinny_kernel<2><<<grid, SKT_THREADS, smem, c.stream>>>(p, tmapC);
    RB_LAUNCH_CHECK(name);
    return RB200_OK;
}

}  // namespace

namespace rb200 {

int encode_tensor_map_f32(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                          const uint32_t* box, bool swizzle128);

bool gemm_tcgen05_skinny_eligible(int mode, int64_t M, int64_t N, int64_t K)
{
    if (mode != RB200_GEMM_TF32 && mode != RB200_GEMM_TF32X3) return false;
    if (M < 2048 || N < 1 || N > 128 || K < 1 || K > 128) return false;
    const int planes = mode == RB200_GEMM_TF32X3 ? 2 : 1;
    const int NP = (int)((N + 31) / 32 * 32), KB = (int)((K + 31) / 32);
    if (rb_ceil_div(M, SKT_BM) > 0x7fffffffLL) return false;
    return skt_smem_bytes(planes, NP, KB, 1, 0) <= SKT_SMEM_LIMIT;
}

int gemm_tcgen05_skinny(int mode, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                        const float* B, int64_t ldB, float beta, float* C, int64_t ldC)
{
    Context& c = ctx();
    const int planes = mode == RB200_GEMM_TF32X3 ? 2 : 1;
    SktParams p;
    p.M = M; p.N = (int)N; p.K = (int)K;
    p.NP = (int)((N + 31) / 32 * 32);
    p.KB = (int)((K + 31) / 32);
    p.alpha = alpha; p.beta = beta;
    p.A = A; p.ldA = ldA; p.B = B; p.ldB = ldB; p.C = C; p.ldC = ldC;
    p.conv = 0; p.bias = nullptr;
    p.cv_tw = p.cv_rows = p.cv_col_blocks = p.cv_row_blocks = p.cv_slab_rows = p.cv_width_x = p.cv_row_elems = p.cv_slab_elems = 0;
    p.im_h = p.im_w = p.chan = p.filt_w = p.str_h = p.str_w = p.dil_h = p.dil_w = p.pad_h = p.pad_w = p.out_h = p.out_w = 0;
    p.num_tiles = (int)rb_ceil_div(M, SKT_BM);
    p.a_contig = (ldA == K) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((SKT_BM * K) % 4 == 0);
    p.k_magic = (unsigned)((1ULL << 32) / (unsigned long long)K) + 1u;
    uint32_t cols = (uint32_t)(2 * planes * p.NP), pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    p.tmem_cols = pow2;
    p.epi_groups = skt_smem_bytes(planes, p.NP, p.KB, 2, 0) <= SKT_SMEM_LIMIT ? 2 : 1;
    p.raw_slots = 0;
    if (p.a_contig && p.KB == 1) {
        // the contiguous-tile fast path needs the raw ring; give up the second epilogue group before it
        if (skt_smem_bytes(planes, p.NP, p.KB, p.epi_groups, SKT_RAW_SLOTS) > SKT_SMEM_LIMIT) p.epi_groups = 1;
        if (skt_smem_bytes(planes, p.NP, p.KB, p.epi_groups, SKT_RAW_SLOTS) <= SKT_SMEM_LIMIT) p.raw_slots = SKT_RAW_SLOTS;
        else p.a_contig = 0;
    }
    const size_t smem = skt_smem_bytes(planes, p.NP, p.KB, p.epi_groups, p.raw_slots);
    p.ring_bytes = (uint32_t)p.raw_slots * SKT_RAW_STAGE;
    CUtensorMap mapC;
    memset(&mapC, 0, sizeof(mapC));
    p.tma_store = 0;
    if (skt_tma_store_enabled() && beta == 0.f && N >= 32 && (N % 4) == 0 && (ldC % 4) == 0 &&
        ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && M <= 0x7fffffffLL) {
        const uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, 1, 1};
        const uint64_t strides[3] = {(uint64_t)ldC * 4, (uint64_t)ldC * 4 * (uint64_t)M, (uint64_t)ldC * 4 * (uint64_t)M};
        const uint32_t box[4] = {32, SKT_BM, 1, 1};
        p.tma_store = encode_tensor_map_f32(&mapC, C, 4, dims, strides, box, true) == RB200_OK;    // else: plain stores
    }
    const int rc = skt_launch(planes, p, mapC, smem, "gemm_tf32_skinny_kernel");
    if (rc != RB200_OK) return rc;
    c.last_gemm_path = planes == 1 ? "tcgen05_tf32_skinny" : "tcgen05_tf32x3_skinny";
    return RB200_OK;
}

// Fused conv2d forward: out[b,oy,ox,f] = sum_{ky,kx,c} images[b, oy*sh+ky*dh-ph,
// ox*sw+kx*dw-pw, c] * W[(ky,kx,c), f] (+ bias[f]) == gemm(im2col2d(images) viewed [B*oh*ow, kh*kw*C], W) with
// the reference's default layouts (channels-last images, cols-channels-last), without ever writing cols.
bool conv2d_tcgen05_eligible(int mode, int64_t M, int64_t N, int64_t K)
{
    if (mode != RB200_GEMM_TF32 && mode != RB200_GEMM_TF32X3) return false;
    if (M < 1 || N < 1 || N > 128 || K < 1 || K > 128) return false;
    const int planes = mode == RB200_GEMM_TF32X3 ? 2 : 1;
    const int NP = (int)((N + 31) / 32 * 32), KB = (int)((K + 31) / 32);
    return skt_smem_bytes(planes, NP, KB, 1, 0) <= SKT_SMEM_LIMIT;
}

int conv2d_tcgen05_rows(int mode, const float* images, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                        int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                        int64_t dilation_h, int64_t dilation_w, int64_t pad_h, int64_t pad_w, int64_t out_h, int64_t out_w,
                        const float* W, int64_t filters, const float* bias, float* out);

int conv2d_tcgen05_tma(int mode, const float* images, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                       int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                       int64_t dilation_h, int64_t dilation_w, int64_t pad_h, int64_t pad_w, int64_t out_h, int64_t out_w,
                       const float* W, int64_t filters, const float* bias, float* out);

int conv2d_tcgen05(int mode, const float* images, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                   int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                   int64_t dilation_h, int64_t dilation_w, int64_t pad_h, int64_t pad_w, int64_t out_h, int64_t out_w,
                   const float* W, int64_t filters, const float* bias, float* out)
{
    // long reductions over channels % 32 == 0 (hidden layers): TMA-fed implicit GEMM (conv_tcgen05_tma.cu)
    if (filter_h * filter_w * channels > 32) {
        const int rc = conv2d_tcgen05_tma(mode, images, batches, im_h, im_w, channels, filter_h, filter_w, stride_h, stride_w,
                                          dilation_h, dilation_w, pad_h, pad_w, out_h, out_w, W, filters, bias, out);
        if (rc != RB200_E_UNSUPPORTED) return rc;
        if (!conv2d_tcgen05_eligible(mode, batches * out_h * out_w, filters, filter_h * filter_w * channels)) {
            set_error("conv2d_forward: shape not supported by the fused kernels (use im2col2d + gemm)");
            return RB200_E_UNSUPPORTED;
        }
    }
    // short reductions (K <= 32) go to the TMEM-operand kernel (conv_tcgen05_rows.cu) when it takes the shape
    {
        const int rc = conv2d_tcgen05_rows(mode, images, batches, im_h, im_w, channels, filter_h, filter_w, stride_h, stride_w,
                                           dilation_h, dilation_w, pad_h, pad_w, out_h, out_w, W, filters, bias, out);
        if (rc != RB200_E_UNSUPPORTED) return rc;
    }
    Context& c = ctx();
    const int planes = mode == RB200_GEMM_TF32X3 ? 2 : 1;
    SktParams p;
    p.M = batches * out_h * out_w;
    p.N = (int)filters; p.K = (int)(filter_h * filter_w * channels);
    p.NP = (p.N + 31) / 32 * 32;
    p.KB = (p.K + 31) / 32;
    p.alpha = 1.f; p.beta = 0.f;
    p.A = images; p.ldA = 0; p.B = W; p.ldB = filters; p.C = out; p.ldC = filters;
    p.num_tiles = (int)rb_ceil_div(p.M, SKT_BM);
    p.a_contig = 0;
    p.k_magic = (unsigned)((1ULL << 32) / (unsigned long long)p.K) + 1u;
    uint32_t cols = (uint32_t)(2 * planes * p.NP), pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    p.tmem_cols = pow2;
    p.conv = 1;
    p.bias = bias;
    p.im_h = (int)im_h; p.im_w = (int)im_w; p.chan = (int)channels; p.filt_w = (int)filter_w;
    p.str_h = (int)stride_h; p.str_w = (int)stride_w; p.dil_h = (int)dilation_h; p.dil_w = (int)dilation_w;
    p.pad_h = (int)pad_h; p.pad_w = (int)pad_w; p.out_h = (int)out_h; p.out_w = (int)out_w;
    // tiling: <= 128 cells per MMA tile, as a block of out columns of one out row (wide images) or as
    // several whole out rows (narrow feature maps)
    if (out_w >= SKT_BM) { p.cv_tw = SKT_BM; p.cv_rows = 1; }
    else { p.cv_tw = (int)out_w; p.cv_rows = (int)(SKT_BM / out_w); if (p.cv_rows > out_h) p.cv_rows = (int)out_h; }
    // the slab of inputs must fit one 16 KB slot of the cp.async ring
    for (;;) {
        p.cv_slab_rows = (p.cv_rows - 1) * p.str_h + (int)(filter_h - 1) * p.dil_h + 1;
        p.cv_width_x = (p.cv_tw - 1) * p.str_w + (int)(filter_w - 1) * p.dil_w + 1;
        p.cv_row_elems = p.cv_width_x * p.chan;
        p.cv_slab_elems = p.cv_slab_rows * p.cv_row_elems;
        if ((int64_t)p.cv_slab_elems * 4 <= SKT_RAW_STAGE) break;
        if (p.cv_rows > 1) { p.cv_rows--; continue; }
        if (p.cv_tw > 16) { p.cv_tw /= 2; continue; }
        set_error("conv2d_forward: the input slab of one tile exceeds the staging ring (use im2col2d + gemm)");
        return RB200_E_UNSUPPORTED;
    }
    p.raw_slots = SKT_RAW_SLOTS;
    p.ring_bytes = (uint32_t)SKT_RAW_SLOTS * SKT_RAW_STAGE;
    p.epi_groups = skt_smem_bytes(planes, p.NP, p.KB, 2, SKT_RAW_SLOTS) <= SKT_SMEM_LIMIT ? 2 : 1;
    if (skt_smem_bytes(planes, p.NP, p.KB, p.epi_groups, SKT_RAW_SLOTS) > SKT_SMEM_LIMIT) {
        set_error("conv2d_forward: shared memory budget exceeded for filters=%d, K=%d (use im2col2d + gemm)", p.N, p.K);
        return RB200_E_UNSUPPORTED;
    }
    p.cv_col_blocks = (int)rb_ceil_div(out_w, p.cv_tw);
    p.cv_row_blocks = (int)rb_ceil_div(out_h, p.cv_rows);
    const int64_t tiles = batches * p.cv_row_blocks * p.cv_col_blocks;
    if (tiles > 0x3fffffffLL) {
        set_error("conv2d_forward: problem too large for 32-bit tile indices");
        return RB200_E_UNSUPPORTED;
    }
    p.num_tiles = (int)tiles;
    const size_t smem = skt_smem_bytes_ring(planes, p.NP, p.KB, p.epi_groups, p.ring_bytes);
    CUtensorMap mapC;
    memset(&mapC, 0, sizeof(mapC));
    p.tma_store = 0;
    if (skt_tma_store_enabled() && filters >= 32 && (filters % 4) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        const uint64_t row = (uint64_t)filters * 4;
        const uint64_t dims[4] = {(uint64_t)filters, (uint64_t)out_w, (uint64_t)out_h, (uint64_t)batches};
        const uint64_t strides[3] = {row, row * (uint64_t)out_w, row * (uint64_t)out_w * (uint64_t)out_h};
        const uint32_t box[4] = {32, (uint32_t)p.cv_tw, (uint32_t)p.cv_rows, 1};
        p.tma_store = encode_tensor_map_f32(&mapC, out, 4, dims, strides, box, true) == RB200_OK;  // else: plain stores
    }
    const int rc = skt_launch(planes, p, mapC, smem, "gemm_tf32_skinny_kernel(conv)");
    if (rc != RB200_OK) return rc;
    c.last_gemm_path = planes == 1 ? "tcgen05_tf32_conv" : "tcgen05_tf32x3_conv";
    return RB200_OK;
}

}  // namespace rb200
