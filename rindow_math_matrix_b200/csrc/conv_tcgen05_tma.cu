// conv_tcgen05_tma.cu -- fused conv2d forward for LONG reductions (hidden layers: channels % 32 == 0, any kh x kw):
// an implicit GEMM whose A operand is fetched by TMA straight from the channels-last image tensor.
//
//   out[b,oy,ox,f] = sum_{ky,kx,c} images[b, oy*sh+ky*dh-ph, ox*sw+kx*dw-pw, c] * W[(ky,kx,c), f] (+ bias[f])
//   == gemm(im2col2d(images) viewed [B*oh*ow, K], W)   (PhpMath::im2col2d :2158-2314 + PhpBlas::gemm :927-947,
//   composed as LinearAlgebraCL::im2col2dclblast does, src/LinearAlgebraCL.php:5001-5144)
//
// For a fixed filter tap (ky,kx) and a block of 32 channels the rows of `cols` that an M tile needs are a BOX of the
// 5-D tensor (c, x, y, b, plane): 32 channels x tw output columns (every sw-th input column) x th output rows (every
// sh-th input row), shifted by the tap.  One cp.async.bulk.tensor.5d with element strides (1, sw, sh, 1, 1) lands
// that box in shared memory as 128 rows of 128 bytes in the 128-byte swizzle -- exactly the K-major UMMA operand --
// and TMA's out-of-bounds zero fill IS the `same` padding.  So the k loop of the ordinary TMA-fed tcgen05 GEMM
// (gemm_tcgen05.cu) simply walks (tap, channel block) and nothing else changes: no cols buffer (the reference path
// writes and re-reads K/C times the image), no producer warps building rows by hand (conv_tcgen05_rows.cu does that
// for K <= 32 where a row is shorter than one swizzle line).  W[(ky,kx,c), f] is the MN-major B operand as in
// `C = A.B`; 3xTF32 reads hi / lo planes of the image and of W written by the streaming split pre-pass.
//
// Tile = th x tw output cells of one image (tw = out_w and th = 128 / out_w for narrow feature maps, row segments of
// <= 128 cells otherwise) x BN filters; 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (TMEM -> registers (+ bias / + previous partial) -> smem transpose -> 128-byte row segments of
// out).  Reductions longer than 2048 are split by taps into launches that accumulate (the TMEM adder truncates,
// see gemm_tcgen05.cu).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace rb200 {
int make_gemm_map(void* map, const float* base, int64_t inner, int64_t outer, int64_t ld, int planes,
                  int64_t plane_stride, int box_inner, int box_outer, int mn_major);
int split_tf32_planes(const float* src, int64_t rows, int64_t cols, int64_t ld, float* dst, int64_t ld_dst,
                      int64_t plane_stride);
int encode_tensor_map_f32_strided(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                                  const uint32_t* box, const uint32_t* elem_strides, bool swizzle128);
}

namespace {

constexpr int CT_BM = 128;
constexpr int CT_BK = 32;
constexpr int CT_THREADS = 256;
constexpr int CT_STG_LD = 36;

__device__ __forceinline__ uint32_t ct_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ct_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ct_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void ct_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                               int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ct_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ct_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ct_ld32(uint32_t taddr, uint32_t (&r)[32])
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
__device__ __forceinline__ void ct_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool ct_elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t ct_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t ct_idesc(int M, int N, int b_mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct CtParams {
    int N;                          // filters
    int chan, cchunks;              // channels, channels / 32
    int filt_w;                     // kw (tap -> ky, kx)
    int tap_begin, tap_end;         // taps reduced by this launch
    int str_h, str_w, dil_h, dil_w, pad_h, pad_w;
    int out_h, out_w;
    int tw, th;                     // tile: th out rows x tw out columns (th * tw <= 128)
    int col_blocks, row_blocks, n_blocks, num_tiles;
    uint32_t a_bytes;               // bytes one A box delivers (th * tw rows of 128 bytes)
    int accumulate;                 // add to the values already in `out` (launches after the first tap range)
    const float* bias;              // added by the first launch only
    float* out;
};

template <int BN, int PLANES, int STAGES>
struct CtSmem {
    static constexpr int A_PLANE = CT_BM * CT_BK * 4;
    static constexpr int B_PLANE = BN * CT_BK * 4;
    static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
    static constexpr int BAR_OFF = STAGES * STAGE;
    static constexpr int STG_OFF = BAR_OFF + 256;
    static constexpr int TOTAL = STG_OFF + 4 * 32 * CT_STG_LD * 4 + 1024;
};

template <int BN, int PLANES, int STAGES>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const CtParams p)
{
    using L = CtSmem<BN, PLANES, STAGES>;
    constexpr int ACC_COLS = PLANES * BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ uint8_t ct_raw[];
    const uint32_t smem_base = (ct_smem_u32(ct_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFF;
    auto full_bar   = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar  = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar  = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_ptr_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_gen = ct_raw + (smem_base - ct_smem_u32(ct_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { ct_mbar_init(full_bar(s), 1); ct_mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { ct_mbar_init(tfull_bar(a), 1); ct_mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tmem_ptr_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // rows of the A stages a short tile never loads (th * tw < 128) must not hold NaN patterns: they feed accumulator
    // rows that are never stored, but keep them finite and deterministic
    if (p.a_bytes < (uint32_t)L::A_PLANE) {
        for (int i = threadIdx.x; i < STAGES * L::STAGE / 16; i += CT_THREADS)
            reinterpret_cast<int4*>(smem_gen)[i] = make_int4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    ct_fence_before();
    __syncthreads();
    ct_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFF + 8 * (2 * STAGES + 4));

    const int nk = (p.tap_end - p.tap_begin) * p.cchunks;            // k blocks of 32 in this launch
    const int tiles_per_img = p.row_blocks * p.col_blocks;

    auto tile_coords = [&](int tile, int& b, int& oy0, int& ox0, int& n0) {
        const int nb = tile % p.n_blocks;
        const int sp = tile / p.n_blocks;
        b = sp / tiles_per_img;
        const int q = sp - b * tiles_per_img;
        const int rb = q / p.col_blocks, cb = q - rb * p.col_blocks;
        oy0 = rb * p.th; ox0 = cb * p.tw; n0 = nb * BN;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (ct_elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int b, oy0, ox0, n0;
                tile_coords(tile, b, oy0, ox0, n0);
                const int x_base = ox0 * p.str_w - p.pad_w, y_base = oy0 * p.str_h - p.pad_h;
                int tap = p.tap_begin, cc = 0;
                int ky = tap / p.filt_w, kx = tap - ky * p.filt_w;
                for (int kb = 0; kb < nk; kb++) {
                    ct_mbar_wait(empty_bar(stage), phase ^ 1);
                    ct_mbar_expect_tx(full_bar(stage), (uint32_t)PLANES * (p.a_bytes + (uint32_t)L::B_PLANE));
                    const uint32_t sA = smem_base + stage * L::STAGE;
                    const uint32_t sB = sA + PLANES * L::A_PLANE;
                    const int k0 = tap * p.chan + cc * CT_BK;               // row of W
#pragma unroll
                    for (int pl = 0; pl < PLANES; pl++) {
                        ct_tma_load_5d(sA + pl * L::A_PLANE, &tmapA, full_bar(stage), cc * CT_BK, x_base + kx * p.dil_w,
                                       y_base + ky * p.dil_h, b, pl);
#pragma unroll
                        for (int j = 0; j < BN / 32; j++)
                            ct_tma_load_3d(sB + pl * L::B_PLANE + j * 4096, &tmapB, full_bar(stage), n0 + 32 * j, k0, pl);
                    }
                    if (++cc == p.cchunks) {
                        cc = 0; tap++;
                        if (++kx == p.filt_w) { kx = 0; ky++; }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (ct_elect_one()) {
            constexpr uint32_t idesc = ct_idesc(CT_BM, BN, 1);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                ct_mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                ct_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
                for (int kb = 0; kb < nk; kb++) {
                    ct_mbar_wait(full_bar(stage), phase);
                    ct_fence_after();
                    const uint32_t sA = smem_base + stage * L::STAGE;
                    const uint32_t sB = sA + PLANES * L::A_PLANE;
#pragma unroll
                    for (int kk = 0; kk < CT_BK / 8; kk++) {
                        const uint64_t a_hi = ct_desc(sA + kk * 32, 16, 1024, 2);              // K-major SW128
                        const uint64_t b_hi = ct_desc(sB + kk * 1024, 4096, 512, 1);           // MN-major SW128 / 32 B atoms
                        const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                        if (PLANES == 1) {
                            ct_mma_tf32(tmem_d, a_hi, b_hi, idesc, first);
                        } else {
                            const uint64_t a_lo = ct_desc(sA + L::A_PLANE + kk * 32, 16, 1024, 2);
                            const uint64_t b_lo = ct_desc(sB + L::B_PLANE + kk * 1024, 4096, 512, 1);
                            ct_mma_tf32(tmem_d + BN, a_hi, b_lo, idesc, first);    // cross terms -> small accumulator
                            ct_mma_tf32(tmem_d + BN, a_lo, b_hi, idesc, 1u);
                            ct_mma_tf32(tmem_d, a_hi, b_hi, idesc, first);         // hi*hi -> big accumulator
                        }
                    }
                    ct_commit(empty_bar(stage));
                    if (kb == nk - 1) ct_commit(tfull_bar(acc));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        const int cells = p.tw * p.th;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            int b, oy0, ox0, n0;
            tile_coords(tile, b, oy0, ox0, n0);
            ct_mbar_wait(tfull_bar(acc), acc_phase);
            ct_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ch++) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * 32);
                ct_ld32(taddr, r);
                if (PLANES == 2) {
                    uint32_t rs[32];
                    ct_ld32(taddr + BN, rs);
                    ct_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(rs[e]));
                } else {
                    ct_wait_ld();
                }
                const int c0 = n0 + ch * 32;
                if (c0 < p.N) {                                   // warp-uniform
                    float* stg = reinterpret_cast<float*>(smem_gen + L::STG_OFF) + (warp - 4) * (32 * CT_STG_LD);
                    __syncwarp();
#pragma unroll
                    for (int v = 0; v < 8; v++) {
                        *reinterpret_cast<float4*>(stg + lane * CT_STG_LD + 4 * v) =
                            make_float4(__uint_as_float(r[4 * v + 0]), __uint_as_float(r[4 * v + 1]),
                                        __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                    }
                    __syncwarp();
                    const int cc = (lane & 7) * 4;
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias && c0 + cc + 4 <= p.N) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + cc));
#pragma unroll
                    for (int pass = 0; pass < 8; pass++) {
                        const int rr = pass * 4 + (lane >> 3);
                        const int cell = q * 32 + rr;                       // row of the tile = cell (ty, tx)
                        const int ty = cell / p.tw, tx = cell - ty * p.tw;
                        const int oy = oy0 + ty, ox = ox0 + tx;
                        if (cell < cells && oy < p.out_h && ox < p.out_w && c0 + cc < p.N) {
                            float4 o = *reinterpret_cast<const float4*>(stg + rr * CT_STG_LD + cc);
                            float* dstp = p.out + (((int64_t)b * p.out_h + oy) * p.out_w + ox) * p.N + c0 + cc;
                            if (c0 + cc + 4 <= p.N) {
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                                if (p.accumulate) {
                                    const float4 old = *reinterpret_cast<const float4*>(dstp);
                                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                }
                                *reinterpret_cast<float4*>(dstp) = o;
                            } else {
                                const float oe[4] = {o.x, o.y, o.z, o.w};
                                for (int e = 0; e < 4; e++) {
                                    if (c0 + cc + e < p.N) {
                                        float v = oe[e] + (p.bias ? __ldg(p.bias + c0 + cc + e) : 0.f);
                                        if (p.accumulate) v += dstp[e];
                                        dstp[e] = v;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            ct_fence_before();
            __syncwarp();
            if (lane == 0) ct_mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    ct_fence_before();
    __syncthreads();
    if (warp == 2) {
        ct_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

template <int BN, int PLANES, int STAGES>
int ct_launch(const CUtensorMap& mA, const CUtensorMap& mB, const CtParams& p)
{
    rb200::Context& c = rb200::ctx();
    using L = CtSmem<BN, PLANES, STAGES>;
    auto kern = conv_tma_kernel<BN, PLANES, STAGES>;
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    const int grid = p.num_tiles < c.sm_count ? p.num_tiles : c.sm_count;
    kern<<<grid, CT_THREADS, L::TOTAL, c.stream>>>(mA, mB, p);
    RB_LAUNCH_CHECK("conv_tma_kernel");
    return RB200_OK;
}

}  // namespace

namespace rb200 {

// RB200_E_UNSUPPORTED = this kernel does not take the shape (the caller falls back); any other non-zero code is an error
int conv2d_tcgen05_tma(int mode, const float* images, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                       int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                       int64_t dilation_h, int64_t dilation_w, int64_t pad_h, int64_t pad_w, int64_t out_h, int64_t out_w,
                       const float* W, int64_t filters, const float* bias, float* out)
{
    Context& c = ctx();
    static const int enabled = getenv("RB200_CONV_TMA") ? atoi(getenv("RB200_CONV_TMA")) : 1;
    if (!enabled) return RB200_E_UNSUPPORTED;
    if (mode != RB200_GEMM_TF32 && mode != RB200_GEMM_TF32X3) return RB200_E_UNSUPPORTED;
    const int planes = mode == RB200_GEMM_TF32X3 ? 2 : 1;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (channels % CT_BK != 0 || filters % 4 != 0 || filters < 8) return RB200_E_UNSUPPORTED;
    if (!al16(images) || !al16(W) || !al16(out) || (bias && !al16(bias))) return RB200_E_UNSUPPORTED;
    if (batches > 0x7fffffffLL || im_h > 0x7fffffff || im_w > 0x7fffffff || filters > 0x7fffffff) return RB200_E_UNSUPPORTED;
    // tile: whole out rows of a narrow feature map, or balanced segments of <= 128 cells of one out row
    int tw, th;
    if (out_w >= CT_BM) {
        const int64_t segs = rb_ceil_div(out_w, (int64_t)CT_BM);
        tw = (int)rb_ceil_div(out_w, segs); th = 1;
    } else {
        tw = (int)out_w; th = (int)(CT_BM / out_w);
        if (th > out_h) th = (int)out_h;
    }
    // the box spans (tw - 1) * sw + 1 input columns and (th - 1) * sh + 1 input rows: TMA boxes are <= 256 per dimension
    while ((int64_t)(tw - 1) * stride_w + 1 > 256 && tw > 1) tw = (tw + 1) / 2;
    while ((int64_t)(th - 1) * stride_h + 1 > 256 && th > 1) th--;
    if ((int64_t)(tw - 1) * stride_w + 1 > 256 || stride_w > 8 || stride_h > 8) return RB200_E_UNSUPPORTED;

    const int64_t K = filter_h * filter_w * channels;
    const int64_t img_elems = batches * im_h * im_w * channels;
    const float* srcA = images; const float* srcW = W;
    int64_t plsA = 0, plsW = 0, ldW = filters;
    if (planes == 2) {
        // hi / lo planes of the images and of W in the workspace (streaming pre-pass, as rb200_sgemm's 3xTF32 mode)
        const int64_t cW = rb_ceil_div(filters, 4) * 4;
        plsA = img_elems; plsW = K * cW;
        int rc = ensure_workspace((size_t)(2 * plsA + 2 * plsW) * sizeof(float));
        if (rc != RB200_OK) return rc;
        float* wA = reinterpret_cast<float*>(c.workspace);
        float* wW = wA + 2 * plsA;
        const int64_t row = im_w * channels;
        rc = split_tf32_planes(images, batches * im_h, row, row, wA, row, plsA);
        if (rc != RB200_OK) return rc;
        rc = split_tf32_planes(W, K, filters, filters, wW, cW, plsW);
        if (rc != RB200_OK) return rc;
        srcA = wA; srcW = wW; ldW = cW;
    }
    // A: 5-D (c, x, y, b, plane), box 32 x ((tw-1)*sw+1) x ((th-1)*sh+1) x 1 x 1 walked with element strides (1, sw, sh)
    CUtensorMap mA, mB;
    {
        const uint64_t dims[5] = {(uint64_t)channels, (uint64_t)im_w, (uint64_t)im_h, (uint64_t)batches, (uint64_t)planes};
        const uint64_t strides[4] = {(uint64_t)channels * 4, (uint64_t)im_w * channels * 4,
                                     (uint64_t)im_h * im_w * channels * 4,
                                     (uint64_t)(planes > 1 ? plsA : img_elems) * 4};
        const uint32_t box[5] = {CT_BK, (uint32_t)((tw - 1) * stride_w + 1), (uint32_t)((th - 1) * stride_h + 1), 1, 1};
        const uint32_t es[5] = {1, (uint32_t)stride_w, (uint32_t)stride_h, 1, 1};
        int rc = encode_tensor_map_f32_strided(&mA, srcA, 5, dims, strides, box, es, true);
        if (rc != RB200_OK) return rc;
    }
    {
        int rc = make_gemm_map(&mB, srcW, filters, K, ldW, planes, plsW, 32, CT_BK, 1);
        if (rc != RB200_OK) return rc;
    }
    CtParams p;
    p.N = (int)filters; p.chan = (int)channels; p.cchunks = (int)(channels / CT_BK); p.filt_w = (int)filter_w;
    p.str_h = (int)stride_h; p.str_w = (int)stride_w; p.dil_h = (int)dilation_h; p.dil_w = (int)dilation_w;
    p.pad_h = (int)pad_h; p.pad_w = (int)pad_w; p.out_h = (int)out_h; p.out_w = (int)out_w;
    p.tw = tw; p.th = th;
    p.col_blocks = (int)rb_ceil_div(out_w, tw); p.row_blocks = (int)rb_ceil_div(out_h, th);
    p.a_bytes = (uint32_t)(tw * th) * 128u;
    p.out = out;
    // filters per tile: the widest accumulator that fits TMEM (2 stages x planes x BN columns <= 512)
    const int max_bn = planes == 1 ? 256 : 128;
    int BN = 64;
    while (BN < max_bn && BN < filters) BN *= 2;
    p.n_blocks = (int)rb_ceil_div(filters, BN);
    const int64_t tiles = batches * p.row_blocks * p.col_blocks * p.n_blocks;
    if (tiles > 0x3fffffffLL) return RB200_E_UNSUPPORTED;
    p.num_tiles = (int)tiles;
    // reductions of more than 2048 terms are accumulated by several launches (truncating TMEM adder, gemm_tcgen05.cu)
    const int taps = (int)(filter_h * filter_w);
    int taps_per = (int)(2048 / channels);
    if (taps_per < 1) taps_per = 1;
    if (planes == 1) taps_per = taps;
    for (int t0 = 0; t0 < taps; t0 += taps_per) {
        p.tap_begin = t0; p.tap_end = t0 + taps_per < taps ? t0 + taps_per : taps;
        p.accumulate = t0 > 0 ? 1 : 0;
        p.bias = t0 == 0 ? bias : nullptr;
        int rc;
        if (planes == 1) {
            rc = BN == 64 ? ct_launch<64, 1, 8>(mA, mB, p) : BN == 128 ? ct_launch<128, 1, 6>(mA, mB, p) : ct_launch<256, 1, 4>(mA, mB, p);
        } else {
            rc = BN == 64 ? ct_launch<64, 2, 4>(mA, mB, p) : ct_launch<128, 2, 3>(mA, mB, p);
        }
        if (rc != RB200_OK) return rc;
    }
    c.last_gemm_path = planes == 1 ? "tcgen05_tf32_tma_conv" : "tcgen05_tf32x3_tma_conv";
    return RB200_OK;
}

}  // namespace rb200
