// transpose.cu -- N-dimensional transpose B = permute(A, perm)  (HBM-bound data movement).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:2959-3025 + transCopy :3043-3067): with source strides
// s[d] (row-major over sourceShape) and target strides t[perm[d]] = prod(sourceShape[perm[e]], e > d), every source
// index (i_0 .. i_{n-1}) is copied:  B[offB + sum i_d * t[d]] = A[offA + sum i_d * s[d]].  A pure copy: bit-exact for
// every dtype (moved as 1 / 2 / 4 / 8-byte words).
//
// The host side first drops size-1 axes and merges source axes that stay adjacent and in order in the target, then
//   * last axis preserved -> rows kernel: one thread per 16-byte (or element) piece of the output, coordinates by
//     precomputed multiply-high divisions, reads and writes both run along the contiguous axis;
//   * last axis changes   -> tile kernel: 32 x 32 tiles over (source-fastest axis, target-fastest axis) through padded
//     shared memory, coalesced on both sides, the remaining axes decomposed once per CTA.
#include "common.cuh"

namespace {

constexpr int TR_MAXD = 8;

struct TrDims {
    int nd;                          // outer (batch) axes, in TARGET order, fastest last
    int64_t size[TR_MAXD];
    int64_t sstride[TR_MAXD];        // source stride of that axis
    int64_t tstride[TR_MAXD];        // target stride of that axis
};

// ---- last axis preserved: runs of `inner` contiguous elements on both sides --------------------------------------
// IDX = unsigned (all counts < 2^32: 32-bit divisions) or int64_t
template <typename W, int V, typename IDX>
__global__ void __launch_bounds__(256)
transpose_rows_kernel(TrDims d, int64_t inner_v64, int64_t total_v64, const W* __restrict__ A, W* __restrict__ B)
{
    struct alignas(sizeof(W) * V) Pack { W v[V]; };
    const IDX inner_v = (IDX)inner_v64, total_v = (IDX)total_v64;
    IDX size[TR_MAXD];
#pragma unroll
    for (int a = 0; a < TR_MAXD; a++) size[a] = a < d.nd ? (IDX)d.size[a] : (IDX)1;
    constexpr int U = 4;                                      // independent 16-byte loads in flight per thread
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < (int64_t)total_v; g0 += stride * U) {
        Pack p[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            // g enumerates the target in order: (outer axes..., inner piece)
            const int64_t g64 = g0 + u * stride;
            if (g64 < (int64_t)total_v) {
                const IDX g = (IDX)g64;
                IDX rest = g / inner_v;
                const IDX piece = g - rest * inner_v;
                int64_t src = 0;
#pragma unroll
                for (int a = TR_MAXD - 1; a >= 0; a--) {
                    if (a < d.nd) {
                        const IDX q = rest / size[a];
                        src += (int64_t)(rest - q * size[a]) * d.sstride[a];
                        rest = q;
                    }
                }
                p[u] = *reinterpret_cast<const Pack*>(A + src + (int64_t)piece * V);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t g64 = g0 + u * stride;
            if (g64 < (int64_t)total_v) *reinterpret_cast<Pack*>(B + g64 * V) = p[u];
        }
    }
}

// ---- last axis changes: 32 x 32 tiles over (x = source-fastest axis, y = target-fastest axis) --------------------
// source: x contiguous, y at stride sy;  target: y contiguous, x at stride tx.  blockIdx.z (+ gridDim striding) walks
// the remaining axes.
template <typename W>
__global__ void __launch_bounds__(256)
transpose_tile_kernel(TrDims d, int64_t nx, int64_t ny, int64_t sy, int64_t tx, int64_t batches,
                      const W* __restrict__ A, W* __restrict__ B)
{
    __shared__ W tile[32][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;         // 32 x 8 threads
    const int64_t x0 = (int64_t)blockIdx.x * 32, y0 = (int64_t)blockIdx.y * 32;
    for (int64_t bt = blockIdx.z; bt < batches; bt += gridDim.z) {
        int64_t rest = bt, src = 0, dst = 0;
#pragma unroll 1
        for (int a = d.nd - 1; a >= 0; a--) {
            const int64_t q = rest / d.size[a];
            const int64_t i = rest - q * d.size[a];
            src += i * d.sstride[a];
            dst += i * d.tstride[a];
            rest = q;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t y = y0 + ly + 8 * j, x = x0 + lx;
            if (y < ny && x < nx) tile[ly + 8 * j][lx] = A[src + y * sy + x];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t x = x0 + ly + 8 * j, y = y0 + lx;
            if (x < nx && y < ny) B[dst + x * tx + y] = tile[lx][ly + 8 * j];
        }
        __syncthreads();
    }
}

// 64 x 64 tiles with 16-byte accesses on both sides (4- and 8-byte elements; every extent, stride and base a multiple
// of the pack width).  A linear grid walks (batch, tile row, tile column).
template <typename W>
__global__ void __launch_bounds__(256)
transpose_tile_vec_kernel(TrDims d, int64_t nx, int64_t ny, int64_t sy, int64_t tx, int64_t batches,
                          const W* __restrict__ A, W* __restrict__ B)
{
    constexpr int V = 16 / (int)sizeof(W);
    constexpr int TS = 64;
    constexpr int VPR = TS / V;                  // packs per tile row
    constexpr int RPP = 256 / VPR;               // tile rows per pass
    constexpr int PASSES = TS / RPP;
    union Pack { int4 raw; W v[V]; };
    __shared__ W tile[TS][TS + 1];
    const int xv = threadIdx.x % VPR, r0 = threadIdx.x / VPR;
    const int64_t tiles_x = (nx + TS - 1) / TS, tiles_y = (ny + TS - 1) / TS;
    const int64_t per_batch = tiles_x * tiles_y;
    for (int64_t t = blockIdx.x; t < per_batch * batches; t += gridDim.x) {
        const int64_t bt = t / per_batch, tt = t - bt * per_batch;
        const int64_t y0 = (tt / tiles_x) * TS, x0 = (tt - (tt / tiles_x) * tiles_x) * TS;
        int64_t rest = bt, src = 0, dst = 0;
#pragma unroll 1
        for (int a = d.nd - 1; a >= 0; a--) {
            const int64_t q = rest / d.size[a];
            const int64_t i = rest - q * d.size[a];
            src += i * d.sstride[a];
            dst += i * d.tstride[a];
            rest = q;
        }
        Pack v[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; p++) {
            const int64_t y = y0 + r0 + p * RPP, x = x0 + xv * V;
            if (y < ny && x < nx) v[p].raw = __ldcs(reinterpret_cast<const int4*>(A + src + y * sy + x));
        }
#pragma unroll
        for (int p = 0; p < PASSES; p++) {
#pragma unroll
            for (int e = 0; e < V; e++) tile[r0 + p * RPP][xv * V + e] = v[p].v[e];
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < PASSES; p++) {
            const int xr = r0 + p * RPP;
            const int64_t x = x0 + xr, y = y0 + xv * V;
            if (x < nx && y < ny) {
                Pack o;
#pragma unroll
                for (int e = 0; e < V; e++) o.v[e] = tile[xv * V + e][xr];
                __stcs(reinterpret_cast<int4*>(B + dst + x * tx + y), o.raw);
            }
        }
        __syncthreads();
    }
}

template <typename W>
int transpose_typed(int nd, const int64_t* size, const int64_t* ss, const int64_t* ts,
                    const void* Av, int64_t offA, void* Bv, int64_t offB)
{
    // axes arrive in TARGET order (slowest first) with their source / target strides, already merged
    rb200::Context& c = rb200::ctx();
    const W* A = reinterpret_cast<const W*>(Av) + offA;
    W* B = reinterpret_cast<W*>(Bv) + offB;
    int64_t total = 1;
    for (int a = 0; a < nd; a++) total *= size[a];
    const int last = nd - 1;
    TrDims d;
    if (ss[last] == 1) {
        // target-fastest axis is also source-fastest: contiguous runs of size[last]
        d.nd = nd - 1;
        for (int a = 0; a < nd - 1; a++) { d.size[a] = size[a]; d.sstride[a] = ss[a]; d.tstride[a] = ts[a]; }
        const int64_t inner = size[last];
        constexpr int VMAX = 16 / (int)sizeof(W);
        bool vec = (inner % VMAX == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
        for (int a = 0; a < nd - 1 && vec; a++) vec = (ss[a] % VMAX == 0);
        const int64_t cap = (int64_t)c.sm_count * 32;
        const bool small = total < (1LL << 32);
        if (vec && VMAX > 1) {
            const int64_t total_v = total / VMAX;
            int64_t grid = rb_ceil_div(total_v, 256 * 4);
            if (grid > cap) grid = cap;
            if (small) transpose_rows_kernel<W, VMAX, unsigned><<<(unsigned)grid, 256, 0, c.stream>>>(d, inner / VMAX, total_v, A, B);
            else       transpose_rows_kernel<W, VMAX, int64_t><<<(unsigned)grid, 256, 0, c.stream>>>(d, inner / VMAX, total_v, A, B);
        } else {
            int64_t grid = rb_ceil_div(total, 256 * 4);
            if (grid > cap) grid = cap;
            if (small) transpose_rows_kernel<W, 1, unsigned><<<(unsigned)grid, 256, 0, c.stream>>>(d, inner, total, A, B);
            else       transpose_rows_kernel<W, 1, int64_t><<<(unsigned)grid, 256, 0, c.stream>>>(d, inner, total, A, B);
        }
        RB_LAUNCH_CHECK("transpose_rows_kernel");
        return RB200_OK;
    }
    // find the source-fastest axis (source stride 1) among the remaining ones
    int xs = -1;
    for (int a = 0; a < nd - 1; a++) if (ss[a] == 1) xs = a;
    if (xs < 0) RB_INVALID("transpose: no unit-stride source axis");
    d.nd = 0;
    int64_t batches = 1;
    for (int a = 0; a < nd - 1; a++) {
        if (a == xs) continue;
        d.size[d.nd] = size[a]; d.sstride[d.nd] = ss[a]; d.tstride[d.nd] = ts[a];
        d.nd++;
        batches *= size[a];
    }
    const int64_t nx = size[xs], ny = size[last];
    if (sizeof(W) >= 4) {
        constexpr int V = 16 / (int)sizeof(W);
        bool vec = nx % V == 0 && ny % V == 0 && ss[last] % V == 0 && ts[xs] % V == 0 &&
                   ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
        for (int a = 0; a < d.nd && vec; a++) vec = d.sstride[a] % V == 0 && d.tstride[a] % V == 0;
        if (vec) {
            int64_t grid = rb_ceil_div(nx, 64) * rb_ceil_div(ny, 64) * batches;
            const int64_t cap = (int64_t)c.sm_count * 32;
            if (grid > cap) grid = cap;
            transpose_tile_vec_kernel<W><<<(unsigned)grid, 256, 0, c.stream>>>(d, nx, ny, ss[last], ts[xs], batches, A, B);
            RB_LAUNCH_CHECK("transpose_tile_vec_kernel");
            return RB200_OK;
        }
    }
    const int64_t gx = rb_ceil_div(nx, 32), gy = rb_ceil_div(ny, 32);
    if (gy > 65535) RB_UNSUPPORTED("transpose: axis of %lld elements is too long for the tile kernel", (long long)ny);
    const int64_t gz = batches < 65535 ? batches : 65535;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
    transpose_tile_kernel<W><<<grid, 256, 0, c.stream>>>(d, nx, ny, ss[last], ts[xs], batches, A, B);
    RB_LAUNCH_CHECK("transpose_tile_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" {

int rb200_transpose(int dtype, int ndim, const int64_t* shape, const int64_t* perm,
                    const void* A, int64_t offA, void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!A || !B || !shape || !perm) RB_INVALID("NULL argument");
    if (ndim < 1) RB_INVALID("Matrix must not be a scalar");
    if (ndim > TR_MAXD) RB_UNSUPPORTED("transpose: at most %d dimensions", TR_MAXD);
    if (offA < 0 || offB < 0) RB_INVALID("bad offset");
    // source strides, and the axes in target order
    int64_t sstride[TR_MAXD], size_t_[TR_MAXD], ss_t[TR_MAXD];
    int64_t s = 1;
    bool seen[TR_MAXD] = {};
    for (int a = ndim - 1; a >= 0; a--) {
        if (shape[a] < 1) RB_INVALID("transpose: empty axis");
        sstride[a] = s;
        s *= shape[a];
    }
    for (int a = 0; a < ndim; a++) {
        const int64_t pa = perm[a];
        if (pa < 0 || pa >= ndim) RB_INVALID("dim value in the perm is out of range.");
        if (seen[pa]) RB_INVALID("duplicate axis in perm option");
        seen[pa] = true;
        size_t_[a] = shape[pa];
        ss_t[a] = sstride[pa];
    }
    // drop size-1 axes; merge neighbours (in target order) that are also neighbours in the source
    int64_t size[TR_MAXD], ss[TR_MAXD], ts[TR_MAXD];
    int nd = 0;
    for (int a = 0; a < ndim; a++) {
        if (size_t_[a] == 1) continue;
        if (nd > 0 && ss[nd - 1] == ss_t[a] * size_t_[a]) {       // previous axis strides over exactly this one
            size[nd - 1] *= size_t_[a];
            ss[nd - 1] = ss_t[a];
        } else {
            size[nd] = size_t_[a];
            ss[nd] = ss_t[a];
            nd++;
        }
    }
    if (nd == 0) { size[0] = 1; ss[0] = 1; nd = 1; }
    int64_t t = 1;
    for (int a = nd - 1; a >= 0; a--) { ts[a] = t; t *= size[a]; }
    switch (rb_dtype_size(dtype)) {
        case 1: return transpose_typed<uint8_t>(nd, size, ss, ts, A, offA, B, offB);
        case 2: return transpose_typed<uint16_t>(nd, size, ss, ts, A, offA, B, offB);
        case 4: return transpose_typed<uint32_t>(nd, size, ss, ts, A, offA, B, offB);
        case 8: return transpose_typed<uint64_t>(nd, size, ss, ts, A, offA, B, offB);
        default: RB_UNSUPPORTED("transpose: unsupported dtype %d", dtype);
    }
}

}  // extern "C"
