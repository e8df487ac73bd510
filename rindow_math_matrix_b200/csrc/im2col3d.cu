// im2col3d.cu -- im2col3d forward (gather) and reverse (col2im, gather form): the Conv3D sibling of im2col2d
// (HBM-bound, write-dominated data movement; SURVEY.md section 8f: the callers either side of config C4).
// im2col1d needs no kernel of its own: it is im2col2d with im_h = filter_h = 1 (same layouts, PhpMath.php:1967-2088).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:2319-2390 copyCell3d, :2396-2600):
//   forward : cols[b,oz,oy,ox,kz,ky,kx,c] = images[b, oz*sd + kz*dd - pd, oy*sh + ky*dh - ph, ox*sw + kx*dw - pw, c]
//             or 0 outside the volume; cols_channels_first -> cols[b,oz,oy,ox,c,kz,ky,kx]; channels_first
//             images are [b,c,d,h,w].
//   reverse : images[...] += cols[...] for every in-bounds tap, in the reference's loop order (oz, oy, ox ascending).
//   out = (in - (k-1)*dil - 1) div s + 1;  padding => out = in and p = ((in-1)*s - in + (k-1)*dil + 1) div 2 --
//   EXCEPT that the reference computes padding_d with im_h in place of im_d (:2459); reproduced here.
// A pure copy: bit-exact for every dtype (forward), exact in the reference's summation order (reverse).
//
// Forward kernels: (1) tiled -- shared-memory staging of the source rows a block of out rows needs (the common case);
// (2) flat -- one thread per 16-byte group of cols with a global gather, for shapes whose working set does not fit
// 48 KB of shared memory.  Reverse: one thread per volume element summing its taps, no atomics.
#include <stdlib.h>

#include "common.cuh"

namespace {

struct Ic3Params {
    int64_t batches, im_d, im_h, im_w, channels;
    int64_t filter_d, filter_h, filter_w, stride_d, stride_h, stride_w, dilation_d, dilation_h, dilation_w;
    int64_t out_d, out_h, out_w, padding_d, padding_h, padding_w;
    int64_t im_w_step, im_h_step, im_d_step, channel_step, batch_step;   // image strides in elements
    int     cols_channels_first;
    int64_t cell;            // filter_d*filter_h*filter_w*channels
    int64_t total_cols, total_images;
};

constexpr int IC3_THREADS = 256;
constexpr int IC3_MAX_CELL = 2048;       // 32 KB of shared memory for the tap table

struct Tap3 { int64_t delta; short dz, dy, dx, pad; };   // 16 bytes

template <typename W> union Ic3Pack { int4 raw; W v[16 / sizeof(W)]; };

__device__ __forceinline__ Tap3 ic3_make_tap(const Ic3Params& p, int r)
{
    const int khw = (int)(p.filter_h * p.filter_w), kdhw = (int)(p.filter_d * p.filter_h * p.filter_w);
    int kz, ky, kx, c, q;
    if (p.cols_channels_first) { c = r / kdhw; q = r - c * kdhw; }
    else { q = r / (int)p.channels; c = r - q * (int)p.channels; }
    kz = q / khw; q -= kz * khw;
    ky = q / (int)p.filter_w; kx = q - ky * (int)p.filter_w;
    Tap3 t;
    t.dz = (short)(kz * p.dilation_d);
    t.dy = (short)(ky * p.dilation_h);
    t.dx = (short)(kx * p.dilation_w);
    t.pad = 0;
    t.delta = t.dz * p.im_d_step + t.dy * p.im_h_step + t.dx * p.im_w_step + c * p.channel_step;
    return t;
}

template <typename W, bool VEC, bool TABLE>
__global__ void __launch_bounds__(IC3_THREADS)
im2col3d_forward(Ic3Params p, const W* __restrict__ images, W* __restrict__ cols)
{
    constexpr int V = VEC ? 16 / (int)sizeof(W) : 1;
    __shared__ Tap3 taps[TABLE ? IC3_MAX_CELL : 1];
    const int cell = (int)p.cell;
    if (TABLE) {
        for (int r = threadIdx.x; r < cell; r += IC3_THREADS) taps[r] = ic3_make_tap(p, r);
        __syncthreads();
    }
    const int64_t ngroups = (p.total_cols + V - 1) / V;
    const int64_t plane = p.out_h * p.out_w, volume = p.out_d * plane;
    for (int64_t g = (int64_t)blockIdx.x * IC3_THREADS + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * IC3_THREADS) {
        const int64_t o = g * V;
        const int64_t cell_id = o / cell;
        int r = (int)(o - cell_id * cell);
        int64_t b = cell_id / volume;
        int64_t q = cell_id - b * volume;
        int64_t oz = q / plane; q -= oz * plane;
        int64_t oy = q / p.out_w;
        int64_t ox = q - oy * p.out_w;
        int64_t z0 = oz * p.stride_d - p.padding_d, y0 = oy * p.stride_h - p.padding_h, x0 = ox * p.stride_w - p.padding_w;
        const W* src = images + b * p.batch_step + z0 * p.im_d_step + y0 * p.im_h_step + x0 * p.im_w_step;
        Ic3Pack<W> out;
#pragma unroll
        for (int e = 0; e < V; e++) {
            W val = (W)0;
            if (o + e < p.total_cols) {
                const Tap3 t = TABLE ? taps[r] : ic3_make_tap(p, r);
                const int64_t iz = z0 + t.dz, iy = y0 + t.dy, ix = x0 + t.dx;
                if (iz >= 0 && iz < p.im_d && iy >= 0 && iy < p.im_h && ix >= 0 && ix < p.im_w) val = __ldg(src + t.delta);
            }
            out.v[e] = val;
            if (++r == cell) {           // next cell: advance (ox, oy, oz, b) with carries
                r = 0;
                ox++; x0 += p.stride_w;
                if (ox == p.out_w) {
                    ox = 0; x0 = -p.padding_w;
                    oy++; y0 += p.stride_h;
                    if (oy == p.out_h) {
                        oy = 0; y0 = -p.padding_h;
                        oz++; z0 += p.stride_d;
                        if (oz == p.out_d) { oz = 0; z0 = -p.padding_d; b++; }
                    }
                }
                src = images + b * p.batch_step + z0 * p.im_d_step + y0 * p.im_h_step + x0 * p.im_w_step;
            }
        }
        if (VEC) {
            if (o + V <= p.total_cols) {
                __stcs(reinterpret_cast<int4*>(cols + o), out.raw);
            } else {
                for (int e = 0; o + e < p.total_cols; e++) cols[o + e] = out.v[e];
            }
        } else {
            cols[o] = out.v[0];
        }
    }
}

// ---- forward, tiled: the im2col2d tile design with one more tap axis -------------------------------------------
// One CTA per (batch, out plane oz, block of `rows` out rows, block of `tox` out columns).  For a fixed oz the cols of
// a row are a 2-D im2col over a "virtual" image whose tap rows are the (kz, ky) pairs: the filter_d groups of source
// rows (one group per z tap, zero filled outside the volume) are staged once in shared memory with batched coalesced
// loads, then every out row of the block -- one contiguous run of tox*cell elements of cols -- is produced by gangs of
// threads with loop-invariant tap offsets (one shared-memory gather + one 128-bit streaming store per step).
struct Ic3Tile {
    int tox, rows;
    int slabs_y;        // staged source rows per z tap: (rows-1)*stride_h + (filter_h-1)*dilation_h + 1
    int width_x;        // input columns staged per source row
    int row_elems;      // elements per staged source row: width_x * channels
    unsigned cell_magic, wx_magic, row_magic, slab_magic;      // floor(2^32 / d) + 1
};

constexpr int IC3_MLP = 8;

__device__ __forceinline__ int ic3_fastdiv(int t, unsigned magic, int d)
{
    return d == 1 ? t : (int)__umulhi((unsigned)t, magic);
}

template <typename W>
__global__ void __launch_bounds__(IC3_THREADS)
im2col3d_forward_tiled(Ic3Params p, Ic3Tile tl, const W* __restrict__ images, W* __restrict__ cols)
{
    extern __shared__ __align__(16) unsigned char ic3_smem_raw[];
    int* lut = reinterpret_cast<int*>(ic3_smem_raw);                                   // [cell] tap -> smem offset
    W* tile = reinterpret_cast<W*>(ic3_smem_raw + ((p.cell * 4 + 15) & ~15LL));        // [filter_d][slabs_y][row_elems]

    const int cell = (int)p.cell;
    const int C = (int)p.channels, KW = (int)p.filter_w, KH = (int)p.filter_h, KD = (int)p.filter_d;
    const int ox0 = blockIdx.x * tl.tox;
    const int oy0 = blockIdx.y * tl.rows;
    const int b = blockIdx.z / (int)p.out_d;
    const int oz = blockIdx.z - b * (int)p.out_d;
    const int tox = min(tl.tox, (int)p.out_w - ox0);
    const int nrows = min(tl.rows, (int)p.out_h - oy0);
    const bool cf = p.channel_step != 1;                                               // channels_first volumes
    const int xs = cf ? 1 : C;
    const int dil_h = (int)p.dilation_h;
    for (int r = threadIdx.x; r < cell; r += IC3_THREADS) {
        int kz, ky, kx, c, q;
        if (p.cols_channels_first) { c = r / (KD * KH * KW); q = r - c * KD * KH * KW; }
        else { q = r / C; c = r - q * C; }
        kz = q / (KH * KW); q -= kz * KH * KW;
        ky = q / KW; kx = q - ky * KW;
        lut[r] = (kz * tl.slabs_y + ky * dil_h) * tl.row_elems + kx * (int)p.dilation_w * xs + (cf ? c * tl.width_x : c);
    }
    const int x_start = ox0 * (int)p.stride_w - (int)p.padding_w;
    const int y_start = oy0 * (int)p.stride_h - (int)p.padding_h;
    const int z_start = oz * (int)p.stride_d - (int)p.padding_d;
    const W* img_b = images + (int64_t)b * p.batch_step;
    const int x_lo = max(0, -x_start);
    const int x_hi = max(x_lo, min(tl.width_x, (int)p.im_w - x_start));
    const int total = KD * tl.slabs_y * tl.row_elems;
    for (int e0 = 0; e0 < total; e0 += IC3_MLP * IC3_THREADS) {
        W v[IC3_MLP];
#pragma unroll
        for (int u = 0; u < IC3_MLP; u++) {
            const int e = e0 + u * IC3_THREADS + (int)threadIdx.x;
            v[u] = (W)0;
            if (e < total) {
                const int sl = ic3_fastdiv(e, tl.row_magic, tl.row_elems);
                const int idx = e - sl * tl.row_elems;
                const int kz = ic3_fastdiv(sl, tl.slab_magic, tl.slabs_y);
                const int iy = y_start + (sl - kz * tl.slabs_y);
                const int iz = z_start + kz * (int)p.dilation_d;
                bool ok = iy >= 0 && iy < (int)p.im_h && iz >= 0 && iz < (int)p.im_d;
                int64_t off = (int64_t)iz * p.im_d_step + (int64_t)iy * p.im_h_step;
                if (!cf) {
                    ok = ok && idx >= x_lo * C && idx < x_hi * C;
                    off += x_start * C + idx;
                } else {
                    const int c = ic3_fastdiv(idx, tl.wx_magic, tl.width_x);
                    const int x = idx - c * tl.width_x;
                    ok = ok && x >= x_lo && x < x_hi;
                    off += (int64_t)c * p.channel_step + x_start + x;
                }
                if (ok) v[u] = __ldg(img_b + off);
            }
        }
#pragma unroll
        for (int u = 0; u < IC3_MLP; u++) {
            const int e = e0 + u * IC3_THREADS + (int)threadIdx.x;
            if (e < total) tile[e] = v[u];
        }
    }
    __syncthreads();

    const int n_out = tox * cell;
    const int sx = (int)p.stride_w * xs;
    constexpr int V = 16 / (int)sizeof(W);
    const int gangs = cell <= IC3_THREADS ? IC3_THREADS / cell : 0;
    const int gang = ic3_fastdiv((int)threadIdx.x, tl.cell_magic, cell);
    const int tx = (int)threadIdx.x - gang * cell;
    for (int rr = 0; rr < nrows; rr++) {
        W* out = cols + ((((int64_t)b * p.out_d + oz) * p.out_h + oy0 + rr) * p.out_w + ox0) * cell;
        const W* trow = tile + rr * (int)p.stride_h * tl.row_elems;
        const int mis = (int)((reinterpret_cast<uintptr_t>(out) / sizeof(W)) & (V - 1));
        int head = (V - mis) & (V - 1);
        if (head > n_out) head = n_out;
        if ((int)threadIdx.x < head) {
            const int t = threadIdx.x;
            const int oxl = ic3_fastdiv(t, tl.cell_magic, cell);
            out[t] = trow[lut[t - oxl * cell] + oxl * sx];
        }
        const int nvec = (n_out - head) / V;
        if (gangs > 0) {
            if (gang < gangs) {
                int off[V];
#pragma unroll
                for (int e = 0; e < V; e++) {
                    const int u = head + V * tx + e;
                    const int ce = ic3_fastdiv(u, tl.cell_magic, cell);
                    off[e] = lut[u - ce * cell] + ce * sx;
                }
                const int step = V * sx;
                for (int j = gang, q = tx + cell * gang; q < nvec; j += gangs, q += cell * gangs) {
                    const int base = j * step;
                    Ic3Pack<W> pk;
#pragma unroll
                    for (int e = 0; e < V; e++) pk.v[e] = trow[off[e] + base];
                    __stcs(reinterpret_cast<int4*>(out + head + V * q), pk.raw);
                }
            }
        } else {
            for (int g = threadIdx.x; g < nvec; g += IC3_THREADS) {
                const int t = head + g * V;
                int oxl = ic3_fastdiv(t, tl.cell_magic, cell);
                int r = t - oxl * cell;
                int xoff = oxl * sx;
                Ic3Pack<W> pk;
#pragma unroll
                for (int e = 0; e < V; e++) {
                    pk.v[e] = trow[lut[r] + xoff];
                    if (++r == cell) { r = 0; xoff += sx; }
                }
                __stcs(reinterpret_cast<int4*>(out + t), pk.raw);
            }
        }
        for (int t = head + nvec * V + threadIdx.x; t < n_out; t += IC3_THREADS) {
            const int oxl = ic3_fastdiv(t, tl.cell_magic, cell);
            out[t] = trow[lut[t - oxl * cell] + oxl * sx];
        }
    }
}

// reverse: one thread per image element; its taps are visited in the reference's accumulation order (cells ascending:
// increasing oz <=> decreasing kz, then oy / ky, then ox / kx) -- no atomics, deterministic.
template <typename T>
__global__ void __launch_bounds__(IC3_THREADS)
col2im3d_reverse(Ic3Params p, T* __restrict__ images, const T* __restrict__ cols)
{
    const int64_t kdhw = p.filter_d * p.filter_h * p.filter_w;
    for (int64_t idx = (int64_t)blockIdx.x * IC3_THREADS + threadIdx.x; idx < p.total_images;
         idx += (int64_t)gridDim.x * IC3_THREADS) {
        int64_t b = idx / p.batch_step;
        int64_t rem = idx - b * p.batch_step;
        int64_t c, iz, iy, ix;
        if (p.channel_step == 1) {          // channels last: [b, d, h, w, c]
            iz = rem / p.im_d_step; rem -= iz * p.im_d_step;
            iy = rem / p.im_h_step; rem -= iy * p.im_h_step;
            ix = rem / p.im_w_step; c = rem - ix * p.im_w_step;
        } else {                            // channels first: [b, c, d, h, w]
            c = rem / p.channel_step; rem -= c * p.channel_step;
            iz = rem / p.im_d_step; rem -= iz * p.im_d_step;
            iy = rem / p.im_h_step; ix = rem - iy * p.im_h_step;
        }
        T acc = images[idx];
        for (int64_t kz = p.filter_d - 1; kz >= 0; kz--) {
            const int64_t nz = iz + p.padding_d - kz * p.dilation_d;
            if (nz < 0 || nz % p.stride_d != 0) continue;
            const int64_t oz = nz / p.stride_d;
            if (oz >= p.out_d) continue;
            for (int64_t ky = p.filter_h - 1; ky >= 0; ky--) {
                const int64_t ny = iy + p.padding_h - ky * p.dilation_h;
                if (ny < 0 || ny % p.stride_h != 0) continue;
                const int64_t oy = ny / p.stride_h;
                if (oy >= p.out_h) continue;
                for (int64_t kx = p.filter_w - 1; kx >= 0; kx--) {
                    const int64_t nx = ix + p.padding_w - kx * p.dilation_w;
                    if (nx < 0 || nx % p.stride_w != 0) continue;
                    const int64_t ox = nx / p.stride_w;
                    if (ox >= p.out_w) continue;
                    const int64_t cellbase = (((b * p.out_d + oz) * p.out_h + oy) * p.out_w + ox) * p.cell;
                    const int64_t kpos = (kz * p.filter_h + ky) * p.filter_w + kx;
                    const int64_t within = p.cols_channels_first ? c * kdhw + kpos : kpos * p.channels + c;
                    acc += __ldg(cols + cellbase + within);
                }
            }
        }
        images[idx] = acc;
    }
}

static unsigned ic3_magic(int64_t d) { return (unsigned)((1ULL << 32) / (unsigned long long)d) + 1u; }

template <typename W>
int forward3_launch(const Ic3Params& p, const void* images, void* cols)
{
    rb200::Context& c = rb200::ctx();
    const W* im = reinterpret_cast<const W*>(images);
    W* co = reinterpret_cast<W*>(cols);
    // tiled kernel: the per-CTA working set must fit 48 KB of shared memory and 32-bit in-tile indices
    {
        const int64_t smem_budget = 48 * 1024 - 16;
        const int64_t lut_bytes = (p.cell * 4 + 15) & ~15LL;
        const int64_t es = (int64_t)sizeof(W);
        const int64_t base_x = (p.filter_w - 1) * p.dilation_w + 1;
        const int64_t base_y = (p.filter_h - 1) * p.dilation_h + 1;
        int64_t max_wx = (smem_budget - lut_bytes) / (p.filter_d * base_y * p.channels * es);
        int64_t tox = max_wx >= base_x ? (max_wx - base_x) / p.stride_w + 1 : 0;
        if (tox > p.out_w) tox = p.out_w;
        const int64_t planes = p.batches * p.out_d;
        const bool dims_ok = tox >= 1 && p.cell <= 65536 && planes <= 65535 && p.out_h <= 65535 &&
                             p.im_h < (1LL << 30) && p.im_w < (1LL << 30) && p.im_d < (1LL << 30) &&
                             p.out_w * p.stride_w < (1LL << 30);
        if (dims_ok) {
            const int64_t width_x = (tox - 1) * p.stride_w + base_x;
            const int64_t row_elems = width_x * p.channels;
            // out rows per CTA: wave fill x staging overhead x occupancy, ties to the larger count
            int64_t rows = 1;
            const int64_t col_blocks = rb_ceil_div(p.out_w, tox);
            double best_eff = -1.0;
            for (int64_t r2 = 1; r2 <= 8 && r2 <= p.out_h; r2++) {
                const int64_t slabs2 = (r2 - 1) * p.stride_h + base_y;
                const int64_t smem2 = lut_bytes + p.filter_d * slabs2 * row_elems * es;
                if (smem2 > smem_budget) break;
                const int64_t ctas = planes * rb_ceil_div(p.out_h, r2) * col_blocks;
                if (r2 > 1 && ctas < (int64_t)c.sm_count * 8) break;
                int64_t per_sm = (220 * 1024) / (smem2 + 1024);
                if (per_sm > 8) per_sm = 8;
                if (per_sm < 1) per_sm = 1;
                const double waves = (double)ctas / (double)(per_sm * c.sm_count);
                double eff = waves / (double)rb_ceil_div(ctas, per_sm * c.sm_count);          // wave quantisation
                // staging traffic per out row against the cols bytes it produces: every z tap re-stages its rows, so
                // few rows per CTA cost far more here than in im2col2d (measured: 1 row 0.57, 4 rows 0.85 of HBM)
                const double out_row = (double)(tox * p.cell), staged = (double)(p.filter_d * slabs2 * row_elems) / (double)r2;
                eff *= out_row / (out_row + staged);
                if (per_sm < 6) eff *= (double)per_sm / 6.0;                                   // too few warps to hide latency
                if (eff >= best_eff) { best_eff = eff; rows = r2; }
            }
            { const char* e = getenv("RB200_IC3_ROWS"); if (e && atoi(e) >= 1) rows = atoi(e); }
            Ic3Tile tl;
            tl.tox = (int)tox;
            tl.rows = (int)rows;
            tl.slabs_y = (int)((rows - 1) * p.stride_h + base_y);
            tl.width_x = (int)width_x;
            tl.row_elems = (int)row_elems;
            tl.cell_magic = ic3_magic(p.cell);
            tl.wx_magic = ic3_magic(tl.width_x);
            tl.row_magic = ic3_magic(tl.row_elems);
            tl.slab_magic = ic3_magic(tl.slabs_y);
            const int64_t total = p.filter_d * tl.slabs_y * (int64_t)tl.row_elems;
            const int64_t smem = lut_bytes + total * es;
            const bool magic_ok = tox * p.cell * p.cell < (1LL << 32) && total * tl.row_elems < (1LL << 32) &&
                                  (int64_t)tl.row_elems * tl.width_x < (1LL << 32) && 256LL * p.cell < (1LL << 32) &&
                                  p.filter_d * tl.slabs_y * (int64_t)tl.slabs_y < (1LL << 32);
            if (smem <= smem_budget + 16 && magic_ok) {
                dim3 grid((unsigned)col_blocks, (unsigned)rb_ceil_div(p.out_h, rows), (unsigned)planes);
                im2col3d_forward_tiled<W><<<grid, IC3_THREADS, (size_t)smem, c.stream>>>(p, tl, im, co);
                RB_LAUNCH_CHECK("im2col3d_forward_tiled");
                return RB200_OK;
            }
        }
    }
    const bool vec = (reinterpret_cast<uintptr_t>(co) & 15) == 0;
    const bool table = p.cell <= IC3_MAX_CELL;
    const int V = vec ? 16 / (int)sizeof(W) : 1;
    int64_t grid = rb_ceil_div(rb_ceil_div(p.total_cols, V), IC3_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (vec && table)       im2col3d_forward<W, true, true><<<(unsigned)grid, IC3_THREADS, 0, c.stream>>>(p, im, co);
    else if (vec && !table) im2col3d_forward<W, true, false><<<(unsigned)grid, IC3_THREADS, 0, c.stream>>>(p, im, co);
    else if (table)         im2col3d_forward<W, false, true><<<(unsigned)grid, IC3_THREADS, 0, c.stream>>>(p, im, co);
    else                    im2col3d_forward<W, false, false><<<(unsigned)grid, IC3_THREADS, 0, c.stream>>>(p, im, co);
    RB_LAUNCH_CHECK("im2col3d_forward");
    return RB200_OK;
}

template <typename T>
int reverse3_launch(const Ic3Params& p, void* images, const void* cols)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(p.total_images, IC3_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    col2im3d_reverse<T><<<(unsigned)grid, IC3_THREADS, 0, c.stream>>>(p, reinterpret_cast<T*>(images),
                                                                     reinterpret_cast<const T*>(cols));
    RB_LAUNCH_CHECK("col2im3d_reverse");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_im2col3d(int dtype, int reverse,
                              void* images, int64_t images_offset, int64_t images_size,
                              int64_t batches, int64_t im_d, int64_t im_h, int64_t im_w, int64_t channels,
                              int64_t filter_d, int64_t filter_h, int64_t filter_w,
                              int64_t stride_d, int64_t stride_h, int64_t stride_w,
                              int padding, int channels_first,
                              int64_t dilation_d, int64_t dilation_h, int64_t dilation_w,
                              int cols_channels_first,
                              void* cols, int64_t cols_offset, int64_t cols_size)
{
    RB_REQUIRE_INIT();
    if (!images || !cols) RB_INVALID("NULL buffer");
    if (batches < 1 || im_d < 1 || im_h < 1 || im_w < 1 || channels < 1 || filter_d < 1 || filter_h < 1 || filter_w < 1 ||
        stride_d < 1 || stride_h < 1 || stride_w < 1 || dilation_d < 1 || dilation_h < 1 || dilation_w < 1) {
        RB_INVALID("im2col3d: shape parameters must be greater than 0");
    }
    if (images_offset < 0 || cols_offset < 0) RB_INVALID("negative offset");
    Ic3Params p;
    p.batches = batches; p.im_d = im_d; p.im_h = im_h; p.im_w = im_w; p.channels = channels;
    p.filter_d = filter_d; p.filter_h = filter_h; p.filter_w = filter_w;
    p.stride_d = stride_d; p.stride_h = stride_h; p.stride_w = stride_w;
    p.dilation_d = dilation_d; p.dilation_h = dilation_h; p.dilation_w = dilation_w;
    if (images_size != batches * im_d * im_h * im_w * channels) RB_INVALID("images buffer size is invalid");
    int64_t out_d = (im_d - (filter_d - 1) * dilation_d - 1) / stride_d + 1;
    int64_t out_h = (im_h - (filter_h - 1) * dilation_h - 1) / stride_h + 1;
    int64_t out_w = (im_w - (filter_w - 1) * dilation_w - 1) / stride_w + 1;
    if (out_h <= 0 || out_w <= 0) RB_INVALID("Invalid shape or parameters.");        // the reference does not test out_d (:2453)
    if (padding) {
        p.padding_d = ((im_d - 1) * stride_d - im_h + (filter_d - 1) * dilation_d + 1) / 2;   // im_h: sic, :2459
        p.padding_h = ((im_h - 1) * stride_h - im_h + (filter_h - 1) * dilation_h + 1) / 2;
        p.padding_w = ((im_w - 1) * stride_w - im_w + (filter_w - 1) * dilation_w + 1) / 2;
        out_d = im_d; out_h = im_h; out_w = im_w;
    } else {
        p.padding_d = p.padding_h = p.padding_w = 0;
    }
    if (out_d <= 0) RB_INVALID("Invalid shape or parameters.");                       // would be an empty / negative loop there
    p.out_d = out_d; p.out_h = out_h; p.out_w = out_w;
    p.cell = filter_d * filter_h * filter_w * channels;
    p.total_cols = batches * out_d * out_h * out_w * p.cell;
    p.total_images = images_size;
    if (cols_size != p.total_cols) RB_INVALID("output buffer size is invalid");
    if (channels_first) {
        p.im_w_step = 1; p.im_h_step = im_w; p.im_d_step = im_w * im_h; p.channel_step = im_w * im_h * im_d;
        p.batch_step = im_w * im_h * im_d * channels;
    } else {
        p.channel_step = 1; p.im_w_step = channels; p.im_h_step = channels * im_w; p.im_d_step = channels * im_w * im_h;
        p.batch_step = channels * im_w * im_h * im_d;
    }
    p.cols_channels_first = cols_channels_first ? 1 : 0;
    if (filter_d * dilation_d >= 32768 || filter_h * dilation_h >= 32768 || filter_w * dilation_w >= 32768) {
        RB_INVALID("im2col3d: filter extent too large");
    }
    if (p.cell >= (1LL << 31)) RB_INVALID("im2col3d: cell too large");

    const int es = rb_dtype_size(dtype);
    if (es == 0) RB_UNSUPPORTED("im2col3d: unsupported dtype %d", dtype);
    char* im = reinterpret_cast<char*>(images) + images_offset * es;
    char* co = reinterpret_cast<char*>(cols) + cols_offset * es;
    if (!reverse) {
        switch (es) {
            case 1: return forward3_launch<uint8_t>(p, im, co);
            case 2: return forward3_launch<uint16_t>(p, im, co);
            case 4: return forward3_launch<uint32_t>(p, im, co);
            default: return forward3_launch<uint64_t>(p, im, co);
        }
    }
    switch (dtype) {
        case RB200_FLOAT32: return reverse3_launch<float>(p, im, co);
        case RB200_FLOAT64: return reverse3_launch<double>(p, im, co);
        case RB200_INT32:   return reverse3_launch<int32_t>(p, im, co);
        case RB200_INT64:   return reverse3_launch<int64_t>(p, im, co);
        default: RB_UNSUPPORTED("col2im3d: unsupported dtype %d", dtype);
    }
}
