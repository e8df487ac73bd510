// im2col.cu -- im2col2d forward (gather) and reverse (col2im, gather-form) (HBM-bound,
// write-dominated).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:2091-2148 copyCell2d, :2158-2314):
//   forward : cols[b,oy,ox,ky,kx,c] = images[b, oy*sh + ky*dh - ph, ox*sw + kx*dw - pw, c]
//             or 0 outside the image;  cols_channels_first -> cols[b,oy,ox,c,ky,kx];
//             channels_first images are [b,c,h,w].
//   reverse : images[...] += cols[...] for every in-bounds tap, accumulated in the
//             reference's loop order (oy, ox ascending), on top of the existing image values.
//   out = (in - (k-1)*d - 1) div s + 1;  padding => out = in,
//   p = ((in-1)*s - in + (k-1)*d + 1) div 2                         (PhpMath.php:2217-2241)
// A pure copy: results are bit-exact for every dtype (forward) and exact in the reference's
// summation order (reverse).
// Algorithmic bytes (forward): images read once + cols written once.
//
// Forward kernels: (1) tiled -- one CTA per (batch, out row, block of out columns): the filter_h
// source rows are staged in shared memory with coalesced loads, the block's contiguous run of cols
// is produced with a tap table + shared-memory gather and coalesced stores (the common case);
// (2) flat -- one thread per 16-byte group of cols with a global gather, for shapes whose working
// set does not fit 48 KB of shared memory.  Reverse kernel: one thread per image element,
// summing its <= kh*kw taps -- no atomics, deterministic order.
#include "common.cuh"

namespace {

struct Im2colParams {
    int64_t batches, im_h, im_w, channels;
    int64_t filter_h, filter_w, stride_h, stride_w, dilation_h, dilation_w;
    int64_t out_h, out_w, padding_h, padding_w;
    int64_t im_w_step, im_h_step, channel_step, batch_step;   // image strides in elements
    int     cols_channels_first;
    int64_t cell;            // filter_h*filter_w*channels
    int64_t total_cols;      // batches*out_h*out_w*cell
    int64_t total_images;
};

constexpr int IC_THREADS = 256;
constexpr int IC_MLP = 8;             // independent global loads in flight per thread while staging
constexpr int IC_MAX_CELL = 4096;     // entries of the shared-memory tap table

struct Tap { int delta; short dy, dx; };   // 8 bytes

template <typename W> union IcPack { int4 raw; W v[16 / sizeof(W)]; };

// VEC: cols + offset is 16-byte aligned -> 128-bit stores of 16/sizeof(W) elements.
template <typename W, bool VEC, bool TABLE>
__global__ void __launch_bounds__(IC_THREADS)
im2col2d_forward(Im2colParams p, const W* __restrict__ images, W* __restrict__ cols)
{
    constexpr int V = VEC ? 16 / (int)sizeof(W) : 1;
    __shared__ Tap taps[TABLE ? IC_MAX_CELL : 1];
    const int cell = (int)p.cell;
    const int khw = (int)(p.filter_h * p.filter_w);
    auto make_tap = [&](int r) {
        int ky, kx, c;
        if (p.cols_channels_first) { c = r / khw; int q = r - c * khw; ky = q / (int)p.filter_w; kx = q - ky * (int)p.filter_w; }
        else { int q = r / (int)p.channels; c = r - q * (int)p.channels; ky = q / (int)p.filter_w; kx = q - ky * (int)p.filter_w; }
        Tap t;
        t.dy = (short)(ky * p.dilation_h);
        t.dx = (short)(kx * p.dilation_w);
        t.delta = (int)(t.dy * p.im_h_step + t.dx * p.im_w_step + c * p.channel_step);
        return t;
    };
    if (TABLE) {
        for (int r = threadIdx.x; r < cell; r += IC_THREADS) taps[r] = make_tap(r);
        __syncthreads();
    }
    const int64_t ngroups = (p.total_cols + V - 1) / V;
    const int64_t cells_per_batch = p.out_h * p.out_w;
    for (int64_t g = (int64_t)blockIdx.x * IC_THREADS + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * IC_THREADS) {
        const int64_t o = g * V;
        int64_t cell_id = o / cell;
        int r = (int)(o - cell_id * cell);
        int64_t b = cell_id / cells_per_batch;
        int64_t q = cell_id - b * cells_per_batch;
        int64_t oy = q / p.out_w;
        int64_t ox = q - oy * p.out_w;
        int64_t y0 = oy * p.stride_h - p.padding_h;
        int64_t x0 = ox * p.stride_w - p.padding_w;
        const W* src = images + b * p.batch_step + y0 * p.im_h_step + x0 * p.im_w_step;
        IcPack<W> out;
#pragma unroll
        for (int e = 0; e < V; e++) {
            W val = (W)0;
            if (o + e < p.total_cols) {
                const Tap t = TABLE ? taps[r] : make_tap(r);
                const int64_t iy = y0 + t.dy, ix = x0 + t.dx;
                if (iy >= 0 && iy < p.im_h && ix >= 0 && ix < p.im_w) val = __ldg(src + t.delta);
            }
            out.v[e] = val;
            if (++r == cell) {           // next cell: advance (ox, oy, b) with carries
                r = 0;
                ox++; x0 += p.stride_w;
                if (ox == p.out_w) {
                    ox = 0; x0 = -p.padding_w;
                    oy++; y0 += p.stride_h;
                    if (oy == p.out_h) { oy = 0; y0 = -p.padding_h; b++; }
                }
                src = images + b * p.batch_step + y0 * p.im_h_step + x0 * p.im_w_step;
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


// ---- forward, tiled: one CTA per (batch, block of ROWS out rows, block of TOX out columns) --------------
// The source rows that feed this block of cells are staged ONCE in shared memory with coalesced loads
// (zero-filled outside the image, so the inner loop has no bounds test; consecutive out rows share
// filter_h - stride_h of their source rows), then every out row of the block -- one contiguous run of
// cols -- is produced by shared-memory gathers and 128-bit stores.  All index arithmetic is 32-bit;
// divisions are multiply-highs by precomputed reciprocals.
struct Im2colTile {
    int tox;            // out columns per CTA
    int rows;           // out rows per CTA
    int slabs;          // staged source rows: (rows-1)*stride_h + (filter_h-1)*dilation_h + 1
    int width_x;        // input columns staged per source row: (tox-1)*stride_w + (filter_w-1)*dilation_w + 1
    int row_elems;      // elements per staged source row: width_x * channels
    unsigned cell_magic;   // floor(2^32 / cell) + 1
    unsigned wx_magic;     // floor(2^32 / width_x) + 1
    unsigned row_magic;    // floor(2^32 / row_elems) + 1
};

// t / d for t*d < 2^32 with magic = floor(2^32/d) + 1 (d == 1 has no 32-bit magic)
__device__ __forceinline__ int ic_fastdiv(int t, unsigned magic, int d)
{
    return d == 1 ? t : (int)__umulhi((unsigned)t, magic);
}

template <typename W>
__global__ void __launch_bounds__(IC_THREADS)
im2col2d_forward_tiled(Im2colParams p, Im2colTile tl, const W* __restrict__ images, W* __restrict__ cols)
{
    extern __shared__ __align__(16) unsigned char ic_smem_raw[];
    int* lut = reinterpret_cast<int*>(ic_smem_raw);                       // [cell] tap -> smem offset
    W* tile = reinterpret_cast<W*>(ic_smem_raw + ((p.cell * 4 + 15) & ~15LL));   // [slabs][row_elems]

    const int cell = (int)p.cell;
    const int C = (int)p.channels, KW = (int)p.filter_w, KH = (int)p.filter_h;
    const int ox0 = blockIdx.x * tl.tox;
    const int oy0 = blockIdx.y * tl.rows;
    const int b = blockIdx.z;
    const int tox = min(tl.tox, (int)p.out_w - ox0);
    const int nrows = min(tl.rows, (int)p.out_h - oy0);
    const bool cf = p.channel_step != 1;                                 // channels_first images
    // smem element (slab s, x, c): channels-last s*row_elems + x*C + c ; channels-first s*row_elems + c*width_x + x
    const int xs = cf ? 1 : C;
    const int dil_h = (int)p.dilation_h;
    for (int r = threadIdx.x; r < cell; r += IC_THREADS) {
        int ky, kx, c;
        if (p.cols_channels_first) { c = r / (KH * KW); int q = r - c * KH * KW; ky = q / KW; kx = q - ky * KW; }
        else { int q = r / C; c = r - q * C; ky = q / KW; kx = q - ky * KW; }
        lut[r] = ky * dil_h * tl.row_elems + kx * (int)p.dilation_w * xs + (cf ? c * tl.width_x : c);
    }
    // stage the source rows: one flat index space; each thread issues up to IC_MLP independent global
    // loads before the first shared-memory store, so the HBM/L2 latency is paid once per CTA.
    const int x_start = ox0 * (int)p.stride_w - (int)p.padding_w;
    const int y_start = oy0 * (int)p.stride_h - (int)p.padding_h;
    const W* img_b = images + (int64_t)b * p.batch_step;
    const int x_lo = max(0, -x_start);                                   // first staged column inside the image
    const int x_hi = max(x_lo, min(tl.width_x, (int)p.im_w - x_start));  // one past the last
    const int slabs = min(tl.slabs, (nrows - 1) * (int)p.stride_h + (KH - 1) * dil_h + 1);
    const int total = slabs * tl.row_elems;
    const int im_h_step = (int)p.im_h_step, chan_step = (int)p.channel_step;
    for (int e0 = 0; e0 < total; e0 += IC_MLP * IC_THREADS) {
        W v[IC_MLP];
#pragma unroll
        for (int u = 0; u < IC_MLP; u++) {
            const int e = e0 + u * IC_THREADS + (int)threadIdx.x;
            v[u] = (W)0;
            if (e < total) {
                const int sl = ic_fastdiv(e, tl.row_magic, tl.row_elems);
                const int idx = e - sl * tl.row_elems;
                const int iy = y_start + sl;
                int off;
                bool ok = iy >= 0 && iy < (int)p.im_h;
                if (!cf) {
                    ok = ok && idx >= x_lo * C && idx < x_hi * C;
                    off = iy * im_h_step + x_start * C + idx;
                } else {
                    const int c = ic_fastdiv(idx, tl.wx_magic, tl.width_x);
                    const int x = idx - c * tl.width_x;
                    ok = ok && x >= x_lo && x < x_hi;
                    off = c * chan_step + iy * im_h_step + x_start + x;
                }
                if (ok) v[u] = __ldg(img_b + off);
            }
        }
#pragma unroll
        for (int u = 0; u < IC_MLP; u++) {
            const int e = e0 + u * IC_THREADS + (int)threadIdx.x;
            if (e < total) tile[e] = v[u];
        }
    }
    __syncthreads();

    const int n_out = tox * cell;
    const int sx = (int)p.stride_w * xs;
    constexpr int V = 16 / (int)sizeof(W);
    const int gangs = cell <= IC_THREADS ? IC_THREADS / cell : 0;
    const int gang = ic_fastdiv((int)threadIdx.x, tl.cell_magic, cell);
    const int tx = (int)threadIdx.x - gang * cell;
    for (int rr = 0; rr < nrows; rr++) {
        W* out = cols + (((int64_t)b * p.out_h + oy0 + rr) * p.out_w + ox0) * cell;
        const W* trow = tile + rr * (int)p.stride_h * tl.row_elems;
        // 128-bit stores: scalar head up to the next 16-byte boundary, vector body, scalar tail
        const int mis = (int)((reinterpret_cast<uintptr_t>(out) / sizeof(W)) & (V - 1));
        int head = (V - mis) & (V - 1);
        if (head > n_out) head = n_out;
        if ((int)threadIdx.x < head) {
            const int t = threadIdx.x;
            const int oxl = ic_fastdiv(t, tl.cell_magic, cell);
            out[t] = trow[lut[t - oxl * cell] + oxl * sx];
        }
        const int nvec = (n_out - head) / V;
        if (gangs > 0) {
            // "Gangs" of `cell` threads: thread tx of a gang always handles vector group tx of a run of
            // V cells (V*cell elements = cell vector groups), so the tap of each of its V elements --
            // and hence its shared-memory offset -- is loop invariant; successive runs add V*sx.
            if (gang < gangs) {
                int off[V];
#pragma unroll
                for (int e = 0; e < V; e++) {
                    const int u = head + V * tx + e;
                    const int ce = ic_fastdiv(u, tl.cell_magic, cell);
                    off[e] = lut[u - ce * cell] + ce * sx;
                }
                const int step = V * sx;
                for (int j = gang, q = tx + cell * gang; q < nvec; j += gangs, q += cell * gangs) {
                    const int base = j * step;
                    IcPack<W> pk;
#pragma unroll
                    for (int e = 0; e < V; e++) pk.v[e] = trow[off[e] + base];
                    __stcs(reinterpret_cast<int4*>(out + head + V * q), pk.raw);
                }
            }
        } else {
            for (int g = threadIdx.x; g < nvec; g += IC_THREADS) {
                const int t = head + g * V;
                int oxl = ic_fastdiv(t, tl.cell_magic, cell);
                int r = t - oxl * cell;
                int xoff = oxl * sx;
                IcPack<W> pk;
#pragma unroll
                for (int e = 0; e < V; e++) {
                    pk.v[e] = trow[lut[r] + xoff];
                    if (++r == cell) { r = 0; xoff += sx; }
                }
                __stcs(reinterpret_cast<int4*>(out + t), pk.raw);
            }
        }
        for (int t = head + nvec * V + threadIdx.x; t < n_out; t += IC_THREADS) {
            const int oxl = ic_fastdiv(t, tl.cell_magic, cell);
            out[t] = trow[lut[t - oxl * cell] + oxl * sx];
        }
    }
}

// reverse: one thread per image element; taps visited in the reference's accumulation order.
template <typename T>
__global__ void __launch_bounds__(IC_THREADS)
col2im2d_reverse(Im2colParams p, T* __restrict__ images, const T* __restrict__ cols)
{
    for (int64_t idx = (int64_t)blockIdx.x * IC_THREADS + threadIdx.x; idx < p.total_images;
         idx += (int64_t)gridDim.x * IC_THREADS) {
        // decompose the image element index (memory order) into (b, iy, ix, c)
        int64_t b = idx / p.batch_step;
        int64_t rem = idx - b * p.batch_step;
        int64_t c, iy, ix;
        if (p.channel_step == 1) {          // channels last: [b, h, w, c]
            iy = rem / p.im_h_step; rem -= iy * p.im_h_step;
            ix = rem / p.im_w_step; c = rem - ix * p.im_w_step;
        } else {                            // channels first: [b, c, h, w]
            c = rem / p.channel_step; rem -= c * p.channel_step;
            iy = rem / p.im_h_step; ix = rem - iy * p.im_h_step;
        }
        T acc = images[idx];
        // increasing oy <=> decreasing ky; within one oy increasing ox <=> decreasing kx
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
                const int64_t cellbase = ((b * p.out_h + oy) * p.out_w + ox) * p.cell;
                const int64_t within = p.cols_channels_first
                    ? (c * p.filter_h + ky) * p.filter_w + kx
                    : (ky * p.filter_w + kx) * p.channels + c;
                acc += __ldg(cols + cellbase + within);
            }
        }
        images[idx] = acc;
    }
}

static unsigned magic_u32(int64_t d) { return (unsigned)((1ULL << 32) / (unsigned long long)d) + 1u; }

template <typename W>
int forward_launch(const Im2colParams& p, const void* images, void* cols)
{
    rb200::Context& c = rb200::ctx();
    const W* im = reinterpret_cast<const W*>(images);
    W* co = reinterpret_cast<W*>(cols);
    // tiled kernel: needs the per-CTA working set in 48 KB of shared memory and 32-bit in-tile indices
    {
        const int64_t smem_budget = 48 * 1024 - 16;
        const int64_t lut_bytes = (p.cell * 4 + 15) & ~15LL;
        const int64_t es = (int64_t)sizeof(W);
        const int64_t base_x = (p.filter_w - 1) * p.dilation_w + 1;
        const int64_t base_y = (p.filter_h - 1) * p.dilation_h + 1;
        // widest column block whose filter_h-row working set fits: slabs(rows=1) * width_x * C * es
        int64_t max_wx = (smem_budget - lut_bytes) / (base_y * p.channels * es);
        int64_t tox = max_wx >= base_x ? (max_wx - base_x) / p.stride_w + 1 : 0;
        if (tox > p.out_w) tox = p.out_w;
        const bool dims_ok = tox >= 1 && p.cell <= 65536 && p.batches <= 65535 && p.out_h <= 65535 &&
                             p.im_h < (1LL << 30) && p.im_w < (1LL << 30) && p.out_w * p.stride_w < (1LL << 30) &&
                             p.batch_step < (1LL << 31);
        if (dims_ok) {
            const int64_t width_x = (tox - 1) * p.stride_w + base_x;
            const int64_t row_elems = width_x * p.channels;
            // Out rows per CTA (they share source rows): among the row counts whose working set fits, take the one
            // whose grid fills whole waves of resident CTAs best -- C4 with 8 rows is 1792 CTAs = 1.51 waves (the
            // second wave half empty), with 6 rows 2368 CTAs = exactly 2 waves.  Ties go to the larger count.
            int64_t rows = 1;
            const int64_t col_blocks = rb_ceil_div(p.out_w, tox);
            double best_eff = -1.0;
            for (int64_t r2 = 1; r2 <= 8 && r2 <= p.out_h; r2++) {
                const int64_t slabs2 = (r2 - 1) * p.stride_h + base_y;
                const int64_t smem2 = lut_bytes + slabs2 * row_elems * es;
                if (smem2 > smem_budget) break;
                const int64_t ctas = p.batches * rb_ceil_div(p.out_h, r2) * col_blocks;
                if (r2 > 1 && ctas < (int64_t)c.sm_count * 8) break;          // keep >= 8 CTAs per SM in the grid
                int64_t per_sm = (220 * 1024) / (smem2 + 1024);
                if (per_sm > 8) per_sm = 8;                                     // 2048 threads / 256
                if (per_sm < 1) per_sm = 1;
                const double waves = (double)ctas / (double)(per_sm * c.sm_count);
                double eff = waves / (double)rb_ceil_div(ctas, per_sm * c.sm_count);
                eff *= 1.0 - 0.02 * (double)(8 - r2) / 8.0;                     // fewer rows re-read more halo rows
                if (eff >= best_eff) { best_eff = eff; rows = r2; }
            }
            // few CTAs (small images): split the rows into narrower column blocks instead
            if (rows == 1 && tox > 64 && p.batches * p.out_h * col_blocks < (int64_t)c.sm_count * 4) {
                const int64_t want = rb_ceil_div((int64_t)c.sm_count * 4, p.batches * p.out_h);
                int64_t t2 = rb_ceil_div(p.out_w, want);
                if (t2 < 32) t2 = 32;
                if (t2 < tox) tox = t2;
            }
            Im2colTile tl;
            tl.tox = (int)tox;
            tl.rows = (int)rows;
            tl.slabs = (int)((rows - 1) * p.stride_h + base_y);
            tl.width_x = (int)((tox - 1) * p.stride_w + base_x);
            tl.row_elems = tl.width_x * (int)p.channels;
            tl.cell_magic = magic_u32(p.cell);
            tl.wx_magic = magic_u32(tl.width_x);
            tl.row_magic = magic_u32(tl.row_elems);
            const int64_t smem = lut_bytes + (int64_t)tl.slabs * tl.row_elems * es;
            const bool magic_ok = tox * p.cell * p.cell < (1LL << 32) &&                       // t / cell
                                  (int64_t)tl.slabs * tl.row_elems * tl.row_elems < (1LL << 32) &&   // e / row_elems
                                  (int64_t)tl.row_elems * tl.width_x < (1LL << 32) &&          // idx / width_x
                                  256LL * p.cell < (1LL << 32);
            if (smem <= smem_budget + 16 && magic_ok) {
                dim3 grid((unsigned)rb_ceil_div(p.out_w, tox), (unsigned)rb_ceil_div(p.out_h, rows), (unsigned)p.batches);
                im2col2d_forward_tiled<W><<<grid, IC_THREADS, (size_t)smem, c.stream>>>(p, tl, im, co);
                RB_LAUNCH_CHECK("im2col2d_forward_tiled");
                return RB200_OK;
            }
        }
    }
    const bool vec = (reinterpret_cast<uintptr_t>(co) & 15) == 0;
    const bool table = p.cell <= IC_MAX_CELL;
    const int V = vec ? 16 / (int)sizeof(W) : 1;
    int64_t grid = rb_ceil_div(rb_ceil_div(p.total_cols, V), IC_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (vec && table)       im2col2d_forward<W, true, true><<<(unsigned)grid, IC_THREADS, 0, c.stream>>>(p, im, co);
    else if (vec && !table) im2col2d_forward<W, true, false><<<(unsigned)grid, IC_THREADS, 0, c.stream>>>(p, im, co);
    else if (table)         im2col2d_forward<W, false, true><<<(unsigned)grid, IC_THREADS, 0, c.stream>>>(p, im, co);
    else                    im2col2d_forward<W, false, false><<<(unsigned)grid, IC_THREADS, 0, c.stream>>>(p, im, co);
    RB_LAUNCH_CHECK("im2col2d_forward");
    return RB200_OK;
}

template <typename T>
int reverse_launch(const Im2colParams& p, void* images, const void* cols)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(p.total_images, IC_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    col2im2d_reverse<T><<<(unsigned)grid, IC_THREADS, 0, c.stream>>>(p, reinterpret_cast<T*>(images),
                                                                    reinterpret_cast<const T*>(cols));
    RB_LAUNCH_CHECK("col2im2d_reverse");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_im2col2d(int dtype, int reverse,
                              void* images, int64_t images_offset, int64_t images_size,
                              int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                              int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                              int padding, int channels_first, int64_t dilation_h, int64_t dilation_w,
                              int cols_channels_first,
                              void* cols, int64_t cols_offset, int64_t cols_size)
{
    RB_REQUIRE_INIT();
    if (!images || !cols) RB_INVALID("NULL buffer");
    if (batches < 1 || im_h < 1 || im_w < 1 || channels < 1 || filter_h < 1 || filter_w < 1 ||
        stride_h < 1 || stride_w < 1 || dilation_h < 1 || dilation_w < 1) {
        RB_INVALID("im2col2d: shape parameters must be greater than 0");
    }
    if (images_offset < 0 || cols_offset < 0) RB_INVALID("negative offset");
    Im2colParams p;
    p.batches = batches; p.im_h = im_h; p.im_w = im_w; p.channels = channels;
    p.filter_h = filter_h; p.filter_w = filter_w; p.stride_h = stride_h; p.stride_w = stride_w;
    p.dilation_h = dilation_h; p.dilation_w = dilation_w;
    if (images_size != batches * im_h * im_w * channels) RB_INVALID("images buffer size is invalid");
    int64_t out_h = (im_h - (filter_h - 1) * dilation_h - 1) / stride_h + 1;
    int64_t out_w = (im_w - (filter_w - 1) * dilation_w - 1) / stride_w + 1;
    if (out_h <= 0 || out_w <= 0) RB_INVALID("Invalid shape or parameters.");
    if (padding) {
        p.padding_h = ((im_h - 1) * stride_h - im_h + (filter_h - 1) * dilation_h + 1) / 2;
        p.padding_w = ((im_w - 1) * stride_w - im_w + (filter_w - 1) * dilation_w + 1) / 2;
        out_h = im_h;
        out_w = im_w;
    } else {
        p.padding_h = 0;
        p.padding_w = 0;
    }
    p.out_h = out_h; p.out_w = out_w;
    p.cell = filter_h * filter_w * channels;
    p.total_cols = batches * out_h * out_w * p.cell;
    p.total_images = images_size;
    if (cols_size != p.total_cols) RB_INVALID("output buffer size is invalid");
    if (channels_first) {
        p.im_w_step = 1; p.im_h_step = im_w; p.channel_step = im_w * im_h; p.batch_step = im_w * im_h * channels;
    } else {
        p.channel_step = 1; p.im_w_step = channels; p.im_h_step = channels * im_w; p.batch_step = channels * im_w * im_h;
    }
    p.cols_channels_first = cols_channels_first ? 1 : 0;
    if (p.batch_step >= (1LL << 31)) RB_INVALID("im2col2d: one image exceeds 2^31 elements");
    if (filter_h * dilation_h >= 32768 || filter_w * dilation_w >= 32768) RB_INVALID("im2col2d: filter extent too large");

    const int es = rb_dtype_size(dtype);
    if (es == 0) RB_UNSUPPORTED("im2col2d: unsupported dtype %d", dtype);
    char* im = reinterpret_cast<char*>(images) + images_offset * es;
    char* co = reinterpret_cast<char*>(cols) + cols_offset * es;
    if (!reverse) {
        switch (es) {
            case 1: return forward_launch<uint8_t>(p, im, co);
            case 2: return forward_launch<uint16_t>(p, im, co);
            case 4: return forward_launch<uint32_t>(p, im, co);
            default: return forward_launch<uint64_t>(p, im, co);
        }
    }
    switch (dtype) {
        case RB200_FLOAT32: return reverse_launch<float>(p, im, co);
        case RB200_FLOAT64: return reverse_launch<double>(p, im, co);
        case RB200_INT32:   return reverse_launch<int32_t>(p, im, co);
        case RB200_INT64:   return reverse_launch<int64_t>(p, im, co);
        default: RB_UNSUPPORTED("col2im: unsupported dtype %d", dtype);
    }
}
