// imagecopy.cu -- PhpMath::imagecopy (PhpMath.php:2877-2957): one image [h,w,c] (channels_first: [c,h,w]) copied with
// a shift, vertical / horizontal flips and an RGB -> BGR swap; source coordinates clamp to the border.  A pure gather:
// any dtype, bit-exact, one thread per destination element.
#include "common.cuh"

namespace {

constexpr int IC_THREADS = 256;

struct ImgCopy {
    int64_t height, width, channels, total;
    int64_t ldY, ldX, ldC;
    int64_t dirY, dirX, biasY, biasX;
    int channels_first, rgb_flip;
};

template <typename W>
__global__ void __launch_bounds__(IC_THREADS)
imagecopy_kernel(ImgCopy p, const W* __restrict__ A, W* __restrict__ B)
{
    for (int64_t t = (int64_t)blockIdx.x * IC_THREADS + threadIdx.x; t < p.total; t += (int64_t)gridDim.x * IC_THREADS) {
        int64_t y, x, c;
        if (p.channels_first) { x = t % p.width; y = (t / p.width) % p.height; c = t / (p.width * p.height); }
        else { c = t % p.channels; x = (t / p.channels) % p.width; y = t / (p.channels * p.width); }
        int64_t sy = y * p.dirY + p.biasY, sx = x * p.dirX + p.biasX;
        sy = sy < 0 ? 0 : (sy >= p.height ? p.height - 1 : sy);
        sx = sx < 0 ? 0 : (sx >= p.width ? p.width - 1 : sx);
        const int64_t sc = (p.rgb_flip && c < 3) ? 2 - c : c;
        B[y * p.ldY + x * p.ldX + c * p.ldC] = A[sy * p.ldY + sx * p.ldX + sc * p.ldC];
    }
}

template <typename W>
int imagecopy_launch(const ImgCopy& p, const void* A, void* B)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(p.total, IC_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    imagecopy_kernel<W><<<(unsigned)grid, IC_THREADS, 0, c.stream>>>(p, reinterpret_cast<const W*>(A), reinterpret_cast<W*>(B));
    RB_LAUNCH_CHECK("imagecopy_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_imagecopy(int dtype, int64_t height, int64_t width, int64_t channels, const void* A, int64_t offA,
                               void* B, int64_t offB, int channels_first, int64_t height_shift, int64_t width_shift,
                               int vertical_flip, int horizontal_flip, int rgb_flip)
{
    RB_REQUIRE_INIT();
    if (!A || !B) RB_INVALID("NULL buffer");
    if (height < 1 || width < 1 || channels < 1 || offA < 0 || offB < 0) RB_INVALID("imagecopy: bad shape / offset");
    if (rgb_flip && channels < 3) RB_INVALID("imagecopy: rgbFlip needs at least 3 channels");
    const int es = rb_dtype_size(dtype);
    ImgCopy p;
    p.height = height; p.width = width; p.channels = channels; p.total = height * width * channels;
    p.channels_first = channels_first != 0; p.rgb_flip = rgb_flip != 0;
    if (channels_first) { p.ldC = width * height; p.ldY = width; p.ldX = 1; }
    else { p.ldY = width * channels; p.ldX = channels; p.ldC = 1; }
    p.dirY = 1; p.dirX = 1; p.biasY = 0; p.biasX = 0;                    // PhpMath.php:2921-2933
    if (vertical_flip) { p.dirY = -1; p.biasY = height - 1; }
    if (horizontal_flip) { p.dirX = -1; p.biasX = width - 1; }
    p.biasY -= height_shift * p.dirY;
    p.biasX -= width_shift * p.dirX;
    const char* a = reinterpret_cast<const char*>(A) + offA * es;
    char* b = reinterpret_cast<char*>(B) + offB * es;
    switch (es) {
        case 1: return imagecopy_launch<uint8_t>(p, a, b);
        case 2: return imagecopy_launch<uint16_t>(p, a, b);
        case 4: return imagecopy_launch<uint32_t>(p, a, b);
        case 8: return imagecopy_launch<uint64_t>(p, a, b);
        default: RB_UNSUPPORTED("imagecopy: unsupported dtype %d", dtype);
    }
}
