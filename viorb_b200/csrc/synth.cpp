/*
 * synth.cpp -- deterministic, integer-only synthetic frame generator (host, test/bench input only).
 *
 * Frames with realistic FAST corner density: mid-grey canvas, h*w/900 random filled shapes (axis
 * aligned rectangles and rotated ellipses, grey U[0,256)), 5-tap binomial blur, +-3 integer noise
 * (SURVEY.md section 8(d) "Synthetic inputs").  Pure integer arithmetic so that a (h, w, seed)
 * triple names the same bytes on every machine (golden fixtures refer to frames by seed).
 */
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) { next(); next(); }
    uint64_t next() {   /* splitmix64 */
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    int uniform(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo)); }   /* [lo,hi) */
};

void draw_shapes(std::vector<uint8_t>& img, int h, int w, Rng& rng) {
    const int n = (int)((int64_t)h * w / 900);
    for (int s = 0; s < n; s++) {
        const int cx = rng.uniform(0, w), cy = rng.uniform(0, h);
        const int a = rng.uniform(3, 30), b = rng.uniform(3, 30);
        const int grey = rng.uniform(0, 256);
        const bool ellipse = rng.next() & 1;
        int ux = rng.uniform(-64, 65), uy = rng.uniform(-64, 65);
        if (ux == 0 && uy == 0) ux = 1;
        const int R = a > b ? a : b;
        const int64_t u2 = (int64_t)ux * ux + (int64_t)uy * uy;
        for (int y = cy - R; y <= cy + R; y++) {
            if (y < 0 || y >= h) continue;
            for (int x = cx - R; x <= cx + R; x++) {
                if (x < 0 || x >= w) continue;
                const int dx = x - cx, dy = y - cy;
                bool in;
                if (ellipse) {
                    const int64_t p = (int64_t)dx * ux + (int64_t)dy * uy;
                    const int64_t q = -(int64_t)dx * uy + (int64_t)dy * ux;
                    in = p * p * b * b + q * q * a * a <= (int64_t)a * a * b * b * u2;
                } else {
                    in = (dx >= -a && dx <= a && dy >= -b && dy <= b);
                }
                if (in) img[(size_t)y * w + x] = (uint8_t)grey;
            }
        }
    }
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void binomial5(std::vector<uint8_t>& img, int h, int w) {
    static const int k[5] = {1, 4, 6, 4, 1};
    std::vector<uint16_t> t((size_t)h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int acc = 0;
            for (int i = 0; i < 5; i++) acc += k[i] * img[(size_t)y * w + clampi(x + i - 2, 0, w - 1)];
            t[(size_t)y * w + x] = (uint16_t)acc;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int acc = 0;
            for (int i = 0; i < 5; i++) acc += k[i] * t[(size_t)clampi(y + i - 2, 0, h - 1) * w + x];
            img[(size_t)y * w + x] = (uint8_t)((acc + 128) >> 8);
        }
}

void add_noise(const std::vector<uint8_t>& base, uint8_t* out, int h, int w, Rng& rng, int amp) {
    for (size_t i = 0; i < (size_t)h * w; i++) {
        int v = base[i] + rng.uniform(-amp, amp + 1);
        out[i] = (uint8_t)clampi(v, 0, 255);
    }
}
}  // namespace

extern "C" {

/* one grayscale frame, row-major, stride = w */
void viorb_synth_frame(int h, int w, uint64_t seed, uint8_t* out) {
    Rng rng(seed);
    std::vector<uint8_t> img((size_t)h * w, 128);
    draw_shapes(img, h, w, rng);
    binomial5(img, h, w);
    add_noise(img, out, h, w, rng, 3);
}

/* n frames (seeds seed0 .. seed0+n-1), packed [n][h][w], generated on nthreads host threads */
void viorb_synth_frames(int n, int h, int w, uint64_t seed0, uint8_t* out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=] {
            for (int i = t; i < n; i += nthreads) viorb_synth_frame(h, w, seed0 + i, out + (size_t)i * h * w);
        });
    for (auto& t : th) t.join();
}

/* rectified stereo pair: right(x) = left(x + d(band)), d piece-wise constant over nbands horizontal
 * bands, d in [dmin, dmax); independent +-2 noise on both images.  disparities[nbands] returned. */
void viorb_synth_stereo(int h, int w, uint64_t seed, int nbands, int dmin, int dmax,
                        uint8_t* left, uint8_t* right, int* disparities) {
    Rng rng(seed);
    std::vector<uint8_t> base((size_t)h * w, 128), rbase((size_t)h * w);
    draw_shapes(base, h, w, rng);
    binomial5(base, h, w);
    for (int b = 0; b < nbands; b++) disparities[b] = rng.uniform(dmin, dmax);
    for (int y = 0; y < h; y++) {
        const int d = disparities[(int)((int64_t)y * nbands / h)];
        for (int x = 0; x < w; x++) rbase[(size_t)y * w + x] = base[(size_t)y * w + clampi(x + d, 0, w - 1)];
    }
    add_noise(base, left, h, w, rng, 2);
    add_noise(rbase, right, h, w, rng, 2);
}

/* uniform random bytes (descriptor maps) */
void viorb_synth_bytes(uint64_t seed, uint8_t* out, size_t n) {
    Rng rng(seed);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t v = rng.next();
        memcpy(out + i, &v, 8);
    }
    if (i < n) {
        uint64_t v = rng.next();
        memcpy(out + i, &v, n - i);
    }
}

}  // extern "C"
