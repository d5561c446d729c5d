// rng.cuh -- the two random sources of the MC kernel.
//
//  ReplayRng : bit-for-bit restatement of what the reference draws from --
//              glibc rand() (TYPE_3 additive feedback, RAND_MAX = 2^31-1;
//              move_funcs.pyx:14, bead_selection.pyx:10, linalg.pyx:9,
//              mc_sim.pyx:11,171) and numpy's legacy MT19937 randint
//              (move_funcs.pyx:819) -- so that a GPU run can be compared with
//              the reference draw for draw.
//  PhiloxRng : production.  Philox4x32-10, key = (seed, replica), a 64-bit
//              block counter per replica that persists across mc_sim calls.
//
// Both expose the same interface, so the proposal code is shared:
//   next31()  -> integer in [0, 2^31-1]   (what rand() returns)
//   uniform() -> (double)next31() / RAND_MAX in [0, 1]
//   randint(n)-> np.random.randint(0, n)
// Draws are consumed by lane 0 of the replica's warp (the reference's stream is
// sequential with data-dependent draw counts).
#pragma once
#include "launch.cuh"
#include "params.cuh"

struct ReplayRng {
    static constexpr bool kBatched = false; // one sequential stream: attempts are prepared one at a time
    uint32_t *st; // shared: r[0..30], f at [31], b at [32]
    uint32_t *mt; // global: mt[0..623], pos at [624]
    __device__ __forceinline__ void seek_attempt(unsigned long long, uint32_t = 0) {}
    __device__ __forceinline__ uint32_t position() const { return 0; }

    __device__ __forceinline__ uint32_t next31() {
        uint32_t f = st[31], b = st[32];
        uint32_t v = st[f] + st[b];
        st[f] = v;
        st[31] = (f + 1 >= 31) ? 0 : f + 1;
        st[32] = (b + 1 >= 31) ? 0 : b + 1;
        return v >> 1;
    }
    __device__ __forceinline__ double uniform() { return (double)next31() / CB_RAND_MAX; }

    __device__ uint32_t mt_next() {
        uint32_t pos = mt[624];
        if (pos >= 624) {
            for (int k = 0; k < 624; k++) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        uint32_t y = mt[pos];
        mt[624] = pos + 1;
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // numpy legacy randint(0, high): masked rejection on 32-bit draws; a range
    // of one value consumes no draw.
    __device__ int randint(int high) {
        uint32_t rng = (uint32_t)(high - 1);
        if (rng == 0) return 0;
        uint32_t mask = rng;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        do {
            v = mt_next() & mask;
        } while (v > rng);
        return (int)v;
    }
    // snapshot / restore of the rand() state (33 words) for the rare large
    // tangent-rotation path, which regenerates its per-bead draws on commit
    __device__ void save(uint32_t *dst) const {
        for (int i = 0; i < 33; i++) dst[i] = st[i];
    }
    __device__ void restore(const uint32_t *src) {
        for (int i = 0; i < 33; i++) st[i] = src[i];
    }
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Counter-based: draw i of attempt t of replica r is word (i & 3) of
// Philox4x32-10(counter = (t_lo, t_hi, i >> 2, r), key = seed).  Any lane can
// therefore produce any attempt's stream, which is what lets a batch of
// proposals be prepared by 32 lanes at once.
struct PhiloxRng {
    static constexpr bool kBatched = true;
    uint32_t k0, k1, rep;
    unsigned long long attempt;
    uint32_t pos;            // next draw index within the attempt's stream
    uint32_t b0, b1, b2, b3; // block pos >> 2 (valid when (pos & 3) != 0)

    // out of line: ~70 instructions that would otherwise be inlined at every draw site
    __device__ CB_NOINLINE void refill() {
        uint32_t o[4];
        philox4x32_10((uint32_t)attempt, (uint32_t)(attempt >> 32), pos >> 2, rep, k0, k1, o);
        b0 = o[0];
        b1 = o[1];
        b2 = o[2];
        b3 = o[3];
    }
    __device__ __forceinline__ void seek_attempt(unsigned long long t, uint32_t p = 0) {
        attempt = t;
        pos = p;
        if (p & 3u) refill();
    }
    __device__ __forceinline__ uint32_t position() const { return pos; }
    __device__ __forceinline__ uint32_t next32() {
        const uint32_t w = pos & 3u;
        if (w == 0) refill();
        ++pos;
        return w == 0 ? b0 : w == 1 ? b1 : w == 2 ? b2 : b3;
    }
    __device__ __forceinline__ uint32_t next31() { return next32() >> 1; }
    __device__ __forceinline__ double uniform() { return (double)next31() / CB_RAND_MAX; }
    __device__ __forceinline__ int randint(int high) {
        uint32_t rng = (uint32_t)(high - 1);
        if (rng == 0) return 0;
        uint32_t mask = rng;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        do {
            v = next32() & mask;
        } while (v > rng);
        return (int)v;
    }
    __device__ __forceinline__ void save(uint32_t *dst) const {
        dst[0] = (uint32_t)attempt;
        dst[1] = (uint32_t)(attempt >> 32);
        dst[2] = pos;
    }
    __device__ __forceinline__ void restore(const uint32_t *src) {
        seek_attempt((unsigned long long)src[0] | ((unsigned long long)src[1] << 32), src[2]);
    }
};

// host-side glibc srand() restatement used by chromo_srand
static inline void glibc_srand_host(uint32_t *st, uint32_t seed) {
    int32_t r[34];
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        long long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    uint32_t *u = (uint32_t *)r;
    int f = 3, b = 0;
    for (int i = 0; i < 310; i++) {
        u[f] += u[b];
        if (++f >= 31) f = 0;
        if (++b >= 31) b = 0;
    }
    for (int i = 0; i < 31; i++) st[i] = u[i];
    st[31] = (uint32_t)f;
    st[32] = (uint32_t)b;
}

static inline void mt_seed_host(uint32_t *mt, uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    mt[624] = 624;
}
