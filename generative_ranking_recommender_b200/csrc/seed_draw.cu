// Host side of KMeans.initialize (balancekmeans/__init__.py:247-255): the reference seeds the centroids with
// np.random.choice(N, K, replace=False) on NumPy's GLOBAL legacy generator, which permutes all N row numbers
// (SURVEY.md H7).  To stay on the reference's random stream the draw has to consume the MT19937 state exactly as
// NumPy's legacy RandomState.permutation does; doing it here (plain C, no Python objects) takes a fifth of NumPy's
// time and, called through ctypes, runs without the interpreter lock, so it overlaps the GPU iterations.
//
// Algorithm (NumPy legacy, numpy/random/mtrand.pyx + src/distributions/distributions.c):
//   permutation(n): a = arange(n); for i = n-1 .. 1: j = random_interval(i); swap(a[i], a[j])
//   random_interval(max): mask = smallest 2^b - 1 >= max; draw 32-bit (max <= 2^32-1) or 64-bit words, masked,
//                         until the value is <= max
//   MT19937: the standard generator; a 64-bit word is (hi << 32) | lo of two consecutive 32-bit outputs
// choice(n, k, replace=False) returns the first k elements.  tests/test_host_logic.py pins this against NumPy.
#include "common.cuh"

namespace rqk {
namespace {
struct Mt {
    uint32_t* key;   // [624]
    int pos;
};
inline void mt_gen(Mt& s) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX = 0x9908b0dfu;
    uint32_t* k = s.key;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; i++) {
        y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX);
    }
    for (; i < 623; i++) {
        y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX);
    }
    y = (k[623] & UPPER) | (k[0] & LOWER);
    k[623] = k[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX);
    s.pos = 0;
}
inline uint32_t mt_next32(Mt& s) {
    if (s.pos == 624) mt_gen(s);
    uint32_t y = s.key[s.pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
inline uint64_t mt_next64(Mt& s) {
    const uint64_t hi = mt_next32(s);
    return (hi << 32) | mt_next32(s);
}
inline uint64_t legacy_interval(Mt& s, uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max, value;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    if (max <= 0xffffffffull) {
        while ((value = (mt_next32(s) & mask)) > max) {}
    } else {
        while ((value = (mt_next64(s) & mask)) > max) {}
    }
    return value;
}
}  // namespace
}  // namespace rqk

extern "C" {

// key: uint32[624] MT19937 state words, *pos: position (both from np.random.get_state(), updated in place).
// scratch: int64[n] (host).  out: int64[k] = np.random.choice(n, k, replace=False) on that state.
int rqk_legacy_choice(uint32_t* key, int32_t* pos, int64_t n, int64_t k, int64_t* scratch, int64_t* out) {
    using namespace rqk;
    if (!key || !pos || !scratch || !out || n < 1 || k < 0 || k > n || *pos < 0 || *pos > 624)
        return fail(RQK_ERR_ARG, "rqk_legacy_choice: bad argument%s");
    Mt s{key, *pos};
    if (n <= 0x7fffffffll) {          // 32-bit row numbers: half the cache footprint of the permutation
        uint32_t* a = reinterpret_cast<uint32_t*>(scratch);
        for (int64_t i = 0; i < n; ++i) a[i] = (uint32_t)i;
        for (int64_t i = n - 1; i >= 1; --i) {
            const int64_t j = (int64_t)legacy_interval(s, (uint64_t)i);
            const uint32_t t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
        for (int64_t i = 0; i < k; ++i) out[i] = (int64_t)a[i];
    } else {
        for (int64_t i = 0; i < n; ++i) scratch[i] = i;
        for (int64_t i = n - 1; i >= 1; --i) {
            const int64_t j = (int64_t)legacy_interval(s, (uint64_t)i);
            const int64_t t = scratch[i];
            scratch[i] = scratch[j];
            scratch[j] = t;
        }
        for (int64_t i = 0; i < k; ++i) out[i] = scratch[i];
    }
    *pos = s.pos;
    return 0;
}

}  // extern "C"
