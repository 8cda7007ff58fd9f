// lw_mtjump.cpp -- host side of the Mersenne-Twister jump-ahead (irng = 1 path of the McICA generator).
//
// The reference draws ONE MT19937 stream per mcica_subcol_lw call in the order (sub-column, column, layer)
// (src/mcica_subcol_gen_lw.f90:478, 523-595; generator src/mcica_random_numbers.f90:116-185).  To generate the 140
// sub-column runs of a column block side by side, each generator CTA has to start n words into the stream.  MT19937
// is a linear recurrence over GF(2) with a primitive characteristic polynomial phi of degree 19937, so with
//     g(x) = x^n mod phi(x)
// every word of the stream obeys  x[n + j] = XOR over the set coefficients g_i of x[i + j]  (F^n = g(F) applied to
// the state, read out at word j): the window that starts n words ahead is 624 GF(2) dot products of g with the next
// 19937 + 623 words.  This file builds, once per process and from nothing but the generator itself,
//   * phi, by Berlekamp-Massey on one output bit stream of the generator (no constants to get wrong), and
//   * the table  x^(2^b) mod phi, b = 0 .. MTJ_NPOW-1  (repeated squaring),
// so that the device jumps by an arbitrary n as a product of the table entries of n's set bits
// (mt_jump_kernel in lw_kernels.cu).  mt_jump_selftest() checks the whole chain against plain stepping on the host.
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "lw_launch.h"

namespace lw {
namespace {

constexpr int DEG = 19937;
constexpr int NW64 = 320;            // 64-bit words of a polynomial of degree < 20480
constexpr int NW64_2 = 2 * NW64;     // of a square

struct HostMt {  // raw (untempered) word stream x[0], x[1], ...; x[0..623] = the seeded array
    uint32_t w[624];
    int pos = 0;  // index of the oldest word inside the circular window
    void seed(uint32_t s) {  // initialize_scalar, random_numbers.f90:169-185
        w[0] = s;
        for (int i = 1; i < 624; ++i) { s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i; w[i] = s; }
        pos = 0;
    }
    uint32_t at(int k) const { return w[(pos + k) % 624]; }
    // appends x[t + 624] and drops x[t]; returns the new word
    uint32_t step() {
        const uint32_t u = at(0), v = at(1), m = at(397);
        const uint32_t mix = (u & 0x80000000u) | (v & 0x7fffffffu);
        const uint32_t nw = m ^ (mix >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
        w[pos] = nw;
        pos = (pos + 1) % 624;
        return nw;
    }
};

inline int parity64(uint64_t v) { return __builtin_parityll(v); }

// phi(x) = sum c_i x^(L - i) from the connection polynomial C of the bit stream (Berlekamp-Massey over GF(2))
bool characteristic_polynomial(std::vector<uint64_t>& phi) {
    HostMt g;
    g.seed(4357u);
    const int N = 2 * DEG + 128;
    std::vector<uint64_t> C(NW64, 0), Bs(NW64, 0), R(NW64, 0), T(NW64);
    C[0] = 1;
    Bs[0] = 2;  // B = 1, shifted by m = 1
    int L = 0;
    for (int i = 0; i < 624; ++i) g.step();  // start from generated words only
    for (int n = 0; n < N; ++n) {
        const int bit = (int)(g.step() & 1u);
        // R: bit i = s[n - i]
        uint64_t carry = (uint64_t)bit;
        const int rw = (n >> 6) + 1 < NW64 ? (n >> 6) + 1 : NW64;
        for (int k = 0; k < rw; ++k) {
            const uint64_t nc = R[k] >> 63;
            R[k] = (R[k] << 1) | carry;
            carry = nc;
        }
        const int lw = (L >> 6) + 1;
        uint64_t acc = 0;
        for (int k = 0; k < lw; ++k) acc ^= C[k] & R[k];
        const int d = parity64(acc);
        if (d) {
            if (2 * L <= n) {
                T = C;
                for (int k = 0; k < NW64; ++k) C[k] ^= Bs[k];
                L = n + 1 - L;
                Bs = T;  // B = old C, then the shift below makes it B x^1
            } else {
                for (int k = 0; k < NW64; ++k) C[k] ^= Bs[k];
            }
        }
        uint64_t c2 = 0;
        for (int k = 0; k < NW64; ++k) {
            const uint64_t nc = Bs[k] >> 63;
            Bs[k] = (Bs[k] << 1) | c2;
            c2 = nc;
        }
    }
    if (L != DEG) return false;
    phi.assign(NW64, 0);
    for (int i = 0; i <= DEG; ++i)
        if ((C[i >> 6] >> (i & 63)) & 1ull) { const int e = DEG - i; phi[e >> 6] |= 1ull << (e & 63); }
    return ((phi[DEG >> 6] >> (DEG & 63)) & 1ull) != 0;
}

struct Table {
    bool ok = false;
    std::vector<uint32_t> pow2;  // MTJ_NPOW x 624 words: bit i of entry b = coefficient of x^i in x^(2^b) mod phi
};

void build_table(Table& tb) {
    std::vector<uint64_t> phi;
    if (!characteristic_polynomial(phi)) return;
    // phi << s for s = 0..63 (NW64 + 1 words each)
    std::vector<uint64_t> sh((size_t)64 * (NW64 + 1), 0);
    for (int s = 0; s < 64; ++s) {
        uint64_t* d = &sh[(size_t)s * (NW64 + 1)];
        for (int k = 0; k < NW64; ++k) {
            d[k] |= phi[k] << s;
            if (s) d[k + 1] |= phi[k] >> (64 - s);
        }
    }
    const int PW = (DEG >> 6) + 2;  // words of phi << s that can be non-zero
    tb.pow2.assign((size_t)MTJ_NPOW * 624, 0u);
    std::vector<uint64_t> p(NW64, 0), sq(NW64_2 + 2, 0);
    p[0] = 2;  // x
    for (int b = 0; b < MTJ_NPOW; ++b) {
        uint32_t* out = &tb.pow2[(size_t)b * 624];
        for (int k = 0; k < 312; ++k) { out[2 * k] = (uint32_t)p[k]; out[2 * k + 1] = (uint32_t)(p[k] >> 32); }
        // square: bit i -> bit 2 i
        std::fill(sq.begin(), sq.end(), 0);
        for (int k = 0; k < 312; ++k) {
            uint64_t lo = 0, hi = 0;
            const uint64_t v = p[k];
            for (int i = 0; i < 32; ++i) lo |= ((v >> i) & 1ull) << (2 * i);
            for (int i = 0; i < 32; ++i) hi |= ((v >> (32 + i)) & 1ull) << (2 * i);
            sq[2 * k] = lo; sq[2 * k + 1] = hi;
        }
        // reduce modulo phi, top bit down
        for (int i = 2 * (DEG - 1); i >= DEG; --i) {
            if (!((sq[i >> 6] >> (i & 63)) & 1ull)) continue;
            const int s = i - DEG;
            const uint64_t* d = &sh[(size_t)(s & 63) * (NW64 + 1)];
            uint64_t* r = &sq[s >> 6];
            for (int k = 0; k < PW; ++k) r[k] ^= d[k];
        }
        for (int k = 0; k < NW64; ++k) p[k] = k < 312 ? sq[k] : 0;
    }
    tb.ok = true;
}

Table& table() {
    static Table tb;
    static std::once_flag once;
    std::call_once(once, [] { build_table(tb); });
    return tb;
}

}  // namespace

const unsigned* mt_jump_pow2_table() {
    Table& tb = table();
    return tb.ok ? tb.pow2.data() : nullptr;
}

// Host check of the jump algebra: the window n words ahead obtained through the table (one GF(2) dot-product pass per
// set bit of n, exactly what the device kernel does) against stepping the generator n times.  Returns the number of
// mismatching words (bit 31 only for word 0: its low 31 bits are not part of the generator state), -1 = no table.
int mt_jump_selftest() {
    const uint32_t* pw = mt_jump_pow2_table();
    if (!pw) return -1;
    int bad = 0;
    const unsigned long long dist[] = {1ull, 623ull, 624ull, 10080ull, 1000003ull};
    for (unsigned long long n : dist) {
        HostMt ref;
        ref.seed(140u);
        for (unsigned long long k = 0; k < n; ++k) ref.step();
        HostMt cur;
        cur.seed(140u);
        std::vector<uint32_t> xs(DEG + 624), y(624);
        for (int b = 0; (n >> b) != 0ull; ++b) {
            if (!((n >> b) & 1ull)) continue;
            for (int k = 0; k < 624; ++k) xs[k] = cur.at(k);
            HostMt run = cur;
            for (int k = 624; k < DEG + 624; ++k) xs[k] = run.step();
            const uint32_t* g = pw + (size_t)b * 624;
            for (int j = 0; j < 624; ++j) {
                uint32_t acc = 0;
                for (int i = 0; i < DEG; ++i)
                    if ((g[i >> 5] >> (i & 31)) & 1u) acc ^= xs[i + j];
                y[j] = acc;
            }
            for (int k = 0; k < 624; ++k) cur.w[k] = y[k];
            cur.pos = 0;
        }
        bad += ((cur.at(0) ^ ref.at(0)) & 0x80000000u) != 0u;
        for (int k = 1; k < 624; ++k) bad += cur.at(k) != ref.at(k);
    }
    return bad;
}

}  // namespace lw
