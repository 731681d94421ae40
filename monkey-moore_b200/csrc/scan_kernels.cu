// scan_kernels.cu -- sm_100a kernels of the relative-search path (see scan_kernels.cuh for the plan).
//
// Reference semantics reproduced here (file:line under /root/reference/):
//   window comparison + advance, simple / value scan   src/core/monkey_moore.cpp:347-405
//   window comparison + advance, wildcard              src/core/monkey_moore.cpp:449-541
//   per-block / per-alignment element views            src/core/search_engine.cpp:129-159
//   endianness normalisation (a byte permute on load)  include/mmoore/byteswap.hpp:70-79
#include "scan_kernels.cuh"

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#define FULL 0xFFFFFFFFu

namespace {

// ------------------------------------------------------------------------------------------
// exact window evaluation (the slow path; also the whole of the generic path)
// ------------------------------------------------------------------------------------------

template <int W, bool BE>
__device__ __forceinline__ uint32_t ld_elem(const uint8_t *p) {
    if (W == 1) return p[0];
    uint32_t a = p[0], b = p[1];
    return BE ? ((a << 8) | b) : ((b << 8) | a);
}

// F(s): returns the advance in bits [7:0] and 0x100 when the window matches.
template <int W, bool BE>
__device__ __noinline__ uint32_t eval_window(const MmgProgram &P, const uint8_t *w) {
    const uint32_t vmask = W == 1 ? 0xFFu : 0xFFFFu;
    for (int c = 0; c < P.ncheck; c++) {
        const int i = P.chk[c].i;
        const int cur = (int)ld_elem<W, BE>(w + i * W);
        const int prv = (int)ld_elem<W, BE>(w + (i - P.chk[c].lag) * W);
        const int d = cur - prv;
        const int ed = P.chk[c].ed;
        const bool pass = P.modular ? ((((uint32_t)(d - ed)) & vmask) == 0) : (d == ed);
        if (!pass) {
            int sk = P.tab_default;
            for (int j = 0; j < P.ntab; j++)
                if (P.tab_key[j] == d) sk = P.tab_val[j];
            return (uint32_t)min(P.chk[c].cap, sk);
        }
    }
    return 0x100u | (uint32_t)P.match_jump;
}

// ------------------------------------------------------------------------------------------
// K1: streaming filter
// ------------------------------------------------------------------------------------------

// 32-bit word starting at byte offset OFF of the 32-byte window x[0..7] (x[0..3]: the 16 bytes
// before this lane's, x[4..7]: this lane's).  BE swaps the bytes of each 16-bit half.
template <int OFF, bool BE>
__device__ __forceinline__ uint32_t extract(const uint32_t (&x)[8]) {
    constexpr int q = OFF >> 2, r = OFF & 3;
    if (!BE) {
        if (r == 0) return x[q];
        return __funnelshift_r(x[q], x[q + 1], 8 * r);
    }
    constexpr uint32_t sel = (uint32_t)(r + 1) | ((uint32_t)r << 4) | ((uint32_t)(r + 3) << 8) | ((uint32_t)(r + 2) << 12);
    return __byte_perm(x[q], r == 0 ? 0u : x[q + 1], sel);
}

// Per-lane candidate detection.  Returns true when any of this lane's 16 positions is flagged;
// f[] receives what slow_row() needs to name the positions.
//   W=1: f[k]    bit 8j+7 set  <=> byte position 4k+j flagged (current element at that byte)
//   W=2: f[c*4+k] has a zero 16-bit half h  <=> the element starting at byte 4k-c+2h is flagged
// NK > 0: number of keys known at compile time (fully unrolled); NK == 0: P.nkeys at run time.
template <int W, int LB, bool BE, int NK>
__device__ __forceinline__ bool filter_lane(const MmgProgram &P, const uint32_t (&x)[8], uint32_t (&f)[8], bool depth2 = false) {
    if (LB == 0) return true;   // evaluate-everything mode
    const int nk = NK > 0 ? NK : P.nkeys;
    if (W == 1) {
        return true;    // 8-bit searches use filter8() below
    } else {
        // t = cur + ~prev = (cur - prev - 1) per 16-bit half; key constant C = 1 - key;
        // min-accumulate t + C: a zero half <=> difference == key
        uint32_t t[8];
        t[0] = __vadd2(extract<16, BE>(x), ~extract<16 - LB, BE>(x));
        t[1] = __vadd2(extract<20, BE>(x), ~extract<20 - LB, BE>(x));
        t[2] = __vadd2(extract<24, BE>(x), ~extract<24 - LB, BE>(x));
        t[3] = __vadd2(extract<28, BE>(x), ~extract<28 - LB, BE>(x));
        t[4] = __vadd2(extract<15, BE>(x), ~extract<15 - LB, BE>(x));
        t[5] = __vadd2(extract<19, BE>(x), ~extract<19 - LB, BE>(x));
        t[6] = __vadd2(extract<23, BE>(x), ~extract<23 - LB, BE>(x));
        t[7] = __vadd2(extract<27, BE>(x), ~extract<27 - LB, BE>(x));
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = 0xFFFFFFFFu;
        for (int j = 0; j < nk; j++) {
            const uint32_t key = P.keys[j];
#pragma unroll
            for (int k = 0; k < 8; k++) f[k] = __viaddmin_u16x2(t[k], key, f[k]);
        }
        uint32_t m = __vminu2(__vminu2(__vminu2(f[0], f[1]), __vminu2(f[2], f[3])),
                              __vminu2(__vminu2(f[4], f[5]), __vminu2(f[6], f[7])));
        return ((m & 0xFFFFu) == 0) || ((m >> 16) == 0);
    }
}


// ------------------------------------------------------------------------------------------
// 8-bit filter: the lane's 16 elements (and the 16 before them) are unpacked into 16-bit pairs so that the
// differences are EXACT signed values (one IADD3 per pair, bias 256 keeps both halves positive) and one
// VIADDMNMX.U16x2 per key and pair min-accumulates "difference - key": a zero half <=> that element's
// comparison-0 difference is a key.  Pair registers: E[q] = elements (4q, 4q+2) of word q, O[q] = (4q+1, 4q+3).
// ------------------------------------------------------------------------------------------

template <int POS>
__device__ __forceinline__ uint32_t pair_at(const uint32_t (&E)[8], const uint32_t (&O)[8]) {
    // elements at byte POS and POS + 2 of the 32-byte window, as (low half, high half)
    constexpr int q = POS >> 2, r = POS & 3;
    if (r == 0) return E[q];
    if (r == 1) return O[q];
    if (r == 2) return __funnelshift_r(E[q], E[q + 1 < 8 ? q + 1 : q], 16);
    return __funnelshift_r(O[q], O[q + 1 < 8 ? q + 1 : q], 16);
}

template <int LB, int POS>
__device__ __forceinline__ uint32_t diff_at(const uint32_t (&E)[8], const uint32_t (&O)[8]) {
    return pair_at<POS>(E, O) - pair_at<POS - LB>(E, O) + 0x01000100u;      // halves: 256 + (cur - prv), never a borrow
}

// a - b as one IMAD (FMA pipe: the ALU pipe is the kernel's bottleneck)
__device__ __forceinline__ uint32_t sub_fma(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 0xFFFFFFFF, %2;" : "=r"(r) : "r"(b), "r"(a));
    return r;
}

// Returns the 16-bit candidate mask of the lane: bit e <=> the element at the lane's own byte e is flagged.
// NK > 0: compile-time key count.  DEPTH2 (simple / value-scan patterns, P.d2ok): a window whose comparison 0
// PASSES (difference == keys[0]) is kept only if its comparison 1 -- the difference one element earlier -- is
// a key as well; every other such window advances by J0 without a match, i.e. is not an event.
template <int LB, int NK, bool DEPTH2>
__device__ __forceinline__ uint32_t filter8(const MmgProgram &P, const uint32_t *x) {
    // Odd lags: the pairs of odd bytes are unpacked WITH the bias (the PRMT takes the 0x01 bytes from its second source),
    // so "current - previous" is a bare subtraction with a known constant in each half -- 256 + d for a biased current
    // pair, d - 256 and d - 257 (the low half always borrows) for an unbiased one -- and runs as an IMAD on the idle FMA
    // pipe.  The two kinds of difference register then use two sets of key constants (keys / pkeys).
    constexpr bool FMA_DIFF = (LB & 1) != 0;
    uint32_t E[8], O[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { E[q] = x[q] & 0x00FF00FFu; O[q] = __byte_perm(x[q], FMA_DIFF ? 0x01010101u : 0u, 0x4341u); }
    // D[1 + 2q + t]: pair (4q + t, 4q + t + 2) of the lane's own bytes; D[0]: pair (-3, -1) (DEPTH2 only)
    uint32_t D[9];
    if (FMA_DIFF) {
        D[1] = sub_fma(pair_at<16>(E, O), pair_at<16 - LB>(E, O)); D[2] = sub_fma(pair_at<17>(E, O), pair_at<17 - LB>(E, O));
        D[3] = sub_fma(pair_at<20>(E, O), pair_at<20 - LB>(E, O)); D[4] = sub_fma(pair_at<21>(E, O), pair_at<21 - LB>(E, O));
        D[5] = sub_fma(pair_at<24>(E, O), pair_at<24 - LB>(E, O)); D[6] = sub_fma(pair_at<25>(E, O), pair_at<25 - LB>(E, O));
        D[7] = sub_fma(pair_at<28>(E, O), pair_at<28 - LB>(E, O)); D[8] = sub_fma(pair_at<29>(E, O), pair_at<29 - LB>(E, O));
        if (DEPTH2) D[0] = sub_fma(pair_at<13>(E, O), pair_at<13 - LB>(E, O));
    } else {
        D[1] = diff_at<LB, 16>(E, O); D[2] = diff_at<LB, 17>(E, O); D[3] = diff_at<LB, 20>(E, O); D[4] = diff_at<LB, 21>(E, O);
        D[5] = diff_at<LB, 24>(E, O); D[6] = diff_at<LB, 25>(E, O); D[7] = diff_at<LB, 28>(E, O); D[8] = diff_at<LB, 29>(E, O);
        if (DEPTH2) D[0] = diff_at<LB, 13>(E, O);
    }
    // key constant of difference register k: odd k holds an unbiased current pair (bytes 4q, 4q + 2)
#define KEY8(j, k) ((FMA_DIFF && ((k) & 1)) ? P.pkeys[j] : P.keys[j])
    const int nk = NK > 0 ? NK : P.nkeys;
    uint32_t c[9];          // per half: 0 <=> candidate, else 1 (the min-accumulation starts from 1)
    if (!DEPTH2) {
#pragma unroll
        for (int k = 1; k < 9; k++) c[k] = __viaddmin_u16x2(D[k], KEY8(0, k), 0x00010001u);      // per-half add; halves stay in {0, 1}
        if (NK > 0) {
#pragma unroll
            for (int j = 1; j < NK; j++) {
#pragma unroll
                for (int k = 1; k < 9; k++) c[k] = __viaddmin_u16x2(D[k], KEY8(j, k), c[k]);
            }
        } else {
#pragma unroll 1
            for (int j = 1; j < nk; j++) {
#pragma unroll
                for (int k = 1; k < 9; k++) c[k] = __viaddmin_u16x2(D[k], KEY8(j, k), c[k]);
            }
        }
    } else {
        uint32_t ap[9], ao[9];      // pass key / the other keys
#pragma unroll
        for (int k = 0; k < 9; k++) { ap[k] = __viaddmin_u16x2(D[k], KEY8(0, k), 0x00010001u); ao[k] = 0x00010001u; }
        if (NK > 0) {
#pragma unroll
            for (int j = 1; j < NK; j++) {
#pragma unroll
                for (int k = 0; k < 9; k++) ao[k] = __viaddmin_u16x2(D[k], KEY8(j, k), ao[k]);
            }
        } else {
#pragma unroll 1
            for (int j = 1; j < nk; j++) {
#pragma unroll
                for (int k = 0; k < 9; k++) ao[k] = __viaddmin_u16x2(D[k], KEY8(j, k), ao[k]);
            }
        }
        // any[k]: zero half <=> that element's difference is some key
        uint32_t any[9];
#pragma unroll
        for (int k = 0; k < 9; k++) any[k] = NK == 1 ? ap[k] : __vminu2(ap[k], ao[k]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // the element before (4q, 4q+2) is (4q-1, 4q+1): upper half of the O pair of word q-1, lower half of word q's
            const uint32_t before_e = __funnelshift_r(any[2 * q], any[2 * q + 2], 16);
            const uint32_t before_o = any[2 * q + 1];                            // before (4q+1, 4q+3) is (4q, 4q+2)
            const uint32_t pe = __vmaxu2(ap[2 * q + 1], before_e), po = __vmaxu2(ap[2 * q + 2], before_o);
            c[2 * q + 1] = NK == 1 ? pe : __vminu2(ao[2 * q + 1], pe);
            c[2 * q + 2] = NK == 1 ? po : __vminu2(ao[2 * q + 2], po);
        }
    }
    // halves (0: candidate, 1: not) -> bits, weighted so that bit e (and 14 + e for the upper halves) names element e
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        m += c[2 * q + 1] << (4 * q);
        m += c[2 * q + 2] << (4 * q + 1);
    }
#undef KEY8
    return ((m | (m >> 14)) & 0xFFFFu) ^ 0xFFFFu;
}

// 16-bit candidate mask in ascending byte order.  Bit b names the element whose first byte is
//   W=1: 16*lane + b            W=2: 16*lane + b - 1
template <int W, int LB>
__device__ __forceinline__ uint32_t candidate_mask(const uint32_t (&f)[8], bool any) {
    if (LB == 0) return 0xFFFFu;
    if (!any) return 0;
    uint32_t cm = 0;
    if (W == 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) cm |= ((((f[k] >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * k);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t e = f[k], o = f[4 + k];      // even class (c=0), odd class (c=1)
            cm |= (uint32_t)((o & 0xFFFFu) == 0) << (4 * k + 0);
            cm |= (uint32_t)((e & 0xFFFFu) == 0) << (4 * k + 1);
            cm |= (uint32_t)((o >> 16) == 0) << (4 * k + 2);
            cm |= (uint32_t)((e >> 16) == 0) << (4 * k + 3);
        }
    }
    return cm;
}

// Cheap 16-bit pre-test: is ANY of this lane's 16 element positions possibly flagged?  The element
// differences are formed with one 32-bit subtraction per word (the borrow between the two halves can
// only turn a hit in the upper half into 1 or 2 instead of 0), then min-accumulated into a single
// register.  A superset of filter_lane<2,...>: false positives ~2^-15 per position, no false negatives.
template <int LB, bool BE, int NK>
__device__ __forceinline__ bool prefilter16(const MmgProgram &P, const uint32_t (&x)[8]) {
    uint32_t cur[8], prv[8];
    cur[0] = extract<16, BE>(x); prv[0] = extract<16 - LB, BE>(x);
    cur[1] = extract<20, BE>(x); prv[1] = extract<20 - LB, BE>(x);
    cur[2] = extract<24, BE>(x); prv[2] = extract<24 - LB, BE>(x);
    cur[3] = extract<28, BE>(x); prv[3] = extract<28 - LB, BE>(x);
    cur[4] = extract<15, BE>(x); prv[4] = extract<15 - LB, BE>(x);
    cur[5] = extract<19, BE>(x); prv[5] = extract<19 - LB, BE>(x);
    cur[6] = extract<23, BE>(x); prv[6] = extract<23 - LB, BE>(x);
    cur[7] = extract<27, BE>(x); prv[7] = extract<27 - LB, BE>(x);
    uint32_t acc = 0xFFFFFFFFu;
    if (NK == 1) {
        const uint32_t c = P.pkeys[0];          // upper half 1 - key, lower half -key
#pragma unroll
        for (int k = 0; k < 8; k += 2)
            acc = __vimin3_u16x2(acc, cur[k] - prv[k] + c, cur[k + 1] - prv[k + 1] + c);
    } else {
        uint32_t d[8];
#pragma unroll
        for (int k = 0; k < 8; k++) d[k] = cur[k] - prv[k];
        if (P.rng_w != 0xFFFFFFFFu) {
            // range stage: does any difference of the row fall into the arc that holds all keys?  (upper halves carry
            // the borrow slack of the 32-bit subtraction: + 1)
            const uint32_t c = P.rng_c;
            uint32_t a0 = __viaddmin_u16x2(d[0], c, 0xFFFFFFFFu), a1 = __viaddmin_u16x2(d[1], c, 0xFFFFFFFFu);
            a0 = __viaddmin_u16x2(d[2], c, a0); a1 = __viaddmin_u16x2(d[3], c, a1);
            a0 = __viaddmin_u16x2(d[4], c, a0); a1 = __viaddmin_u16x2(d[5], c, a1);
            a0 = __viaddmin_u16x2(d[6], c, a0); a1 = __viaddmin_u16x2(d[7], c, a1);
            const uint32_t a = __vminu2(a0, a1);
            const bool maybe = ((a & 0xFFFFu) <= P.rng_w) || ((a >> 16) <= P.rng_w + 1u);
            if (!__any_sync(FULL, maybe)) return false;
        }
        const int nk = P.nkeys;
        uint32_t a4[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};      // four independent min chains
#pragma unroll 2
        for (int j = 0; j < nk; j++) {
            const uint32_t c = P.pkeys[j];
#pragma unroll
            for (int k = 0; k < 8; k++) a4[k & 3] = __viaddmin_u16x2(d[k], c, a4[k & 3]);
        }
        acc = __vimin3_u16x2(__vminu2(a4[0], a4[1]), a4[2], a4[3]);
    }
    return ((acc & 0xFFFFu) == 0) || ((acc >> 16) <= 2u);
}

struct WarpState {
    uint32_t open_t;       // sub-tile whose event list is being filled
    uint32_t open_start;   // event index where that list starts
    uint32_t cursor;       // next free event index of this warp's private region (unclamped)
    // extent words {first event, number of events} of the chunk's sub-tiles, lane k holding the one of sub-tile t0 + k:
    // they leave the warp in ONE coalesced store when the chunk is done (flush_extents) instead of one 8-byte store
    // per sub-tile as they close.
    uint32_t t0, ex, ey;
};

__device__ __forceinline__ WarpState close_until(const MmgScratch &X, WarpState st, uint32_t t, int lane) {
    // every sub-tile the warp passes gets its extent word -- a count of 0 where it has no events, so the resolve kernels
    // read the extents without any zeroing by the host
    if (st.open_t < t) {
        const uint32_t k = st.open_t - st.t0, ke = t - st.t0;
        if ((uint32_t)lane == k) { st.ex = st.open_start; st.ey = st.cursor - st.open_start; }
        else if ((uint32_t)lane > k && (uint32_t)lane < ke) { st.ex = 0u; st.ey = 0u; }      // passed without events
        st.open_start = st.cursor;
        st.open_t = t;
    }
    return st;
}

__device__ __forceinline__ void flush_extents(const MmgScratch &X, const WarpState &st, uint32_t t1, int lane) {
    if ((uint32_t)lane < t1 - st.t0) X.ext[st.t0 + lane] = make_uint2(st.ex, st.ey);
}

// The filter kernel keeps the part of the pattern program that exact evaluation needs in shared memory
// (file-scope __shared__: every device function reaches it with plain LDS instead of generic loads from
// the kernel-parameter space).
struct SProg {
    int32_t ncheck, ntab, tab_default, match_jump, J0, modular;
    int16_t ci[MMG_MAXL], clag[MMG_MAXL];
    int32_t ced[MMG_MAXL], ccap[MMG_MAXL];
    int32_t tkey[MMG_MAXL], tval[MMG_MAXL];
    uint8_t tab8[512];      // 8-bit searches: skip for every possible difference d, indexed d + 255
    uint8_t tab0[512];      // ditto with the cap of comparison 0 applied: the advance when comparison 0 fails on d
    uint8_t tab1[512];      // ditto for comparison 1
    // scalars of the 8-bit candidate evaluation (read once per thread with volatile loads: they then live in
    // registers instead of being re-derived from the kernel parameters inside the hot loop)
    uint32_t sg, pmask, o1c, o1p, matchword;
    int32_t ed0, ed1, L;
};
__shared__ SProg g_sprog;

__device__ __forceinline__ void load_sprog(const MmgProgram &P) {
    if (threadIdx.x == 0) {
        g_sprog.ncheck = P.ncheck; g_sprog.ntab = P.ntab; g_sprog.tab_default = P.tab_default;
        g_sprog.match_jump = P.match_jump; g_sprog.J0 = P.J0; g_sprog.modular = P.modular;
        g_sprog.sg = (uint32_t)P.chk[0].i * (uint32_t)P.W;
        g_sprog.pmask = P.modular ? (P.W == 1 ? 0xFFu : 0xFFFFu) : 0xFFFFFFFFu;
        g_sprog.o1c = (uint32_t)(P.chk[0].i - P.chk[1].i);
        g_sprog.o1p = (uint32_t)(P.chk[0].i - P.chk[1].i + P.chk[1].lag);
        g_sprog.matchword = 0x100u | (uint32_t)P.match_jump;
        g_sprog.ed0 = P.chk[0].ed; g_sprog.ed1 = P.chk[1].ed; g_sprog.L = P.L;
    }
    for (int i = threadIdx.x; i < P.ncheck; i += blockDim.x) {
        g_sprog.ci[i] = P.chk[i].i; g_sprog.clag[i] = P.chk[i].lag; g_sprog.ced[i] = P.chk[i].ed; g_sprog.ccap[i] = P.chk[i].cap;
    }
    for (int i = threadIdx.x; i < P.ntab; i += blockDim.x) { g_sprog.tkey[i] = P.tab_key[i]; g_sprog.tval[i] = P.tab_val[i]; }
    if (P.W == 1) {
        for (int i = threadIdx.x; i < 511; i += blockDim.x) {
            int sk = P.tab_default;
            for (int j = 0; j < P.ntab; j++)
                if (P.tab_key[j] == i - 255) sk = P.tab_val[j];
            g_sprog.tab8[i] = (uint8_t)sk;
            g_sprog.tab0[i] = (uint8_t)min(sk, P.ncheck > 0 ? P.chk[0].cap : sk);
            g_sprog.tab1[i] = (uint8_t)min(sk, P.ncheck > 1 ? P.chk[1].cap : sk);
        }
    }
    __syncthreads();
}

// F(s) from the shared-memory program: advance in bits [7:0], 0x100 when the window matches.
template <int W, bool BE>
__device__ __forceinline__ uint32_t eval_window_s(const uint8_t *w) {
    const uint32_t vmask = W == 1 ? 0xFFu : 0xFFFFu;
    const int nc = g_sprog.ncheck;
    for (int c = 0; c < nc; c++) {
        const int i = g_sprog.ci[c];
        const int cur = (int)ld_elem<W, BE>(w + i * W);
        const int prv = (int)ld_elem<W, BE>(w + (i - g_sprog.clag[c]) * W);
        const int d = cur - prv;
        const int ed = g_sprog.ced[c];
        const bool pass = g_sprog.modular ? ((((uint32_t)(d - ed)) & vmask) == 0) : (d == ed);
        if (!pass) {
            int sk;
            if (W == 1) {
                sk = g_sprog.tab8[d + 255];
            } else {
                sk = g_sprog.tab_default;
                const int nt = g_sprog.ntab;
#pragma unroll 1
                for (int j = 0; j < nt; j++)
                    if (g_sprog.tkey[j] == d) sk = g_sprog.tval[j];
            }
            return (uint32_t)min(g_sprog.ccap[c], sk);
        }
    }
    return 0x100u | (uint32_t)g_sprog.match_jump;
}

// per-chunk constants of the exact-evaluation path.  eval_batch is a real call (not inlined), so a ChunkCtx handed to
// it by reference has to live in memory: as a local variable that meant nine 8-byte local-memory stores per THREAD and
// chunk -- 2.3 KB per warp and chunk, 38 MB per 512 MiB scanned, which the input streaming through the L2 kept pushing
// out to DRAM (the "unexplained" 34 MB of DRAM writes of round 1; profiles/r2_dram_writes.txt).  The warp's copy now
// sits in shared memory, written once per chunk by one lane.
struct ChunkCtx {
    int64_t s_lo, s_hi;      // window starts owned by this chunk
    int64_t q_base;          // queue entries are (window start - q_base)
    int64_t blk_off;
    int64_t max_rel[2];      // last byte offset (from the block start) that begins a complete window, per alignment; -1: none
    const uint8_t *data;
    uint32_t *ev;
    uint32_t reg_hi;
    uint32_t npads;
};

// Exact evaluation of up to 32 queued candidate windows (one per lane, ascending window start)
// and ordered append of the resulting events to the warp's private event region.
#define MMG_FILTER_WARPS 8
__shared__ ChunkCtx g_ctx[MMG_FILTER_WARPS];

template <int W, bool BE>
__device__ __noinline__ WarpState eval_batch(const MmgScratch &X, WarpState st, const ChunkCtx &C, uint32_t entry, bool have,
                                             int lane) {
    bool is_event = false;
    uint32_t word = 0, ts = 0;
    if (have) {
        const int64_t s = C.q_base + (int64_t)entry;
        const int64_t rel = s - C.blk_off;
        const uint32_t pad = (uint32_t)rel & (uint32_t)(W - 1);
        if (s >= C.s_lo && s < C.s_hi && pad < C.npads && rel <= C.max_rel[pad]) {
            const uint32_t r = eval_window_s<W, BE>(C.data + s);
            const uint32_t jump = r & 0xFFu;
            if ((r & 0x100u) || jump != (uint32_t)g_sprog.J0) {     // default advance without a match is not an event
                is_event = true;
                ts = (uint32_t)((uint64_t)s >> MMG_SUBTILE_SHIFT);
                word = ((uint32_t)s & (MMG_SUBTILE - 1)) | (jump << 16) | ((r & 0x100u) ? MMG_EV_MATCH : 0u);
            }
        }
    }
    uint32_t pending = __ballot_sync(FULL, is_event);
    const uint32_t lt = (1u << lane) - 1u;
    while (pending) {                                        // one round per sub-tile present in the batch
        const uint32_t tt = __shfl_sync(FULL, ts, __ffs(pending) - 1);
        const uint32_t grp = __ballot_sync(FULL, is_event && ts == tt);
        st = close_until(X, st, tt, lane);
        if (is_event && ts == tt) {
            const uint32_t at = st.cursor + __popc(grp & lt);
            if (at < C.reg_hi) C.ev[at] = word;
        }
        st.cursor += __popc(grp);
        pending &= ~grp;
    }
    return st;
}

// ---- TMA (bulk async copy) + mbarrier helpers: one private ring per warp ------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// policy != 0: the lines of this copy are the first to leave the L2 again (createpolicy evict_first).  The input is
// read exactly once; without the hint it pushes everything else out of the 126 MB L2 as it streams through -- the
// event lists and extent words the resolve kernel is about to read, and dirty lines that then cost DRAM writes.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    if (policy)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t l2_stream_policy(uint32_t hint) {
    uint64_t p = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(p));
    return p;
}

#ifndef MMG_STAGE_BYTES
#define MMG_STAGE_BYTES 4096u                       // 8 rows (two 4 KiB stages per warp: half the per-stage bookkeeping of three 2 KiB stages, -7 %)
#endif
#define MMG_STAGE_STRIDE (MMG_STAGE_BYTES + 16u)    // + 16-byte left halo
#ifndef MMG_NSTAGES
#define MMG_NSTAGES 2
#endif
#ifndef MMG_FILTER_MIN_CTAS
#define MMG_FILTER_MIN_CTAS 3
#endif
#define MMG_QUEUE_CAP 320u                          // < 32 carried + <= 256 new candidates per half row (u16 entries)
#define MMG_WARP_SMEM ((MMG_NSTAGES * MMG_STAGE_STRIDE + MMG_QUEUE_CAP * 2u + MMG_NSTAGES * 8u + 15u) & ~15u)

// Stage `k` of a chunk holds the slice bytes [p0 + k*2048 - 16, p0 + (k+1)*2048), clipped to
// [0, copy_end) where copy_end is the 16-byte-aligned end of what the chunk needs; an unaligned tail
// of the slice (< 16 bytes) is patched in by the lanes.  All positions are relative to p0 (32 bit).
template <uint32_t STAGE = MMG_STAGE_BYTES>
__device__ __forceinline__ void issue_stage(const uint8_t *chunk_base, bool at_slice_start, uint32_t rel_stage,
                                            uint32_t copy_end_rel, uint32_t dst, uint32_t bar, uint64_t policy) {
    uint32_t skip = 0, lo = rel_stage - 16u;          // rel_stage == 0 && at_slice_start: no left halo exists
    if (rel_stage == 0 && at_slice_start) { skip = 16u; lo = 0; }
    const uint32_t hi = min(rel_stage + STAGE, copy_end_rel);
    if ((int32_t)(hi - lo) > 0) {
        const uint32_t bytes = hi - lo;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(dst + skip, chunk_base + (int32_t)lo, bytes, bar, policy);
    } else {
        mbar_arrive(bar);
    }
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}

__device__ __forceinline__ int lds32(uint32_t addr) {
    int r;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ int lds16s(uint32_t addr) {
    int r;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}

// F(s) of an 8-bit window that lies in shared memory at address wa, from comparison c0 on (the earlier ones
// passed).  sp: shared address of g_sprog.  32-bit shared addresses throughout.
__device__ __noinline__ uint32_t eval_window_lds8(uint32_t wa, uint32_t sp, int c0, uint32_t pmask) {
    const int nc = lds32(sp + (uint32_t)offsetof(SProg, ncheck));
    for (int c = c0; c < nc; c++) {
        const int i = lds16s(sp + (uint32_t)offsetof(SProg, ci) + 2u * c);
        const int lag = lds16s(sp + (uint32_t)offsetof(SProg, clag) + 2u * c);
        const int d = (int)lds8(wa + i) - (int)lds8(wa + i - lag);
        const int ed = lds32(sp + (uint32_t)offsetof(SProg, ced) + 4u * c);
        if ((((uint32_t)(d - ed)) & pmask) != 0u) {
            const int sk = (int)lds8(sp + (uint32_t)offsetof(SProg, tab8) + 255u + d);
            return (uint32_t)min(lds32(sp + (uint32_t)offsetof(SProg, ccap) + 4u * c), sk);
        }
    }
    return 0x100u | (uint32_t)lds32(sp + (uint32_t)offsetof(SProg, match_jump));
}

// record the extent of sub-tile t's event list: [st.open_start, cut)
__device__ __forceinline__ WarpState close_at(const MmgScratch &X, WarpState st, uint32_t t, uint32_t cut, int lane) {
    if ((uint32_t)lane == t - st.t0) { st.ex = st.open_start; st.ey = cut - st.open_start; }   // (every sub-tile gets one: nothing to zero before a scan)
    st.open_start = cut;
    st.open_t = t + 1;
    return st;
}

// ------------------------------------------------------------------------------------------
// Resolve FUSED into the filter kernels, for sparse scans (X.fuse).
// A pattern like cfg2's leaves a handful of events per 512 KiB engine block.  The general resolve kernel spends a CTA
// of 128 threads, some ten block-wide barriers and a decoupled look-back on every block whatever it holds, and costs
// a second launch.  Instead, the warps of the (persistent, fully resident, cooperatively launched) filter grid meet at a
// grid-wide barrier once the chunks are used up, and then share out the engine blocks: a warp reads the extent words of
// its block's sub-tiles (one round of loads), the few events behind them (a second round), replays the block's chains
// through them once -- positions are absolute within the block, so "is this event visited" is a plain lattice test and
// no per-sub-tile phase is needed -- and leaves the block's matches in a small per-block record.  The last warp to
// finish turns the per-block counts into output positions, writes the (few) matches in file order, hands the status
// words to the host and restores the workspace's zero state: the scan is ONE launch, and nothing is fenced or counted
// per chunk.  A block with more events, or more matches, than the staging area / its record holds raises a flag; the
// host then runs k_resolve over the same event lists.
// ------------------------------------------------------------------------------------------

// a filter warp ran out of its private event region: the host grows the buffer and re-runs
__device__ __forceinline__ bool events_overflowed(const MmgScratch &X) { return *reinterpret_cast<volatile uint64_t *>(X.status) > X.ev_per_warp; }

#define SPARSE_EV_CAP 1024u     // events of one engine block the resolving warp stages (in its idle TMA ring)
#define SPARSE_REC 32u          // words of a block record: [0] match count, [1..31] byte offsets of the matches in the block

// sm: >= SPARSE_EV_CAP words of this warp's shared memory
template <int W>
__device__ __noinline__ void sparse_block(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint32_t bi, uint32_t *sm,
                                          int lane) {
    uint32_t m = 0;                                       // matches of this block
    const uint32_t t0 = bi * G.spb;
    const uint32_t nsb = min(G.spb, G.nsub - t0);         // sub-tiles of this block (<= 128)
    // round 1: the extent words of the lane's four sub-tiles
    uint2 ex[4];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t j = 4u * (uint32_t)lane + (uint32_t)k;
        ex[k] = j < nsb ? __ldcg(X.ext + t0 + j) : make_uint2(0u, 0u);
        if ((uint64_t)ex[k].x + ex[k].y > X.ev_total) ex[k].y = SPARSE_EV_CAP + 1u;      // a warp overflowed its event region: the host re-runs
        mine += ex[k].y;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const uint32_t n = __shfl_sync(FULL, incl, 31);
    bool dense = n > SPARSE_EV_CAP;
    if (!dense && n != 0) {
        // round 2: the events, as  byte offset in the block [18:0] | advance [26:19] | match [27]
        uint32_t at = incl - mine;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t base = (4u * (uint32_t)lane + (uint32_t)k) << MMG_SUBTILE_SHIFT;
            for (uint32_t j = 0; j < ex[k].y; j++) {
                const uint32_t w = __ldcg(X.ev + ex[k].x + j);
                sm[at++] = (base + MMG_EV_OFF(w)) | (MMG_EV_JUMP(w) << 19) | ((w & MMG_EV_MATCH) ? (1u << 27) : 0u);
            }
        }
        __syncwarp();
        // replay, every lane the same: the chains of the block's alignment classes start at its first element
        const uint32_t J0 = P.J0;
        const uint32_t magic = 0xFFFFFFFFu / J0 + 1u;     // floor(2^32 / J0) + 1 (2^32 / J0 for powers of two): exact quotients below 2^19
        uint32_t xc[2] = {0u, 0u};
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t e = sm[i];
            const uint32_t bo = e & 0x7FFFFu;
            const uint32_t c = (W == 2) ? (bo & 1u) : 0u;
            const uint32_t q = bo / W, x = xc[c];
            if (x <= q) {
                const uint32_t d = q - x;
                if (J0 == 1u || d - __umulhi(d, magic) * J0 == 0u) {
                    xc[c] = q + ((e >> 19) & 0xFFu);
                    if (e & (1u << 27)) {
                        __syncwarp();                     // every lane has read entry i
                        if (lane == 0) sm[m] = bo;        // m <= i: never ahead of the read position
                        m++;
                    }
                }
            }
        }
        __syncwarp();
        dense = m >= SPARSE_REC;
    }
    if (dense) {
        if (lane == 0) atomicOr(X.ticket + 2, 1u);        // not sparse after all: the host falls back to k_resolve
        m = 0;
    }
    uint32_t *rec = X.brec + (size_t)bi * SPARSE_REC;
    if (lane == 0) { rec[0] = m; X.bcount[bi] = m; }
    if (lane < (int)m) rec[1 + lane] = sm[lane];
    __syncwarp();
}

// Emission of one resolved block: its position in the output is the sum of the counts of the blocks before it (a warp
// reads them coalesced: a few KB from L2), then one lane per match.
template <int W, bool BE>
__device__ __forceinline__ void sparse_emit_block(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint32_t b, int lane) {
    const uint32_t cnt = __ldcg(X.bcount + b);
    if (cnt == 0) return;
    uint64_t before = 0;
    for (uint32_t j = lane; j < b; j += 32) before += __ldcg(X.bcount + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(FULL, before, o);
    if ((uint32_t)lane < cnt && before + lane < X.capacity) {
        const uint32_t o0 = (uint32_t)P.first_lit * W;
        const bool has1 = P.opp_idx >= 0;
        const uint32_t o1 = has1 ? (uint32_t)P.opp_idx * W : 0u;
        const uint64_t sb = (uint64_t)b * G.B + __ldcg(X.brec + (size_t)b * SPARSE_REC + 1 + lane);
        uint32_t v = ld_elem<W, BE>(G.data + sb + o0);
        if (has1) v |= ld_elem<W, BE>(G.data + sb + o1) << 16;
        X.out_off[before + lane] = (G.base_offset + sb) >> G.report_shift;
        X.out_val[before + lane] = v;
    }
}

// The last CTA of the grid (its first warp): total, status words for the host, zero state of the workspace.
__device__ __noinline__ void sparse_last(const MmgGeom &G, const MmgScratch &X, int lane) {
    uint64_t total = 0;
    for (uint32_t j = lane; j < G.nblocks; j += 32) total += __ldcg(X.bcount + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
    if (lane < 2) X.host_status[lane] = reinterpret_cast<volatile uint64_t *>(X.status)[lane];
    if (lane == 2) X.host_status[2] = total;
    if (lane == 3) X.host_status[3] = 0;
    if (lane == 4) X.host_status[4] = reinterpret_cast<volatile uint32_t *>(X.ticket)[2];
    // (no system fence: the status slot is read by the host after the kernel has completed)
    __syncwarp();
    if (lane < 4) X.status[lane] = 0;
    if (lane < 8) X.ticket[lane] = 0;
}

// CTA-level arrival at a grid-wide barrier: the grid is persistent and launched cooperatively, so every CTA is resident.
// ONE thread per CTA polls, with a back-off -- thousands of pollers on one word saturate its L2 slice and stall the bulk
// copies of the CTAs that still filter.
__device__ __forceinline__ void grid_barrier(uint32_t *counter) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1u);
        volatile uint32_t *arrived = counter;
        uint32_t ns = 100;
        while (*arrived < gridDim.x) { __nanosleep(ns); ns = min(ns * 2u, 400u); }
    }
    __syncthreads();
    __threadfence();
}

// End of a filter CTA in a fused scan: barrier (every chunk is filtered) -- the warps share out the engine blocks and
// resolve them -- barrier (every block's match count is known) -- the warps write their blocks' matches -- the last CTA
// to get here reports.  ticket[1], [3]: the barriers; [4]: CTAs that are done.
__shared__ uint32_t g_fuse_last;

template <int W, bool BE>
__device__ __forceinline__ void fused_tail(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint32_t *sm, int lane) {
    const uint32_t wpc = blockDim.x >> 5;
    const uint32_t nwarps = gridDim.x * wpc;
    const uint32_t me = blockIdx.x * wpc + (threadIdx.x >> 5);
    grid_barrier(X.ticket + 1);
    const bool ok = !events_overflowed(X);                // (otherwise the lists are incomplete and the host re-runs the scan)
    for (uint32_t b = me; b < G.nblocks; b += nwarps) {
        if (ok) sparse_block<W>(P, G, X, b, sm, lane);
        else if (lane == 0) X.bcount[b] = 0;
    }
    grid_barrier(X.ticket + 3);
    if (ok)
        for (uint32_t b = me; b < G.nblocks; b += nwarps) sparse_emit_block<W, BE>(P, G, X, b, lane);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) g_fuse_last = atomicAdd(X.ticket + 4, 1u) + 1u == gridDim.x ? 1u : 0u;
    __syncthreads();
    if (g_fuse_last && threadIdx.x < 32) {
        __threadfence();
        sparse_last(G, X, lane);
    }
}

template <int W, int LB, bool BE, int NK>
__global__ void __launch_bounds__(MMG_FILTER_WARPS * 32, MMG_FILTER_MIN_CTAS)
k_filter(const __grid_constant__ MmgProgram P, const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * MMG_FILTER_WARPS + wib;
    const uint32_t reg_lo = warp * X.ev_per_warp;
    // window start = first byte of the current element of comparison 0, minus sigma
    const int64_t sigma = (LB == 0) ? 0 : (int64_t)P.chk[0].i * W;

    uint8_t *ring = smem_raw + (size_t)wib * MMG_WARP_SMEM;
    uint16_t *queue = reinterpret_cast<uint16_t *>(ring + MMG_NSTAGES * MMG_STAGE_STRIDE);      // positions relative to the chunk (< 2^16)
    const uint32_t ring_a = smem_u32(ring);
    const uint32_t bar_a = smem_u32(queue + MMG_QUEUE_CAP);
    if (lane == 0) {
        for (int i = 0; i < MMG_NSTAGES; i++) mbar_init(bar_a + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    load_sprog(P);      // includes the only __syncthreads() of the kernel

    WarpState st;
    st.cursor = reg_lo;
    uint32_t slot = 0, parity = 0;    // ring slot / mbarrier phase of the next stage to consume
    uint32_t static_round = 0;
    const uint64_t l2pol = l2_stream_policy(G.l2_hint);

    for (;;) {
        // dynamic chunk scheduling: warps draw chunks from a global counter (balances the tail).  Small inputs, where
        // every warp gets the same few chunks anyway, take them round robin: the draw's round trip (~1 us) would be a
        // tenth of the whole kernel
        uint32_t chunk = 0;
        if (G.static_chunks) {
            chunk = warp + static_round * gridDim.x * MMG_FILTER_WARPS;
            static_round++;
        } else {
            if (lane == 0) chunk = (uint32_t)atomicAdd((unsigned long long *)&X.status[3], 1ull);
            chunk = __shfl_sync(FULL, chunk, 0);
        }
        if (chunk >= G.nchunks) break;

        const uint32_t t0 = chunk * G.chunk_subs;
        const uint32_t t1 = min(t0 + G.chunk_subs, G.nsub);
        const uint32_t bi = t0 / G.spb;
        ChunkCtx C;
        C.blk_off = (int64_t)((uint64_t)bi * G.B);
        {
            // element count of each alignment's view of the block (src/core/search_engine.cpp:136-141)
            const uint64_t blk_size = min(G.B + (uint64_t)G.ov, G.S - (uint64_t)C.blk_off);
            for (uint32_t pad = 0; pad < 2; pad++) {
                uint64_t count = blk_size / W;
                if (pad + count * W > blk_size) count -= 1;
                C.max_rel[pad] = (pad < G.npads && count >= (uint64_t)P.L) ? (int64_t)(pad + (count - (uint64_t)P.L) * W) : -1;
            }
        }
        C.data = G.data;
        C.ev = X.ev;
        C.npads = G.npads;
        C.s_lo = (int64_t)t0 << MMG_SUBTILE_SHIFT;
        C.s_hi = (int64_t)t1 << MMG_SUBTILE_SHIFT;
        C.reg_hi = reg_lo + X.ev_per_warp;
        st.open_t = t0;
        st.open_start = st.cursor;
        st.t0 = t0; st.ex = 0u; st.ey = 0u;

        // current-element positions p run over [p0, p_end); the window start is p - sigma (- 1 for
        // the odd 16-bit class), so the chunk needs sigma + 1 extra bytes past its own window starts
        const uint64_t p0 = (uint64_t)C.s_lo;
        const uint64_t s16 = (G.S + 15) & ~(uint64_t)15;
        const uint64_t p_end = min((uint64_t)C.s_hi + (uint64_t)sigma + 1, s16);
        const uint32_t len = p_end > p0 ? (uint32_t)(p_end - p0) : 0u;         // bytes of p-space to process
        const uint32_t nst = (len + MMG_STAGE_BYTES - 1) / MMG_STAGE_BYTES;
        // bulk copies stop at the aligned end of the slice and never go past what the chunk needs
        const uint64_t aligned_end = G.S & ~(uint64_t)15;
        const uint64_t want_end = min((p_end + 15) & ~(uint64_t)15, aligned_end);
        const uint32_t copy_end_rel = want_end > p0 ? (uint32_t)(want_end - p0) : 0u;
        // unaligned tail of the slice inside this chunk?  (rel position of the first tail byte)
        const bool has_tail = aligned_end != G.S && aligned_end >= p0 && aligned_end < p0 + len;
        const uint32_t tail_rel = has_tail ? (uint32_t)(aligned_end - p0) : 0xFFFFFFFFu;
        const uint8_t *chunk_base = G.data + p0;
        const bool at_start = p0 == 0;
        C.q_base = (int64_t)p0 - sigma - (W == 2 ? 1 : 0);
        __syncwarp();                       // (eval_batch calls of the previous chunk are done with the old copy)
        if (lane == 0) g_ctx[wib] = C;
        __syncwarp();
        uint32_t qn = 0;
        if (lane == 0) {
            uint32_t sl = slot;
            for (uint32_t k = 0; k < nst && k < MMG_NSTAGES; k++) {
                issue_stage(chunk_base, at_start, k * MMG_STAGE_BYTES, copy_end_rel, ring_a + sl * MMG_STAGE_STRIDE, bar_a + 8 * sl, l2pol);
                sl = sl + 1 == MMG_NSTAGES ? 0 : sl + 1;
            }
        }
        for (uint32_t k = 0, rel_stage = 0; k < nst; k++, rel_stage += MMG_STAGE_BYTES) {
            mbar_wait(bar_a + 8 * slot, parity);
            if (tail_rel - rel_stage < MMG_STAGE_BYTES) {      // patch the last (S mod 16) bytes of the slice
                if (lane < 16 && aligned_end + lane < G.S)
                    ring[slot * MMG_STAGE_STRIDE + 16 + (tail_rel - rel_stage) + lane] = G.data[aligned_end + lane];
                __syncwarp();
            }
            const uint32_t rows = min(MMG_STAGE_BYTES / MMG_ROW, (len - rel_stage + MMG_ROW - 1) / MMG_ROW);
            uint32_t sa = ring_a + slot * MMG_STAGE_STRIDE + (uint32_t)lane * 16u;
#pragma unroll 1
            for (uint32_t r = 0; r < rows; r++, sa += MMG_ROW) {
                const uint4 prv = lds128(sa);
                const uint4 own = lds128(sa + 16);
                uint32_t x[8];
                x[0] = prv.x; x[1] = prv.y; x[2] = prv.z; x[3] = prv.w;
                x[4] = own.x; x[5] = own.y; x[6] = own.z; x[7] = own.w;
                uint32_t f[8];
                bool any;
                if (LB == 0) any = true;
                else if (W == 2) any = prefilter16<LB, BE, NK>(P, x);
                if (__any_sync(FULL, any)) {
                    // exact per-position flags (16-bit: only now), then ordered enqueue of the candidates
                    if (LB != 0 && W == 2) any = any && filter_lane<2, LB, BE, 0>(P, x, f);
                    const uint32_t cm = candidate_mask<W, LB>(f, any);
                    const uint32_t rel = rel_stage + r * MMG_ROW + (uint32_t)lane * 16u;   // candidate bit 0 of this lane
                    // lanes 0-15, then lanes 16-31 (ascending positions; at most 256 new queue entries per pass)
#pragma unroll 1
                    for (uint32_t hp = 0; hp < 2; hp++) {
                        const uint32_t cmh = ((uint32_t)lane >> 4) == hp ? cm : 0u;
                        if (!__any_sync(FULL, cmh != 0u)) continue;
                        // exclusive prefix of the per-lane counts (<= 16) from bit-sliced ballots
                        const uint32_t cnt = __popc(cmh);
                        const uint32_t lt = (1u << lane) - 1u;
                        const uint32_t b0 = __ballot_sync(FULL, cnt & 1u), b1 = __ballot_sync(FULL, cnt & 2u);
                        uint32_t pre = __popc(b0 & lt) + 2u * __popc(b1 & lt);
                        uint32_t total = __popc(b0) + 2u * __popc(b1);
                        if (__any_sync(FULL, cnt >= 4u)) {
                            const uint32_t b2 = __ballot_sync(FULL, cnt & 4u), b3 = __ballot_sync(FULL, cnt & 8u),
                                           b4 = __ballot_sync(FULL, cnt & 16u);
                            pre += 4u * __popc(b2 & lt) + 8u * __popc(b3 & lt) + 16u * __popc(b4 & lt);
                            total += 4u * __popc(b2) + 8u * __popc(b3) + 16u * __popc(b4);
                        }
                        uint32_t at = qn + pre;
                        uint32_t m = cmh;
                        while (m) {
                            queue[at++] = (uint16_t)(rel + (uint32_t)(__ffs(m) - 1));
                            m &= m - 1;
                        }
                        qn += total;
                        __syncwarp();
                        uint32_t qh = 0;
                        while (qn - qh >= 32) {
                            st = eval_batch<W, BE>(X, st, g_ctx[wib], queue[qh + lane], true, lane);
                            qh += 32;
                        }
                        if (qh) {          // move the < 32 left-overs to the front
                            const uint32_t left = qn - qh;
                            const uint32_t v = lane < left ? queue[qh + lane] : 0u;
                            __syncwarp();
                            if (lane < left) queue[lane] = (uint16_t)v;
                            qn = left;
                            __syncwarp();
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && k + MMG_NSTAGES < nst)
                issue_stage(chunk_base, at_start, rel_stage + MMG_NSTAGES * MMG_STAGE_BYTES, copy_end_rel,
                            ring_a + slot * MMG_STAGE_STRIDE, bar_a + 8 * slot, l2pol);
            if (++slot == MMG_NSTAGES) { slot = 0; parity ^= 1u; }
        }
        if (qn) st = eval_batch<W, BE>(X, st, g_ctx[wib], lane < qn ? queue[lane] : 0u, lane < qn, lane);
        __syncwarp();
        st = close_until(X, st, t1, lane);
        flush_extents(X, st, t1, lane);
    }
    if (lane == 0) {
        atomicMax((unsigned long long *)&X.status[0], (unsigned long long)(st.cursor - reg_lo));
        atomicAdd((unsigned long long *)&X.status[1], (unsigned long long)(st.cursor - reg_lo));
    }
    if (X.fuse) fused_tail<W, BE>(P, G, X, reinterpret_cast<uint32_t *>(ring), lane);
}


// ------------------------------------------------------------------------------------------
// K1 (8-bit): same per-warp TMA ring as k_filter, but every lane owns 32 contiguous bytes of a 1 KiB row, the
// candidates of a row are compacted with one warp scan and each becomes an event word in place (no queue):
//   * comparison 0 fails (the usual case): the advance comes from a 511-entry table in shared memory;
//   * it passes: comparison 1 is evaluated from the row in shared memory, deeper ones (rare) by a helper.
// Flagged windows that turn out to advance by J0 without a match are written as "null" events (a no-op in the
// replay), so the event offsets can be assigned before the evaluation.
// ------------------------------------------------------------------------------------------

#define MMG_ROW8 1024u
// the 8-bit kernel runs two 4 KiB stages per warp (same bytes in flight as three 2 KiB stages, half the per-stage
// bookkeeping per row)
#define MMG_STAGE8 4096u
#define MMG_NSTAGES8 2
#define MMG_STRIDE8 (MMG_STAGE8 + 16u)
#define MMG_WARP_SMEM8 ((MMG_NSTAGES8 * MMG_STRIDE8 + MMG_NSTAGES8 * 8u + 15u) & ~15u)     // ring + mbarriers, no queue
#ifndef MMG_FILTER8_MIN_CTAS
#define MMG_FILTER8_MIN_CTAS 3
#endif

template <int LB, int NK>
__global__ void __launch_bounds__(MMG_FILTER_WARPS * 32, MMG_FILTER8_MIN_CTAS)
k_filter8(const __grid_constant__ MmgProgram P, const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * MMG_FILTER_WARPS + wib;
    const uint32_t reg_lo = warp * X.ev_per_warp, reg_hi = reg_lo + X.ev_per_warp;

    uint8_t *ring = smem_raw + (size_t)wib * MMG_WARP_SMEM8;
    const uint32_t ring_a = smem_u32(ring);
    const uint32_t bar_a = ring_a + MMG_NSTAGES8 * MMG_STRIDE8;
    if (lane == 0) {
        for (int i = 0; i < MMG_NSTAGES8; i++) mbar_init(bar_a + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    load_sprog(P);      // includes the only __syncthreads() of the kernel

    // constants of the exact evaluation (volatile shared loads: see SProg)
    const volatile SProg &VP = g_sprog;
    const uint32_t sg = VP.sg;                            // window start = position of comparison 0's current element - sg
    const int ev_ed0 = VP.ed0, ev_ed1 = VP.ed1, ev_L = VP.L;
    const int ev_nc = P.ncheck;
    const uint32_t ev_pmask = VP.pmask, ev_o1c = VP.o1c, ev_o1p = VP.o1p, ev_match = VP.matchword;
    const uint32_t sprog_a = smem_u32(&g_sprog);
    uint32_t tab0_a = sprog_a + (uint32_t)offsetof(SProg, tab0) + 255u, tab1_a = sprog_a + (uint32_t)offsetof(SProg, tab1) + 255u;
    // keep the table addresses in registers: left alone, ptxas re-derives the shared window base for every candidate
    asm volatile("" : "+r"(tab0_a), "+r"(tab1_a));
    const uint32_t lt = (1u << lane) - 1u;
    const bool d2ok = P.d2ok != 0;

    WarpState st;
    st.cursor = reg_lo;
    // stages the depth-2 refinement stays on before the candidate density is probed again; with a single key every
    // candidate is a "comparison 0 passes" window, so the refinement is always worth its cost
    const bool always_dense = d2ok && (NK == 1 || P.nkeys == 1);
    uint32_t dense_left = always_dense ? 0xFFFFFFFFu : 0u;
    uint32_t slot = 0, parity = 0;    // ring slot / mbarrier phase of the next stage to consume
    uint32_t static_round = 0;
    const uint64_t l2pol = l2_stream_policy(G.l2_hint);

    for (;;) {
        uint32_t chunk = 0;
        if (G.static_chunks) {              // small inputs: round robin instead of a draw (see k_filter)
            chunk = warp + static_round * gridDim.x * MMG_FILTER_WARPS;
            static_round++;
        } else {
            if (lane == 0) chunk = (uint32_t)atomicAdd((unsigned long long *)&X.status[3], 1ull);
            chunk = __shfl_sync(FULL, chunk, 0);
        }
        if (chunk >= G.nchunks) break;

        const uint32_t t0 = chunk * G.chunk_subs;
        const uint32_t t1 = min(t0 + G.chunk_subs, G.nsub);
        const uint64_t blk_off = (uint64_t)(t0 / G.spb) * G.B;
        const uint64_t blk_size = min(G.B + (uint64_t)G.ov, G.S - blk_off);      // src/core/search_engine.cpp:227-230
        const uint64_t p0 = (uint64_t)t0 << MMG_SUBTILE_SHIFT;                    // first window start of the chunk
        const uint32_t chunk_bytes = (t1 - t0) << MMG_SUBTILE_SHIFT;
        st.open_t = t0;
        st.open_start = st.cursor;
        st.t0 = t0; st.ex = 0u; st.ey = 0u;

        // a candidate at chunk-relative position rel (current element of comparison 0) is a valid window iff
        // sg <= rel < v_hi: the window starts inside the chunk and is complete inside the block's view
        uint32_t v_hi = 0;
        if (blk_size >= (uint64_t)ev_L) {
            const int64_t lim = min((int64_t)chunk_bytes, (int64_t)(blk_off + blk_size - (uint64_t)ev_L + 1) - (int64_t)p0);
            if (lim > 0) v_hi = (uint32_t)lim + sg;
        }
        const uint64_t s16 = (G.S + 15) & ~(uint64_t)15;
        const uint64_t p_end = min(p0 + chunk_bytes + sg + 1, s16);
        const uint32_t len = p_end > p0 ? (uint32_t)(p_end - p0) : 0u;            // bytes of position space to process
        const uint32_t nst = (len + MMG_STAGE8 - 1) / MMG_STAGE8;
        const uint64_t aligned_end = G.S & ~(uint64_t)15;
        const uint64_t want_end = min((p_end + 15) & ~(uint64_t)15, aligned_end);
        const uint32_t copy_end_rel = want_end > p0 ? (uint32_t)(want_end - p0) : 0u;
        const bool has_tail = aligned_end != G.S && aligned_end >= p0 && aligned_end < p0 + len;
        const uint32_t tail_rel = has_tail ? (uint32_t)(aligned_end - p0) : 0xFFFFFFFFu;
        const uint8_t *chunk_base = G.data + p0;
        const bool at_start = p0 == 0;

        if (lane == 0) {
            uint32_t sl = slot;
            for (uint32_t k = 0; k < nst && k < MMG_NSTAGES8; k++) {
                issue_stage<MMG_STAGE8>(chunk_base, at_start, k * MMG_STAGE8, copy_end_rel, ring_a + sl * MMG_STRIDE8, bar_a + 8 * sl, l2pol);
                sl = sl + 1 == MMG_NSTAGES8 ? 0 : sl + 1;
            }
        }
        for (uint32_t k = 0, rel_stage = 0; k < nst; k++, rel_stage += MMG_STAGE8) {
            mbar_wait(bar_a + 8 * slot, parity);
            if (tail_rel - rel_stage < MMG_STAGE8) {      // patch the last (S mod 16) bytes of the slice
                if (lane < 16 && aligned_end + lane < G.S)
                    ring[slot * MMG_STRIDE8 + 16 + (tail_rel - rel_stage) + lane] = G.data[aligned_end + lane];
                __syncwarp();
            }
            // bytes of this stage that the bulk copy filled (exact evaluation from shared memory stays inside them)
            const uint32_t stage_fill = copy_end_rel > rel_stage ? min(MMG_STAGE8, copy_end_rel - rel_stage) : 0u;
            const uint32_t stage_a = ring_a + slot * MMG_STRIDE8;
            // a candidate whose current element sits at shared address ca has its whole window in this stage's buffer
            // (halo included) iff  win_lo <= ca <= win_lo + win_span
            const uint32_t win_lo = stage_a + sg;
            const int win_span_i = 16 + (int)stage_fill - ev_L;
            const uint32_t win_span = win_span_i > 0 ? (uint32_t)win_span_i : 0u;
            const bool win_ok = win_span_i >= 0;
            const bool stage_edge = rel_stage < sg || rel_stage + MMG_STAGE8 > v_hi;
            const bool dense = dense_left != 0;
            const uint32_t rows = min(MMG_STAGE8 / MMG_ROW8, (len - rel_stage + MMG_ROW8 - 1) / MMG_ROW8);
            uint32_t stage_cands = 0;
            uint32_t sa = stage_a + (uint32_t)lane * 32u;                               // the 16 bytes before the lane's own 32
            uint32_t rowrel = rel_stage;
#pragma unroll 1
            for (uint32_t r = 0; r < rows; r++, sa += MMG_ROW8, rowrel += MMG_ROW8) {
                const uint32_t lanerel = rowrel + (uint32_t)lane * 32u;
                uint32_t x[12];
                {
                    const uint4 a = lds128(sa), b = lds128(sa + 16), c = lds128(sa + 32);
                    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                    x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w;
                }
                uint32_t cm;
                if (dense) cm = filter8<LB, NK, true>(P, x) | (filter8<LB, NK, true>(P, x + 4) << 16);
                else cm = filter8<LB, NK, false>(P, x) | (filter8<LB, NK, false>(P, x + 4) << 16);
                if (stage_edge) {                                   // clip to the valid windows
                    const int lo = min(max((int)sg - (int)lanerel, 0), 32), hi = min(max((int)v_hi - (int)lanerel, 0), 32);
                    const uint32_t below_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
                    const uint32_t below_lo = lo >= 32 ? 0xFFFFFFFFu : ((1u << lo) - 1u);
                    cm &= below_hi & ~below_lo;
                }
                const bool boundary = (rowrel & (MMG_SUBTILE - 1)) == 0 && rowrel != 0;
                if (!__any_sync(FULL, cm != 0u)) {
                    if (boundary) st = close_at(X, st, t0 + (rowrel >> MMG_SUBTILE_SHIFT) - 1u, st.cursor, lane);
                    continue;
                }
                // exclusive prefix of the per-lane candidate counts
                const uint32_t cnt = __popc(cm);
                uint32_t inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                    asm volatile("{ .reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }"
                                 : "+r"(inc) : "r"(o));
                const uint32_t total = __shfl_sync(FULL, inc, 31);
                if (boundary) {     // candidates below rowrel + sg still belong to the previous sub-tile
                    const int bl = min(max((int)(rowrel + sg) - (int)lanerel, 0), 32);
                    const uint32_t below = bl >= 32 ? 0xFFFFFFFFu : ((1u << bl) - 1u);
                    const uint32_t nbefore = __reduce_add_sync(FULL, __popc(cm & below));
                    st = close_at(X, st, t0 + (rowrel >> MMG_SUBTILE_SHIFT) - 1u, st.cursor + nbefore, lane);
                }
                if (st.cursor + total <= reg_hi) {                 // otherwise: counted only, the host re-runs with more room
                    uint32_t *out = X.ev + (st.cursor + inc - cnt);
                    const uint32_t ws0 = lanerel - sg;              // window start of the lane's bit 0, chunk relative
                    while (cm) {
                        const uint32_t b = (uint32_t)__ffs(cm) - 1u;
                        cm &= cm - 1u;
                        const uint32_t ca = sa + 16u + b;                       // current element of comparison 0
                        const int dd = (int)lds8(ca) - (int)lds8(ca - LB);
                        const uint32_t ws = ws0 + b;
                        uint32_t res;
                        if ((((uint32_t)(dd - ev_ed0)) & ev_pmask) != 0u) {
                            res = lds8(tab0_a + dd);
                        } else if (ev_nc == 1) {
                            res = ev_match;
                        } else {
                            if (win_ok && ca - win_lo <= win_span) {          // the window usually lies in this stage's buffer
                                const int d1 = (int)lds8(ca - ev_o1c) - (int)lds8(ca - ev_o1p);
                                if ((((uint32_t)(d1 - ev_ed1)) & ev_pmask) != 0u) res = lds8(tab1_a + d1);
                                else if (ev_nc == 2) res = ev_match;
                                else res = eval_window_lds8(ca - sg, sprog_a, 2, ev_pmask);
                            } else {
                                res = eval_window<1, false>(P, chunk_base + ws);
                            }
                        }
                        *out++ = (ws & (MMG_SUBTILE - 1)) | (res << 16);       // 0x100 << 16 == MMG_EV_MATCH
                    }
                }
                st.cursor += total;
                stage_cands += total;
            }
            __syncwarp();
            if (lane == 0 && k + MMG_NSTAGES8 < nst)
                issue_stage<MMG_STAGE8>(chunk_base, at_start, rel_stage + MMG_NSTAGES8 * MMG_STAGE8, copy_end_rel,
                            ring_a + slot * MMG_STRIDE8, bar_a + 8 * slot, l2pol);
            if (++slot == MMG_NSTAGES8) { slot = 0; parity ^= 1u; }
            // candidate-dense data (low entropy): run the next 15 stages with the depth-2 refinement, then probe again
            if (always_dense) { }
            else if (dense_left) dense_left--;
            else if (d2ok && stage_cands >= MMG_STAGE8 * 3u / 64u) dense_left = 15;
        }
        // the row at chunk-relative 4096 * (t1 - t0) closed the last sub-tile unless the data ended before it
        if (st.open_t < t1) st = close_at(X, st, st.open_t, st.cursor, lane);
        flush_extents(X, st, t1, lane);
    }
    if (lane == 0) {
        atomicMax((unsigned long long *)&X.status[0], (unsigned long long)(st.cursor - reg_lo));
        atomicAdd((unsigned long long *)&X.status[1], (unsigned long long)(st.cursor - reg_lo));
    }
    if (X.fuse) fused_tail<1, false>(P, G, X, reinterpret_cast<uint32_t *>(ring), lane);
}


// ------------------------------------------------------------------------------------------
// K2: resolve -- everything after the filter in ONE kernel, one warp per engine block
//   (a) lane per sub-tile with events: map entry phase -> exit phase for each alignment class
//   (b) the maps are composed along the block's chains (sequentially over the few sub-tiles that
//       have events, closed form for the event-free stretches in between)
//   (c) lane per sub-tile: replay the TRUE chains through the events, mark visited matches
//   (d) decoupled look-back over blocks (ticket order) gives the block's base in the output
//   (e) ordered emission of (file offset, table base values)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t lattice_advance(uint32_t x, uint32_t n, uint32_t J0) {
    // first chain position >= n when the chain sits at x and advances by J0; returned relative to n
    if (x >= n) return x - n;
    if (J0 == 1) return 0;
    const uint32_t r = (n - x) % J0;
    return r ? J0 - r : 0;
}

#define LB_AGG (1ull << 62)
#define LB_INCL (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)

template <int W, bool BE>
__device__ __forceinline__ void emit_subtile(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint32_t t,
                                             uint64_t at, uint64_t *__restrict__ out_off, uint32_t *__restrict__ out_val) {
    const uint2 ext = X.ext[t];
    const uint32_t n = ext.y;
    const uint32_t *__restrict__ ev = X.ev + ext.x;
    const uint8_t *__restrict__ data = G.data;
    const uint32_t o0 = (uint32_t)P.first_lit * W;
    const bool has1 = P.opp_idx >= 0;
    const uint32_t o1 = has1 ? (uint32_t)P.opp_idx * W : 0u;
    const uint64_t tbase = (uint64_t)t << MMG_SUBTILE_SHIFT;
    // eight events at a time: their words, then the element loads of the visited ones (independent of each other: one
    // memory round trip per batch instead of one per match), then the stores
    for (uint32_t i0 = 0; i0 < n; i0 += 8) {
        uint32_t w[8], v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = i0 + k < n ? ev[i0 + k] : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = 0;
            if (w[k] & MMG_EV_VISITED) {
                const uint64_t s = tbase + MMG_EV_OFF(w[k]);
                v[k] = ld_elem<W, BE>(data + s + o0);
                if (has1) v[k] |= ld_elem<W, BE>(data + s + o1) << 16;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (w[k] & MMG_EV_VISITED) {
                out_off[at] = (G.base_offset + tbase + MMG_EV_OFF(w[k])) >> G.report_shift;
                out_val[at] = v[k];
                at++;
            }
    }
}

#ifdef MMG_RESOLVE_PROF
__device__ unsigned long long g_phase_ns[8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PHASE_MARK(k) do { __syncthreads(); if (threadIdx.x == 0) { unsigned long long now_ = gtimer(); atomicAdd(&g_phase_ns[k], now_ - t_prev_); t_prev_ = now_; } } while (0)
#define PHASE_INIT unsigned long long t_prev_ = gtimer()
#else
#define PHASE_MARK(k) do { } while (0)
#define PHASE_INIT do { } while (0)
#endif
#define RESOLVE_THREADS 128
#define RESOLVE_FAST_J 16u
#define RESOLVE_BATCH 16
#define SE_STRIDE (RESOLVE_THREADS + 1)     // words per residue row of the fast maps: odd, so that the rows of one sub-tile lie in different banks

// SPLIT > 1 (small inputs, where a CTA has an SM to itself and step (a) is one thread applying ~80 events at the issue
// latency of a lone warp): SPLIT threads share the event list of a sub-tile.  Thread (sub-tile, part) applies its
// quarter of the events to a map of its own that starts as the identity on residues (the rightmost part: as the
// sub-tile's default exit map); the owner thread (part 0) then chains the SPLIT partial maps, exactly as step (b)
// chains the maps of whole sub-tiles.  The helper warps leave after that.
template <int W, bool BE, bool MAPS_ONLY, int SPLIT>
__global__ void __launch_bounds__(RESOLVE_THREADS * SPLIT)
k_resolve(const __grid_constant__ MmgProgram P, const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X,
          uint64_t *out_off, uint32_t *out_val, uint64_t capacity, uint32_t jp) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    // shared: maps [128][npads][jp] bytes (general) or words [npads][16][128] (fast) | entry phases [128][2] | has flags [128][2]
    const uint32_t npads = G.npads;
    // Fast maps: when no advance exceeds J0 (simple / value-scan patterns) a chain that leaves an event lands on a
    // lattice residue and cannot meet an event of that residue inside the gap it jumps over, so ONE right-to-left
    // pass over the events yields, per residue r, the exit phase and the number of matches visited by a chain that
    // sits on residue r in front of the events seen so far:  E[r] <- E[(r + advance) mod J0] (+1 for a match).
    const bool fast = P.J0 == P.Jmax && P.Jmax <= RESOLVE_FAST_J;
    uint8_t *s_map = rs_smem;
    uint32_t *s_E = reinterpret_cast<uint32_t *>(rs_smem);      // word = exit phase | matches << 8, index (c * 16 + r) * SE_STRIDE + tid
    uint8_t *s_ph = s_map + (fast ? (size_t)SE_STRIDE * npads * RESOLVE_FAST_J * 4 * SPLIT : (size_t)RESOLVE_THREADS * npads * jp);
    uint8_t *s_has = s_ph + RESOLVE_THREADS * 2;
    __shared__ uint32_t s_bi, s_cnt[RESOLVE_THREADS / 32], s_phase[2];
    __shared__ uint64_t s_before;
    // fast maps, step (b): [warp][class * 16 + residue][k] = phase with which the chain that enters the warp's 32
    // sub-tiles on that residue enters sub-tile k; s_wmap = where it leaves the 32 sub-tiles
    __shared__ uint8_t s_traj[RESOLVE_THREADS / 32][32][33], s_wmap[RESOLVE_THREADS / 32][32];
    __shared__ uint8_t s_seen[SPLIT][RESOLVE_THREADS];
    // tid = sub-tile of the round this thread works for; part 0 owns it (and is all there is when SPLIT == 1)
    const int part = SPLIT == 1 ? 0 : (int)(threadIdx.x / RESOLVE_THREADS);
    const int tid = SPLIT == 1 ? (int)threadIdx.x : (int)(threadIdx.x % RESOLVE_THREADS), lane = tid & 31, wid = tid >> 5;
    const uint32_t part_words = SE_STRIDE * npads * RESOLVE_FAST_J;      // one partial map array (fast maps)

    // blocks are taken in ticket order, so every predecessor of a block is already running (look-back is safe)
    if (threadIdx.x == 0) s_bi = MAPS_ONLY ? blockIdx.x : atomicAdd(X.ticket, 1u);
    __syncthreads();
    const uint32_t bi = s_bi;
    PHASE_INIT;
    const bool bad = events_overflowed(X);
    // segment bi = segment (bi mod segs_per_block) of engine block (bi / segs_per_block)
    const uint32_t rb = bi / G.segs_per_block, si = bi - rb * G.segs_per_block;
    const uint32_t t_begin = rb * G.spb + si * RESOLVE_THREADS;
    const uint32_t t_end = G.segs_per_block == 1 ? min((rb + 1) * G.spb, G.nsub)
                                                 : min(min(t_begin + RESOLVE_THREADS, (rb + 1) * G.spb), G.nsub);
    // the chain restarts at every engine block (entry phase 0); a slice of a longer chain (mmg_chain_*) is entered with
    // the phase the slices before it leave behind, and segment k > 0 of a block with what k_segphase / k_chainphase found
    if (threadIdx.x < 2) s_phase[tid] = MAPS_ONLY ? 0u : G.chain ? X.segphase[bi * 2 + tid] : si == 0 ? 0u : X.segphase[bi * 2 + tid];
    __syncthreads();
    const uint32_t J0 = P.J0, Jmax = P.Jmax, NP = MMG_SUBTILE / W;
    const uint32_t magic = 65536u / J0 + 1u;      // q mod J0 = q - J0 * ((q * magic) >> 16), exact for q < 4096, J0 <= 16
    uint32_t total = 0;       // matches of this block (valid in every thread after the loop)

    for (uint32_t tb = t_begin; tb < t_end && !bad; tb += RESOLVE_THREADS) {
        const uint32_t t = tb + tid;
        const uint2 ext = t < t_end ? X.ext[t] : make_uint2(0u, 0u);
        const bool he = ext.y != 0;
        const uint32_t nvalid = min((uint32_t)RESOLVE_THREADS, t_end - tb);
        if (!__syncthreads_or(he)) {
            if (MAPS_ONLY) {            // no events in the whole segment: every entry phase just follows its lattice
                for (uint32_t i = tid; i < npads * Jmax; i += RESOLVE_THREADS)
                    X.segmap[((size_t)bi * 2 + i / Jmax) * jp + i % Jmax] = (uint8_t)lattice_advance(i % Jmax, nvalid * NP, J0);
                return;
            }
            if (threadIdx.x == 0)       // (one thread of the CTA: with SPLIT > 1 every part has a tid 0)
                for (uint32_t c = 0; c < npads; c++) s_phase[c] = lattice_advance(s_phase[c], nvalid * NP, J0);
            continue;
        }
        PHASE_MARK(0);
        // (a) maps of this thread's sub-tile, both alignment classes
        uint32_t n = 0;
        uint32_t *ev = nullptr;
        if (part == 0) { s_has[tid * 2] = 0; s_has[tid * 2 + 1] = 0; }
        uint32_t *const s_P = s_E + (size_t)part * part_words;       // this thread's (partial) map
        if (fast && t < t_end) {
            // no events: the lattice of residue r leaves the sub-tile at (r - NP) mod J0.  With the events shared out,
            // only the part that applies the rightmost events starts from that; the others start as the identity.
            const uint32_t base = NP % J0;
            const bool exit_map = part == SPLIT - 1 || !he;
            for (uint32_t c = 0; c < npads; c++)
                for (uint32_t r = 0; r < J0; r++)
                    s_P[(c * RESOLVE_FAST_J + r) * SE_STRIDE + tid] = !exit_map ? r : r >= base ? r - base : r + J0 - base;
        }
        if (he && fast) {
            n = ext.y;
            ev = X.ev + ext.x;
            const uint32_t ev_lo = SPLIT == 1 ? 0u : n * (uint32_t)part / SPLIT, ev_hi = SPLIT == 1 ? n : n * (uint32_t)(part + 1) / SPLIT;
            uint32_t seen = 0;
            auto step = [&](uint32_t w) {
                const uint32_t off = MMG_EV_OFF(w);
                const uint32_t c = (W == 2) ? (off & 1u) : 0u;
                const uint32_t q = off / W;
                const uint32_t r = q - J0 * ((q * magic) >> 16);
                uint32_t ry = r + MMG_EV_JUMP(w);       // advance <= J0
                if (ry >= J0) ry -= J0;
                s_P[(c * RESOLVE_FAST_J + r) * SE_STRIDE + tid] =
                    s_P[(c * RESOLVE_FAST_J + ry) * SE_STRIDE + tid] + ((w >> 16) & 0x100u);
                seen |= 1u << c;
            };
            // right to left, RESOLVE_BATCH events per batch; the next batch is requested before the current one is
            // applied, so the memory round trips of a long event list overlap with the dependent shared-memory updates
            // (batches of 16 or of 8 measure the same: on small inputs the step is bound by the issue latency of a lone
            // warp working through ~25 instructions per event, which is what SPLIT addresses, not by the round trips)
            {
                const uint32_t *evp = ev + ev_lo;                 // this part's events: [0, i) of evp
                uint32_t cur[RESOLVE_BATCH], nxt[RESOLVE_BATCH];
                uint32_t i = ev_hi - ev_lo;
#pragma unroll
                for (int k = 0; k < RESOLVE_BATCH; k++) cur[k] = (uint32_t)k < i ? evp[i - 1 - k] : 0u;
                while (i > 0) {
                    const uint32_t ni = i > RESOLVE_BATCH ? i - RESOLVE_BATCH : 0u;
#pragma unroll
                    for (int k = 0; k < RESOLVE_BATCH; k++) nxt[k] = (uint32_t)k < ni ? evp[ni - 1 - k] : 0u;
#pragma unroll
                    for (int k = 0; k < RESOLVE_BATCH; k++)
                        if ((uint32_t)k < i) step(cur[k]);
#pragma unroll
                    for (int k = 0; k < RESOLVE_BATCH; k++) cur[k] = nxt[k];
                    i = ni;
                }
            }
            if (SPLIT == 1) { s_has[tid * 2] = seen & 1u; s_has[tid * 2 + 1] = (seen >> 1) & 1u; }
            else s_seen[part][tid] = (uint8_t)seen;
        } else if (he && part == 0) {
            n = ext.y;
            ev = X.ev + ext.x;
            for (uint32_t c = 0; c < npads; c++) {
                uint32_t x[MMG_MAXL];
                for (uint32_t e = 0; e < Jmax; e++) x[e] = e;
                bool any = false;
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t w = ev[i], off = MMG_EV_OFF(w);
                    if (W == 2 && (off & 1u) != c) continue;
                    const uint32_t q = off / W, j = MMG_EV_JUMP(w);
                    any = true;
                    for (uint32_t e = 0; e < Jmax; e++) {
                        const uint32_t xe = x[e];
                        if (xe <= q && (J0 == 1 || (q - xe) % J0 == 0)) x[e] = q + j;
                    }
                }
                if (any) {
                    s_has[tid * 2 + c] = 1;
                    uint8_t *m = s_map + ((size_t)tid * npads + c) * jp;
                    for (uint32_t e = 0; e < Jmax; e++) m[e] = (uint8_t)lattice_advance(x[e], NP, J0);
                }
            }
        }
        __syncthreads();
        if (SPLIT > 1) {
            // the owner chains the partial maps of its sub-tile, left to right: residue -> residue -> ... -> exit phase,
            // matches added up (in place: entry r of the result depends on entry r of part 0 only)
            if (part == 0 && he && fast) {
                for (uint32_t c = 0; c < npads; c++)
                    for (uint32_t r = 0; r < J0; r++) {
                        uint32_t v = s_E[(c * RESOLVE_FAST_J + r) * SE_STRIDE + tid];
                        uint32_t m = v >> 8;
#pragma unroll
                        for (int p = 1; p < SPLIT; p++) {
                            v = s_E[(size_t)p * part_words + (c * RESOLVE_FAST_J + (v & 0xFFu)) * SE_STRIDE + tid];
                            m += v >> 8;
                        }
                        s_E[(c * RESOLVE_FAST_J + r) * SE_STRIDE + tid] = (v & 0xFFu) | (m << 8);
                    }
                uint32_t seen = 0;
#pragma unroll
                for (int p = 0; p < SPLIT; p++) seen |= s_seen[p][tid];
                s_has[tid * 2] = seen & 1u; s_has[tid * 2 + 1] = (seen >> 1) & 1u;
            }
            __syncthreads();
            if (part != 0) return;          // the helper warps are done (whole warps: the barriers below count the rest)
        }
        // (b, fast maps) Every sub-tile of the round has a map now, so the composition needs no bookkeeping, and it runs
        // in two levels instead of one warp walking all 128 sub-tiles (which was a third of the kernel's run time on a
        // dense 16 MiB input): lane (class, residue) of warp w follows its residue through sub-tiles 32w .. 32w+31 and
        // notes the phase in front of each one; the four warp maps are then chained from the block's entry phase, which
        // selects the trajectory that is real.
        uint32_t fin[2] = {0u, 0u};
        if (fast) {
            const uint32_t c = (uint32_t)lane >> 4, r = (uint32_t)lane & 15u;
            const bool act = c < npads && r < J0;
            const uint32_t w0 = (uint32_t)wid * 32u;
            const uint32_t kmax = nvalid > w0 ? min(32u, nvalid - w0) : 0u;
            uint32_t ph = r;
            for (uint32_t k = 0; k < kmax; k++) {
                s_traj[wid][lane][k] = (uint8_t)ph;
                if (act) ph = s_E[(c * RESOLVE_FAST_J + ph) * SE_STRIDE + w0 + k] & 0xFFu;
            }
            s_wmap[wid][lane] = (uint8_t)ph;
            __syncthreads();
            for (uint32_t cc = 0; cc < npads; cc++) {
                uint32_t e = MAPS_ONLY ? 0u : s_phase[cc];
                if (!MAPS_ONLY) {
                    for (int w = 0; w < wid; w++) e = s_wmap[w][cc * 16u + e];
                    if ((uint32_t)tid < nvalid) s_ph[tid * 2 + cc] = s_traj[wid][cc * 16u + e][lane];
                    for (int w = wid; w < RESOLVE_THREADS / 32; w++) e = s_wmap[w][cc * 16u + e];
                }
                fin[cc] = e;
            }
            if (MAPS_ONLY) {
                if (tid < 32 && act) {
                    uint32_t e = r;
                    for (int w = 0; w < RESOLVE_THREADS / 32; w++) e = s_wmap[w][c * 16u + e];
                    X.segmap[((size_t)bi * 2 + c) * jp + r] = (uint8_t)e;
                }
                return;
            }
        }
        if (MAPS_ONLY) {            // (general maps)
            // (b') the map of the whole segment: lane e follows entry phase e through the sub-tile maps (a segment is
            // one round, so this is all the kernel has to produce)
            if (wid < (int)npads) {
                const uint32_t c = wid;
                for (uint32_t e0 = 0; e0 < Jmax; e0 += 32) {
                    const uint32_t e = e0 + lane;
                    uint32_t ph = e < Jmax ? e : 0u, done = 0;
                    for (uint32_t g = 0; g < RESOLVE_THREADS && g < nvalid; g += 32) {
                        uint32_t m = __ballot_sync(FULL, g + lane < nvalid && s_has[(g + lane) * 2 + c]);
                        while (m) {
                            const uint32_t l = g + __ffs(m) - 1;
                            m &= m - 1;
                            if (l > done) ph = lattice_advance(ph, (l - done) * NP, J0);
                            ph = fast ? (s_E[(c * RESOLVE_FAST_J + ph) * SE_STRIDE + l] & 0xFFu) : s_map[((size_t)l * npads + c) * jp + ph];
                            done = l + 1;
                        }
                    }
                    if (nvalid > done) ph = lattice_advance(ph, (nvalid - done) * NP, J0);
                    if (e < Jmax) X.segmap[((size_t)bi * 2 + c) * jp + e] = (uint8_t)ph;
                }
            }
            return;
        }
        PHASE_MARK(1);
        // (b, general maps) phases: warp c composes the maps of class c over the sub-tiles of this round, in order
        if (!fast && wid < (int)npads) {
            const uint32_t c = wid;
            uint32_t ph = s_phase[c], done = 0;
            for (uint32_t g = 0; g < RESOLVE_THREADS && g < nvalid; g += 32) {
                uint32_t m = __ballot_sync(FULL, g + lane < nvalid && s_has[(g + lane) * 2 + c]);
                while (m) {
                    const uint32_t l = g + __ffs(m) - 1;
                    m &= m - 1;
                    if (l > done) ph = lattice_advance(ph, (l - done) * NP, J0);
                    if (lane == 0) s_ph[l * 2 + c] = (uint8_t)ph;
                    ph = fast ? (s_E[(c * RESOLVE_FAST_J + ph) * SE_STRIDE + l] & 0xFFu) : s_map[((size_t)l * npads + c) * jp + ph];
                    done = l + 1;
                }
            }
            if (nvalid > done) ph = lattice_advance(ph, (nvalid - done) * NP, J0);
            if (lane == 0) s_phase[c] = ph;
        }
        __syncthreads();
        if (fast && tid < 2) s_phase[tid] = fin[tid];       // (every thread has read the old value before the barrier)
        PHASE_MARK(2);
        // (c) replay the true chains through this thread's events
        uint32_t cnt = 0;
        if (G.complete) {
            // complete-match mode: every match event counts, whatever the chain does (a matching window is always an event)
            if (he) {
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t w = ev[i];
                    if (w & MMG_EV_MATCH) { ev[i] = w | MMG_EV_VISITED; cnt++; }
                }
                X.mcount[t] = cnt;
            }
        } else if (he && fast) {
            // the match count of the chain that really enters is already known; the replay only marks its matches
            uint32_t xc[2] = {0u, 0u}, xm[2] = {0u, 0u};
            for (uint32_t c = 0; c < npads; c++)
                if (s_has[tid * 2 + c]) {
                    xc[c] = xm[c] = s_ph[tid * 2 + c];      // entry phase < J0: position and residue coincide
                    cnt += s_E[(c * RESOLVE_FAST_J + xc[c]) * SE_STRIDE + tid] >> 8;
                }
            if (cnt) {
                auto visit = [&](uint32_t i, uint32_t w) {
                    const uint32_t off = MMG_EV_OFF(w);
                    const uint32_t c = (W == 2) ? (off & 1u) : 0u;
                    const uint32_t q = off / W;
                    const uint32_t r = q - J0 * ((q * magic) >> 16);
                    if (xc[c] <= q && r == xm[c]) {
                        if (w & MMG_EV_MATCH) ev[i] = w | MMG_EV_VISITED;
                        const uint32_t j = MMG_EV_JUMP(w);
                        xc[c] = q + j;
                        xm[c] = r + j >= J0 ? r + j - J0 : r + j;
                    }
                };
                uint32_t cur[8], nxt[8];
#pragma unroll
                for (int k = 0; k < 8; k++) cur[k] = (uint32_t)k < n ? ev[k] : 0u;
                for (uint32_t i = 0; i < n; i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; k++) nxt[k] = i + 8 + k < n ? ev[i + 8 + k] : 0u;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (i + k < n) visit(i + k, cur[k]);
#pragma unroll
                    for (int k = 0; k < 8; k++) cur[k] = nxt[k];
                }
            }
            X.mcount[t] = cnt;
        } else if (he) {
            uint32_t xc[2] = {s_has[tid * 2] ? s_ph[tid * 2] : 0u, s_has[tid * 2 + 1] ? s_ph[tid * 2 + 1] : 0u};
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t w = ev[i], off = MMG_EV_OFF(w);
                const uint32_t c = (W == 2) ? (off & 1u) : 0u;
                const uint32_t q = off / W, x = xc[c];
                if (x <= q && (J0 == 1 || (q - x) % J0 == 0)) {
                    if (w & MMG_EV_MATCH) { ev[i] = w | MMG_EV_VISITED; cnt++; }
                    xc[c] = q + MMG_EV_JUMP(w);
                }
            }
            X.mcount[t] = cnt;
        }
        const uint32_t wsum = __reduce_add_sync(FULL, cnt);
        if (lane == 0) s_cnt[wid] = wsum;
        __syncthreads();
        for (int i = 0; i < RESOLVE_THREADS / 32; i++) total += s_cnt[i];
        __syncthreads();
    }

    if (SPLIT > 1 && part != 0) return;
    if (MAPS_ONLY) return;       // (an overflowed event buffer skips the loop: nothing to do, the resolve kernel reports it)

    PHASE_MARK(3);
    // (d) base of this block in the output: decoupled look-back, 8 x 32 predecessors per step (the eight
    // window loads are independent, so a block far from the nearest inclusive prefix still needs few round trips)
    if (wid == 0) {
        volatile uint64_t *lb = X.lookback;
        uint64_t before = 0;
        if (bi > 0) {
            if (lane == 0) lb[bi] = LB_AGG | total;
            int64_t hi = (int64_t)bi - 1;             // nearest predecessor not yet accounted for
            bool finished = false;
            while (!finished) {
                uint64_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int64_t j = hi - 32 * k - lane;
                    v[k] = j >= 0 ? lb[j] : LB_INCL;                  // "block -1" has an inclusive prefix of 0
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t flag = (uint32_t)(v[k] >> 62);
                    const uint32_t incl = __ballot_sync(FULL, flag == 2);
                    const uint32_t stop = incl ? (uint32_t)__ffs(incl) - 1 : 32u;      // nearest inclusive prefix
                    const uint32_t need = stop == 32 ? FULL : ((2u << stop) - 1u);     // lanes 0..stop
                    if (__ballot_sync(FULL, flag == 0) & need) break;                  // not published yet: reload from hi
                    uint64_t part = (lane <= (int)stop) ? (v[k] & LB_MASK) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                    before += part;
                    hi -= 32;
                    if (stop < 32) { finished = true; break; }
                }
            }
        }
        if (lane == 0) {
            lb[bi] = LB_INCL | (before + total);
            s_before = before;
            if (bi == G.nseg - 1) X.status[2] = before + total;
        }
    }
    __syncthreads();

    PHASE_MARK(4);
    // (e) ordered emission
    if (total != 0) {
        uint64_t running = s_before;
        for (uint32_t tb = t_begin; tb < t_end; tb += RESOLVE_THREADS) {
            const uint32_t t = tb + tid;
            const bool he = t < t_end && X.ext[t].y != 0;
            const uint32_t cnt = he ? X.mcount[t] : 0u;
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            __syncthreads();
            if (lane == 31) s_cnt[wid] = incl;
            __syncthreads();
            uint32_t wbase = 0, round_total = 0;
            for (int i = 0; i < RESOLVE_THREADS / 32; i++) { if (i < wid) wbase += s_cnt[i]; round_total += s_cnt[i]; }
            const uint64_t at = running + wbase + (incl - cnt);
            if (he) X.mbase[t] = at;
            if (cnt && at + cnt <= capacity) emit_subtile<W, BE>(P, G, X, t, at, out_off, out_val);
            running += round_total;
        }
    }

    PHASE_MARK(5);
    // (f) the last CTA to finish hands the status words to the host (pinned, device-visible slot) and restores the
    // all-zero state of the workspace, so the next scan on this stream needs no memset and no status copy
    __shared__ uint32_t s_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(X.ticket + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (tid < 4) X.host_status[tid] = reinterpret_cast<volatile uint64_t *>(X.status)[tid];
        if (tid == 4) X.host_status[4] = 0;                       // (the fused sparse resolve reports "too dense" here)
        __threadfence_system();
        __syncthreads();
        if (tid < 4) X.status[tid] = 0;
        if (tid < 2) X.ticket[tid] = 0;
        for (uint32_t i = tid; i < G.nseg; i += RESOLVE_THREADS) X.lookback[i] = 0;
    }
}

// ------------------------------------------------------------------------------------------
// K2b: entry phase of every segment of a block that is cut into several segments: the chain enters segment 0 with
// phase 0 and segment k+1 with segmap[k][phase of k].  One warp per engine block; the maps of 32 segments at a
// time are staged in shared memory with coalesced loads, the dependent look-ups then run at shared-memory latency.
// ------------------------------------------------------------------------------------------

#define SEGPHASE_WARPS 4

__global__ void __launch_bounds__(SEGPHASE_WARPS * 32)
k_segphase(const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X, uint32_t jp) {
    extern __shared__ __align__(16) uint8_t sp_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t rb = blockIdx.x * SEGPHASE_WARPS + wid;
    if (rb >= G.nblocks || events_overflowed(X)) return;       // (overflow: no maps were written, the host re-runs the scan)
    const uint32_t row = 2u * jp;                           // bytes of one segment's maps (both classes)
    uint8_t *stage = sp_smem + (size_t)wid * 32u * row;
    const uint32_t first = rb * G.segs_per_block;
    const uint32_t nseg = min(G.segs_per_block, G.nseg - first);
    uint32_t ph = 0;                                        // lanes 0 and 1 follow classes 0 and 1
    for (uint32_t s0 = 0; s0 < nseg; s0 += 32) {
        const uint32_t ns = min(32u, nseg - s0);
        const uint4 *src = reinterpret_cast<const uint4 *>(X.segmap + (size_t)(first + s0) * row);
        for (uint32_t i = lane; i < ns * row / 16u; i += 32) reinterpret_cast<uint4 *>(stage)[i] = src[i];
        __syncwarp();
        if (lane < (int)G.npads) {
            for (uint32_t k = 0; k < ns; k++) {
                X.segphase[(size_t)(first + s0 + k) * 2 + lane] = (uint8_t)ph;
                ph = stage[k * row + lane * jp + ph];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K2c: ONE chain over a buffer that is cut into slices (mmg_chain_*, one slice per GPU).  A slice is scanned before
// its entry phase is known: the filter and the segment maps do not depend on it.  What the other slices need from
// this one is its MAP -- exit phase for every possible entry phase -- i.e. the composition of all its segment maps.
// 8 GiB are 16384 segments, so the composition runs in two levels:
//   k_rangemap   warp j composes the segments of range j (<= 128 ranges), lane = entry phase, 32 segments staged in
//                shared memory at a time (as in k_segphase)
//   k_slicemap   one CTA composes the range maps (thread = class x entry phase) into the pinned host buffer and
//                hands the filter's event counts to the host (no resolve kernel has run yet to do that)
// and, once the entry phase arrived (exchange between the ranks, then mmg_chain_finish):
//   k_chainphase warp j walks the range maps 0..j-1 from the slice's entry phase, then its own segments, writing the
//                entry phase of every segment; k_resolve follows as for any block cut into segments.
// ------------------------------------------------------------------------------------------

#define CHAIN_RANGES 128
#define CHAIN_WARPS 4

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
k_rangemap(const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X, uint32_t jp, uint32_t Jmax,
           uint32_t nsegs, uint32_t per_range, uint32_t nranges) {
    extern __shared__ __align__(16) uint8_t cr_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x * CHAIN_WARPS + wid;
    if (j >= nranges || events_overflowed(X)) return;          // (overflow: no maps were written, the host re-runs the scan)
    const uint32_t row = 2u * jp;
    uint8_t *stage = cr_smem + (size_t)wid * 32u * row;
    const uint32_t first = j * per_range;
    const uint32_t ns_all = min(per_range, nsegs - first);
    uint32_t ph[2][4];
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int g = 0; g < 4; g++) ph[c][g] = (uint32_t)(g * 32 + lane);
    for (uint32_t s0 = 0; s0 < ns_all; s0 += 32) {
        const uint32_t ns = min(32u, ns_all - s0);
        const uint4 *src = reinterpret_cast<const uint4 *>(X.segmap + (size_t)(first + s0) * row);
        for (uint32_t i = lane; i < ns * row / 16u; i += 32) reinterpret_cast<uint4 *>(stage)[i] = src[i];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int g = 0; g < 4; g++)
                if (c < (int)G.npads && (uint32_t)(g * 32 + lane) < Jmax)
                    for (uint32_t k = 0; k < ns; k++) ph[c][g] = stage[k * row + c * jp + ph[c][g]];
        __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int g = 0; g < 4; g++)
            if (c < (int)G.npads && (uint32_t)(g * 32 + lane) < Jmax)
                X.rangemap[((size_t)j * 2 + c) * jp + g * 32 + lane] = (uint8_t)ph[c][g];
}

__global__ void __launch_bounds__(256)
k_slicemap(const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X, uint32_t jp, uint32_t Jmax, uint32_t nranges) {
    extern __shared__ __align__(16) uint8_t cs_smem[];
    const uint32_t row = 2u * jp;
    for (uint32_t i = threadIdx.x; i < nranges * row / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(cs_smem)[i] = reinterpret_cast<const uint4 *>(X.rangemap)[i];
    __syncthreads();
    const uint32_t c = threadIdx.x >> 7, e = threadIdx.x & 127u;
    if (c < G.npads && e < Jmax && !events_overflowed(X)) {
        uint32_t ph = e;
        for (uint32_t j = 0; j < nranges; j++) ph = cs_smem[j * row + c * jp + ph];
        X.slicemap_host[c * jp + e] = (uint8_t)ph;
    }
    if (threadIdx.x == 0) {            // the overflow check of the host needs the filter's event counts now
        X.host_status[0] = reinterpret_cast<volatile uint64_t *>(X.status)[0];
        X.host_status[1] = reinterpret_cast<volatile uint64_t *>(X.status)[1];
        X.host_status[2] = 0;
        X.host_status[4] = 0;
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
k_chainphase(const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X, uint32_t jp, uint32_t nsegs,
             uint32_t per_range, uint32_t nranges) {
    extern __shared__ __align__(16) uint8_t cp_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t row = 2u * jp;
    uint8_t *ranges = cp_smem;                                         // [nranges][row]
    uint8_t *stage = cp_smem + (size_t)CHAIN_RANGES * row + (size_t)wid * 32u * row;
    const uint32_t j = blockIdx.x * CHAIN_WARPS + wid;
    const uint32_t jmax_cta = min(blockIdx.x * CHAIN_WARPS + CHAIN_WARPS - 1, nranges - 1);     // ranges this CTA looks back over
    for (uint32_t i = threadIdx.x; i < jmax_cta * row / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(ranges)[i] = reinterpret_cast<const uint4 *>(X.rangemap)[i];
    __syncthreads();
    if (j >= nranges || events_overflowed(X)) return;
    uint32_t ph = lane < 2 ? G.entry[lane] : 0u;                      // lanes 0 and 1 follow classes 0 and 1
    if (lane < (int)G.npads)
        for (uint32_t i = 0; i < j; i++) ph = ranges[i * row + lane * jp + ph];
    const uint32_t first = j * per_range;
    const uint32_t ns_all = min(per_range, nsegs - first);
    for (uint32_t s0 = 0; s0 < ns_all; s0 += 32) {
        const uint32_t ns = min(32u, ns_all - s0);
        const uint4 *src = reinterpret_cast<const uint4 *>(X.segmap + (size_t)(first + s0) * row);
        for (uint32_t i = lane; i < ns * row / 16u; i += 32) reinterpret_cast<uint4 *>(stage)[i] = src[i];
        __syncwarp();
        if (lane < (int)G.npads) {
            for (uint32_t k = 0; k < ns; k++) {
                X.segphase[(size_t)(first + s0 + k) * 2 + lane] = (uint8_t)ph;
                ph = stage[k * row + lane * jp + ph];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K4: exclusive prefix sum (u32 counts -> u64 bases); 1024 items per CTA
// ------------------------------------------------------------------------------------------

#define SCAN_ITEMS 1024

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *warp_sums, uint64_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    uint64_t base = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; i++) { if (i < wid) base += warp_sums[i]; tot += warp_sums[i]; }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(256) k_scan_sums(const uint32_t *in, uint32_t n, uint64_t *bsum) {
    __shared__ uint64_t ws[8];
    const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
    uint64_t v = 0;
    for (int i = 0; i < 4; i++) if (base + i < n) v += in[base + i];
    uint64_t tot;
    block_exclusive_scan(v, ws, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) k_scan_top(uint64_t *bsum, uint32_t nb, uint64_t *total_out) {
    __shared__ uint64_t ws[8];
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += 256) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = i < nb ? bsum[i] : 0;
        uint64_t tot;
        const uint64_t ex = block_exclusive_scan(v, ws, &tot);
        if (i < nb) bsum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(256) k_scan_final(const uint32_t *in, uint32_t n, const uint64_t *bsum, uint64_t *out) {
    __shared__ uint64_t ws[8];
    const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
    uint32_t a[4];
    uint64_t v = 0;
    for (int i = 0; i < 4; i++) { a[i] = base + i < n ? in[base + i] : 0; v += a[i]; }
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(v, ws, &tot) + bsum[blockIdx.x];
    for (int i = 0; i < 4; i++) { if (base + i < n) out[base + i] = ex; ex += a[i]; }
}

// ------------------------------------------------------------------------------------------
// re-emission with an exactly sized buffer (only when the optimistic capacity of k_resolve was too small)
// ------------------------------------------------------------------------------------------

template <int W, bool BE>
__global__ void __launch_bounds__(128)
k_emit(const __grid_constant__ MmgProgram P, const __grid_constant__ MmgGeom G, const __grid_constant__ MmgScratch X,
       uint64_t *out_off, uint32_t *out_val) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G.nsub || X.ext[t].y == 0) return;
    if (X.mcount[t] == 0) return;
    emit_subtile<W, BE>(P, G, X, t, X.mbase[t], out_off, out_val);
}

// ------------------------------------------------------------------------------------------
// G*: generic path -- one thread walks one (block, alignment) chain, any geometry
// ------------------------------------------------------------------------------------------

template <int W, bool BE>
__global__ void __launch_bounds__(128)
g_walk(const __grid_constant__ MmgProgram P, const __grid_constant__ MmgGeom G, uint32_t *counts, const uint64_t *bases,
       uint64_t *out_off, uint32_t *out_val) {
    const uint64_t chain = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= (uint64_t)G.nblocks * G.npads) return;
    const uint64_t bi = chain / G.npads, pad = chain % G.npads;
    const uint64_t off = bi * G.B;
    const uint64_t size = min(G.B + (uint64_t)G.ov, G.S - off);
    uint64_t count = size / W;
    if (pad + count * W > size) count -= 1;
    const uint8_t *base = G.data + off + pad;
    const uint64_t L = (uint64_t)P.L;
    uint64_t k = 0;
    uint32_t n = 0;
    uint64_t at = bases ? bases[chain] : 0;
    while (k + L <= count) {
        const uint32_t r = eval_window<W, BE>(P, base + k * W);
        if (r & 0x100u) {
            if (bases) {
                const uint64_t s = off + pad + k * W;
                out_off[at] = (G.base_offset + s) >> G.report_shift;
                const uint32_t v0 = ld_elem<W, BE>(G.data + s + (uint32_t)P.first_lit * W);
                const uint32_t v1 = P.opp_idx >= 0 ? ld_elem<W, BE>(G.data + s + (uint32_t)P.opp_idx * W) : 0u;
                out_val[at] = v0 | (v1 << 16);
                at++;
            }
            n++;
        }
        k += G.complete ? 1u : (r & 0xFFu);
    }
    if (!bases) counts[chain] = n;
}

// ---- keywords longer than MMG_MAXL: the same walk with the program's arrays in device memory (MmgLongProgram)
template <int W, bool BE>
__device__ __forceinline__ uint32_t eval_window_long(const MmgLongProgram &P, const uint8_t *w) {      // bit 31: match
    const uint32_t vmask = W == 1 ? 0xFFu : 0xFFFFu;
    for (int c = 0; c < P.ncheck; c++) {
        const MmgCheck k = P.chk[c];
        const int d = (int)ld_elem<W, BE>(w + (int)k.i * W) - (int)ld_elem<W, BE>(w + ((int)k.i - (int)k.lag) * W);
        const bool pass = P.modular ? ((((uint32_t)(d - k.ed)) & vmask) == 0) : (d == k.ed);
        if (!pass) {
            int sk = P.tab_default;
            for (int j = 0; j < P.ntab; j++)
                if (P.tab_key[j] == d) sk = P.tab_val[j];
            return (uint32_t)min(k.cap, sk);
        }
    }
    return 0x80000000u | (uint32_t)P.match_jump;
}

template <int W, bool BE>
__global__ void __launch_bounds__(128)
g_walk_long(const __grid_constant__ MmgLongProgram P, const __grid_constant__ MmgGeom G, uint32_t *counts, const uint64_t *bases,
            uint64_t *out_off, uint32_t *out_val) {
    const uint64_t chain = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= (uint64_t)G.nblocks * G.npads) return;
    const uint64_t bi = chain / G.npads, pad = chain % G.npads;
    const uint64_t off = bi * G.B;
    const uint64_t size = min(G.B + (uint64_t)G.ov, G.S - off);
    uint64_t count = size / W;
    if (pad + count * W > size) count -= 1;
    const uint8_t *base = G.data + off + pad;
    const uint64_t L = (uint64_t)P.L;
    uint64_t k = 0;
    uint32_t n = 0;
    uint64_t at = bases ? bases[chain] : 0;
    while (k + L <= count) {
        const uint32_t r = eval_window_long<W, BE>(P, base + k * W);
        if (r & 0x80000000u) {
            if (bases) {
                const uint64_t s = off + pad + k * W;
                out_off[at] = (G.base_offset + s) >> G.report_shift;
                const uint32_t v0 = ld_elem<W, BE>(G.data + s + (uint32_t)P.first_lit * W);
                const uint32_t v1 = P.opp_idx >= 0 ? ld_elem<W, BE>(G.data + s + (uint32_t)P.opp_idx * W) : 0u;
                out_val[at] = v0 | (v1 << 16);
                at++;
            }
            n++;
        }
        k += G.complete ? 1u : (r & 0x7FFFFFFFu);
    }
    if (!bases) counts[chain] = n;
}

// two alignments of one block interleave by offset: merge them (offsets are unique)
__global__ void __launch_bounds__(128)
g_merge(uint32_t nblocks, const uint32_t *counts, const uint64_t *bases, const uint64_t *in_off, const uint32_t *in_val,
        uint64_t *out_off, uint32_t *out_val) {
    const uint64_t chain = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= (uint64_t)nblocks * 2) return;
    const uint64_t sib = chain ^ 1;
    const uint64_t blockbase = bases[chain & ~1ull];
    const uint64_t *mine = in_off + bases[chain], *other = in_off + bases[sib];
    const uint32_t nm = counts[chain], no = counts[sib];
    for (uint32_t a = 0; a < nm; a++) {
        const uint64_t v = mine[a];
        uint32_t lo = 0, hi = no;   // elements of the sibling list smaller than v
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (other[mid] < v) lo = mid + 1; else hi = mid; }
        out_off[blockbase + a + lo] = v;
        out_val[blockbase + a + lo] = in_val[bases[chain] + a];
    }
}

// ------------------------------------------------------------------------------------------
// synthetic ROM generator (bench / tests): byte i = byte (i mod 8) of splitmix64(seed ^ (i / 8)), AND mask
// (counter based, chunk addressable, reproducible on CPU -- see synth.py)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_synth(uint64_t *out, uint64_t nwords, uint64_t seed, uint64_t first_word, uint64_t mask8) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride)
        out[i] = splitmix64(seed ^ (first_word + i)) & mask8;
}

}  // namespace

cudaError_t mmg_launch_synth(uint64_t *out, uint64_t nwords, uint64_t seed, uint64_t first_word, uint32_t byte_mask,
                             cudaStream_t stream) {
    if (nwords == 0) return cudaSuccess;
    const uint64_t mask8 = 0x0101010101010101ull * (uint64_t)(byte_mask & 0xFFu);
    const unsigned grid = (unsigned)std::min<uint64_t>((nwords + 255) / 256, 148ull * 16);
    k_synth<<<grid, 256, 0, stream>>>(out, nwords, seed, first_word, mask8);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// launch helpers (called from capi.cu)
// ------------------------------------------------------------------------------------------

#include "launch.h"

// NK (compile-time key count): the 8-bit filter unrolls up to 8 keys (lags up to 4 bytes), 16-bit fuses the single-key case.
template <int W, int LB, bool BE>
static const void *filter_for_keys(int nkeys) {
    if (W == 1 && LB != 0) {
        constexpr int L8 = LB ? LB : 1;
        if (LB > 4) return (const void *)k_filter8<L8, 0>;      // long wildcard gaps: run-time key loop only
        switch (nkeys) {
            case 1: return (const void *)k_filter8<L8, 1>;
            case 2: return (const void *)k_filter8<L8, 2>;
            case 3: return (const void *)k_filter8<L8, 3>;
            case 4: return (const void *)k_filter8<L8, 4>;
            case 5: return (const void *)k_filter8<L8, 5>;
            case 6: return (const void *)k_filter8<L8, 6>;
            case 7: return (const void *)k_filter8<L8, 7>;
            case 8: return (const void *)k_filter8<L8, 8>;
            default: return (const void *)k_filter8<L8, 0>;
        }
    }
    if (LB == 0) return (const void *)k_filter<W, LB, BE, 0>;
    return nkeys == 1 ? (const void *)k_filter<W, LB, BE, 1> : (const void *)k_filter<W, LB, BE, 0>;
}

#define FILTER_CASE(W_, LB_)                                                                             \
    case LB_:                                                                                            \
        fn = (W_ == 2 && be) ? filter_for_keys<W_, LB_, (W_ == 2)>(nkeys) : filter_for_keys<W_, LB_, false>(nkeys); \
        break;

// Picks the filter instantiation for (W, lag bytes, endianness, key count); nullptr when the lag is not tiled.
static const void *filter_kernel(int W, int lag_bytes, bool be, int nkeys) {
    const void *fn = nullptr;
    if (W == 1) {
        switch (lag_bytes) {
            FILTER_CASE(1, 0) FILTER_CASE(1, 1) FILTER_CASE(1, 2) FILTER_CASE(1, 3) FILTER_CASE(1, 4)
            FILTER_CASE(1, 5) FILTER_CASE(1, 6) FILTER_CASE(1, 7) FILTER_CASE(1, 8)
            default: break;
        }
    } else {
        switch (lag_bytes) {
            FILTER_CASE(2, 0) FILTER_CASE(2, 2) FILTER_CASE(2, 4) FILTER_CASE(2, 6) FILTER_CASE(2, 8)
            default: break;
        }
    }
    return fn;
}

static size_t filter_smem(int W, int lag_bytes) {
    return (size_t)MMG_FILTER_WARPS * ((W == 1 && lag_bytes != 0) ? MMG_WARP_SMEM8 : MMG_WARP_SMEM);
}

bool mmg_filter_supported(int W, int lag_bytes) { return filter_kernel(W, lag_bytes, false, 1) != nullptr; }

cudaError_t mmg_filter_occupancy(int W, int lag_bytes, bool be, int nkeys, int *blocks_per_sm) {
    const void *fn = filter_kernel(W, lag_bytes, be, nkeys);
    if (!fn) return cudaErrorInvalidValue;
    // per (device, kernel) cache: the attribute call and the occupancy query cost microseconds per scan otherwise
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, fn});
    if (it != cache.end()) { *blocks_per_sm = it->second; return cudaSuccess; }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)filter_smem(W, lag_bytes));
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, MMG_FILTER_WARPS * 32, filter_smem(W, lag_bytes));
    if (e == cudaSuccess) cache[{dev, fn}] = *blocks_per_sm;
    return e;
}

cudaError_t mmg_launch_filter(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, int lag_bytes, int grid,
                              cudaStream_t stream) {
    const void *fn = filter_kernel(P.W, lag_bytes, G.big_endian != 0, P.nkeys);
    if (!fn) return cudaErrorInvalidValue;
    void *args[] = {(void *)&P, (void *)&G, (void *)&X};
    // a fused scan ends in a grid-wide barrier: cooperative launch guarantees that the whole grid is resident
    if (X.fuse) return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(MMG_FILTER_WARPS * 32), args, filter_smem(P.W, lag_bytes), stream);
    return cudaLaunchKernel(fn, dim3(grid), dim3(MMG_FILTER_WARPS * 32), args, filter_smem(P.W, lag_bytes), stream);
}

// chain slices: segments of the slice's own block (a trailing overlap sub-tile may add one more segment to the grid)
static uint32_t chain_segments(const MmgGeom &G) { return min(G.segs_per_block, G.nseg); }

#define RESOLVE_SPLIT 4
#define RESOLVE_SPLIT_MAX_CTAS 592u     // up to four CTAs per SM: beyond that the SMs are busy without the helper warps

template <bool MAPS_ONLY>
static cudaError_t launch_resolve_kernel(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                                         uint32_t *out_val, uint64_t capacity, cudaStream_t stream) {
    const unsigned grid = G.nseg;
    const uint32_t jp = (uint32_t)((P.Jmax + 15) / 16 * 16);
    const bool fast = P.J0 == P.Jmax && (uint32_t)P.Jmax <= RESOLVE_FAST_J;
    static const bool nosplit = getenv("MMG_NO_RESOLVE_SPLIT") != nullptr;
    const bool one_round = G.segs_per_block > 1 || G.spb <= RESOLVE_THREADS;       // (always, but for the MMG_NOSEG debug switch)
    const bool split = fast && P.W == 1 && grid <= RESOLVE_SPLIT_MAX_CTAS && one_round && !nosplit;
    const size_t smem = (fast ? (size_t)SE_STRIDE * G.npads * RESOLVE_FAST_J * 4 * (split ? RESOLVE_SPLIT : 1)
                              : (size_t)RESOLVE_THREADS * G.npads * jp) + RESOLVE_THREADS * 4;
    if (split) k_resolve<1, false, MAPS_ONLY, RESOLVE_SPLIT><<<grid, RESOLVE_THREADS * RESOLVE_SPLIT, smem, stream>>>(P, G, X, out_off, out_val, capacity, jp);
    else if (P.W == 1) k_resolve<1, false, MAPS_ONLY, 1><<<grid, RESOLVE_THREADS, smem, stream>>>(P, G, X, out_off, out_val, capacity, jp);
    else if (G.big_endian) k_resolve<2, true, MAPS_ONLY, 1><<<grid, RESOLVE_THREADS, smem, stream>>>(P, G, X, out_off, out_val, capacity, jp);
    else k_resolve<2, false, MAPS_ONLY, 1><<<grid, RESOLVE_THREADS, smem, stream>>>(P, G, X, out_off, out_val, capacity, jp);
    return cudaGetLastError();
}

// (the two-level phase prefix of the chain kernels also serves search() on one large buffer: a single block of
// thousands of segments, which one warp of k_segphase would walk alone)
static void chain_ranges(const MmgGeom &G, uint32_t &nsegs, uint32_t &per_range, uint32_t &nranges) {
    nsegs = chain_segments(G);
    per_range = (nsegs + CHAIN_RANGES - 1) / CHAIN_RANGES;
    nranges = (nsegs + per_range - 1) / per_range;
}

static cudaError_t launch_chainphase(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, cudaStream_t stream) {
    const uint32_t jp = (uint32_t)((P.Jmax + 15) / 16 * 16);
    uint32_t nsegs, per_range, nranges;
    chain_ranges(G, nsegs, per_range, nranges);
    const size_t smem = (size_t)CHAIN_RANGES * 2 * jp + (size_t)CHAIN_WARPS * 32 * 2 * jp;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_chainphase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_chainphase<<<(nranges + CHAIN_WARPS - 1) / CHAIN_WARPS, CHAIN_WARPS * 32, smem, stream>>>(G, X, jp, nsegs, per_range, nranges);
    return cudaGetLastError();
}

static cudaError_t launch_rangemap(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, cudaStream_t stream) {
    const uint32_t jp = (uint32_t)((P.Jmax + 15) / 16 * 16);
    uint32_t nsegs, per_range, nranges;
    chain_ranges(G, nsegs, per_range, nranges);
    k_rangemap<<<(nranges + CHAIN_WARPS - 1) / CHAIN_WARPS, CHAIN_WARPS * 32, (size_t)CHAIN_WARPS * 32 * 2 * jp, stream>>>(
        G, X, jp, (uint32_t)P.Jmax, nsegs, per_range, nranges);
    return cudaGetLastError();
}

bool mmg_resolve_two_level(const MmgGeom &G) { return G.nblocks == 1 && G.segs_per_block > CHAIN_RANGES; }

cudaError_t mmg_launch_resolve(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                               uint32_t *out_val, uint64_t capacity, cudaStream_t stream) {
    if (G.segs_per_block > 1) {
        // blocks cut into segments: segment maps, then the phase prefix along each block
        cudaError_t e = launch_resolve_kernel<true>(P, G, X, out_off, out_val, capacity, stream);
        if (e != cudaSuccess) return e;
        if (mmg_resolve_two_level(G)) {
            e = launch_rangemap(P, G, X, stream);
            if (e == cudaSuccess) e = launch_chainphase(P, G, X, stream);
        } else {
            const uint32_t jp = (uint32_t)((P.Jmax + 15) / 16 * 16);
            k_segphase<<<(G.nblocks + SEGPHASE_WARPS - 1) / SEGPHASE_WARPS, SEGPHASE_WARPS * 32, (size_t)SEGPHASE_WARPS * 32 * 2 * jp, stream>>>(G, X, jp);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return e;
    }
    return launch_resolve_kernel<false>(P, G, X, out_off, out_val, capacity, stream);
}

// slice of a longer chain, first half: segment maps -> range maps -> the slice's map (pinned host buffer) + event counts
cudaError_t mmg_launch_chain_maps(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, cudaStream_t stream) {
    const uint32_t jp = (uint32_t)((P.Jmax + 15) / 16 * 16);
    uint32_t nsegs, per_range, nranges;
    chain_ranges(G, nsegs, per_range, nranges);
    cudaError_t e = launch_resolve_kernel<true>(P, G, X, nullptr, nullptr, 0, stream);
    if (e == cudaSuccess) e = launch_rangemap(P, G, X, stream);
    if (e != cudaSuccess) return e;
    k_slicemap<<<1, 256, (size_t)nranges * 2 * jp, stream>>>(G, X, jp, (uint32_t)P.Jmax, nranges);
    return cudaGetLastError();
}

// second half, with G.entry set: entry phase of every segment, then the resolve kernel
cudaError_t mmg_launch_chain_resolve(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                                     uint32_t *out_val, uint64_t capacity, cudaStream_t stream) {
    cudaError_t e = launch_chainphase(P, G, X, stream);
    if (e != cudaSuccess) return e;
    return launch_resolve_kernel<false>(P, G, X, out_off, out_val, capacity, stream);
}

bool mmg_sparse_resolve_supported(const MmgGeom &G) { return G.segs_per_block == 1 && G.spb <= 128; }

// exclusive scan of n u32 counts into u64 bases; bsum must hold ceil(n/1024) entries; *total receives the sum
cudaError_t mmg_launch_scan(const uint32_t *counts, uint32_t n, uint64_t *bsum, uint64_t *bases, uint64_t *total,
                            cudaStream_t stream) {
    if (n == 0) return cudaMemsetAsync(total, 0, sizeof(uint64_t), stream);
    const uint32_t nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    k_scan_sums<<<nb, 256, 0, stream>>>(counts, n, bsum);
    k_scan_top<<<1, 256, 0, stream>>>(bsum, nb, total);
    k_scan_final<<<nb, 256, 0, stream>>>(counts, n, bsum, bases);
    return cudaGetLastError();
}

cudaError_t mmg_launch_emit(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                            uint32_t *out_val, cudaStream_t stream) {
    const unsigned grid = (G.nsub + 127) / 128;
    if (P.W == 1) k_emit<1, false><<<grid, 128, 0, stream>>>(P, G, X, out_off, out_val);
    else if (G.big_endian) k_emit<2, true><<<grid, 128, 0, stream>>>(P, G, X, out_off, out_val);
    else k_emit<2, false><<<grid, 128, 0, stream>>>(P, G, X, out_off, out_val);
    return cudaGetLastError();
}

cudaError_t mmg_launch_generic_walk(const MmgProgram &P, const MmgGeom &G, uint32_t *counts, const uint64_t *bases,
                                    uint64_t *out_off, uint32_t *out_val, cudaStream_t stream) {
    const uint64_t chains = (uint64_t)G.nblocks * G.npads;
    const unsigned grid = (unsigned)((chains + 127) / 128);
    if (P.W == 1) g_walk<1, false><<<grid, 128, 0, stream>>>(P, G, counts, bases, out_off, out_val);
    else if (G.big_endian) g_walk<2, true><<<grid, 128, 0, stream>>>(P, G, counts, bases, out_off, out_val);
    else g_walk<2, false><<<grid, 128, 0, stream>>>(P, G, counts, bases, out_off, out_val);
    return cudaGetLastError();
}

cudaError_t mmg_launch_generic_walk_long(const MmgLongProgram &P, const MmgGeom &G, uint32_t *counts, const uint64_t *bases,
                                         uint64_t *out_off, uint32_t *out_val, cudaStream_t stream) {
    const uint64_t chains = (uint64_t)G.nblocks * G.npads;
    const unsigned grid = (unsigned)((chains + 127) / 128);
    if (P.W == 1) g_walk_long<1, false><<<grid, 128, 0, stream>>>(P, G, counts, bases, out_off, out_val);
    else if (G.big_endian) g_walk_long<2, true><<<grid, 128, 0, stream>>>(P, G, counts, bases, out_off, out_val);
    else g_walk_long<2, false><<<grid, 128, 0, stream>>>(P, G, counts, bases, out_off, out_val);
    return cudaGetLastError();
}

cudaError_t mmg_launch_generic_merge(uint32_t nblocks, const uint32_t *counts, const uint64_t *bases,
                                     const uint64_t *in_off, const uint32_t *in_val, uint64_t *out_off,
                                     uint32_t *out_val, cudaStream_t stream) {
    const uint64_t chains = (uint64_t)nblocks * 2;
    g_merge<<<(unsigned)((chains + 127) / 128), 128, 0, stream>>>(nblocks, counts, bases, in_off, in_val, out_off, out_val);
    return cudaGetLastError();
}

#ifdef MMG_RESOLVE_PROF
extern "C" void mmg_debug_resolve_phases(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_ns, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase_ns, z, sizeof(z)); }
}
#endif
