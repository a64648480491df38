// Word-wide fast path of the GAF record parser (host/device inline code, like ptx_core.cuh).
//
// parse_record (ptx_core.cuh) walks a line one byte at a time.  The code here reads the same columns from
//   * a tab bitmap and a newline bitmap of the staged text (one bit per byte, built once per tile by all
//     threads from 16-byte pieces: classify16),
//   * unaligned 4/8-byte words of the text (two aligned loads + a funnel shift: ld4 / ld8),
//   * SWAR decimal conversion (four digits per multiply-add pair: swar4) and word-wise id hashing,
// and touches no byte individually.  It accepts the records whose columns have the plain shape every aligner
// writes - 12+ tab-separated columns, integer columns that are 1-8 unsigned digits or any non-numeric text
// ('*' = null), walk ids of at most 9 digits, "\n" line ends - and reports
// everything else (signs, 9+ digit integers, 10+ digit ids, "\r\n", fewer than 12 columns, lines that leave the
// staged window) as "not handled"; those records go through parse_record, which stays the definition of the
// dialect.  On a record it accepts, fast_parse returns exactly what parse_record returns
// (tests/hostcheck.cpp runs both on every record of every CPU test and of the dialect fuzzer).
//
// Reference semantics: GAF columns 1,2,6,7,8,9,12 (rcls.rs:119-146), digit runs of the walk (rcls.rs:237-258).
#pragma once
#include "ptx_core.cuh"

namespace ptx {

PTX_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // low word of (hi:lo) >> (sh & 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
PTX_HD uint32_t ffs32(uint32_t m) {  // index of the lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)m) - 1u;
#else
    return (uint32_t)__builtin_ctz(m);
#endif
}

// ---- byte classes, four bytes at a time: 0x80 in every byte of the result that belongs to the class
PTX_HD uint32_t eq_mask4(uint32_t w, uint32_t rep) {  // bytes equal to the byte replicated in `rep`
    const uint32_t y = w ^ rep;
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u;
}
PTX_HD uint32_t nondigit_mask4(uint32_t w) {  // bytes outside '0'..'9'
    const uint32_t t = w ^ 0x30303030u;
    return (((t & 0x7f7f7f7fu) + 0x76767676u) | t) & 0x80808080u;
}
// Two 0x80-per-byte masks (text bytes 0-3 and 4-7) -> 8 bits, bit i = byte i.  One multiply gathers the eight
// flags: the partial products land on pairwise different bit positions (3,7,..,31 shifted by 0,7,14,21 fall into
// four different residue classes mod 4), so nothing carries.
PTX_HD uint32_t pack8(uint32_t m03, uint32_t m47) { return ((m47 | (m03 >> 4)) * 0x00204081u) >> 24; }

// sixteen 0x80-per-byte flags (four mask words, text order) -> 16 bits, bit i = byte i.  Four dot products weigh the flags
// (128 * 2^i per byte, the second word of a pair 16 times as much), one multiply-add joins the halves: IDP.4A and IMAD run
// on the FMA pipe, beside the LOP3/IADD3 stream that forms the masks.
PTX_HD uint32_t dot4(uint32_t flags, uint32_t weights, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(flags, weights, acc);
#else
    for (int i = 0; i < 4; ++i) acc += ((flags >> (8 * i)) & 0xFFu) * ((weights >> (8 * i)) & 0xFFu);
    return acc;
#endif
}
PTX_HD uint32_t pack16(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t lo = dot4(b, 0x80402010u, dot4(a, 0x08040201u, 0u)), hi = dot4(d, 0x80402010u, dot4(c, 0x08040201u, 0u));
    return (hi * 256u + lo) >> 7;
}
// bytes equal to the 7-bit byte replicated in `rep`: 0x80 flags (w7 = w & 0x7f7f7f7f, shared between the classes)
PTX_HD uint32_t eq7_mask4(uint32_t w, uint32_t w7, uint32_t rep) { return ~((w7 ^ rep) + 0x7f7f7f7fu) & ~w & 0x80808080u; }

PTX_HD void classify16_mul(uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t& nl16, uint32_t& tab16) {  // multiply-packed (A/B)
    nl16 = pack8(eq_mask4(x, 0x0a0a0a0au), eq_mask4(y, 0x0a0a0a0au)) | (pack8(eq_mask4(z, 0x0a0a0a0au), eq_mask4(w, 0x0a0a0a0au)) << 8);
    tab16 = pack8(eq_mask4(x, 0x09090909u), eq_mask4(y, 0x09090909u)) | (pack8(eq_mask4(z, 0x09090909u), eq_mask4(w, 0x09090909u)) << 8);
}
// newline and tab flags of one 16-byte piece of text, bit i = byte i
PTX_HD void classify16(uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t& nl16, uint32_t& tab16) {
    const uint32_t x7 = x & 0x7f7f7f7fu, y7 = y & 0x7f7f7f7fu, z7 = z & 0x7f7f7f7fu, w7 = w & 0x7f7f7f7fu;
    nl16 = pack16(eq7_mask4(x, x7, 0x0a0a0a0au), eq7_mask4(y, y7, 0x0a0a0a0au), eq7_mask4(z, z7, 0x0a0a0a0au), eq7_mask4(w, w7, 0x0a0a0a0au));
    tab16 = pack16(eq7_mask4(x, x7, 0x09090909u), eq7_mask4(y, y7, 0x09090909u), eq7_mask4(z, z7, 0x09090909u), eq7_mask4(w, w7, 0x09090909u));
}

// non-digit flags of one 16-byte piece (the long-read kernel finds the ends of the walk ids from them), bit i = byte i
PTX_HD uint32_t nondigit16(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    return pack16(nondigit_mask4(x), nondigit_mask4(y), nondigit_mask4(z), nondigit_mask4(w));
}

// ---- the staged window as aligned 32-bit words (little endian).  On the device it lives in shared memory and is read
// with explicit ld.shared from a 32-bit shared address (a generic pointer made the compiler rebuild the shared base
// inside every loop); on the host (tests/hostcheck.cpp) it is a plain array.
struct Words {
    uint32_t base;      // device: shared-memory address of word 0
    const uint32_t* p;  // host: the array
    PTX_HD uint32_t at(uint32_t byte_off) const {  // word at a 4-byte aligned byte offset
#if defined(__CUDA_ARCH__)
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
        return v;
#else
        return p[byte_off >> 2];
#endif
    }
    PTX_HD uint32_t word(uint32_t i) const { return at(i << 2); }
};

// unaligned text words: 4 / 8 bytes starting at byte p
PTX_HD uint32_t ld4(const Words& W, uint32_t p) {
    const uint32_t o = p & ~3u;
    return funnel_r(W.at(o), W.at(o + 4u), p << 3);
}
PTX_HD void ld8(const Words& W, uint32_t p, uint32_t& lo, uint32_t& hi) {
    const uint32_t o = p & ~3u, sh = p << 3;
    const uint32_t w0 = W.at(o), w1 = W.at(o + 4u), w2 = W.at(o + 8u);
    lo = funnel_r(w0, w1, sh);
    hi = funnel_r(w1, w2, sh);
}

// four decimal digits held as byte VALUES 0..9, byte 0 the most significant -> their number: one dot product weighs the
// first three (100, 10, 1), one multiply-add appends the fourth - both on the FMA pipe.  (Bytes that are no digits give some
// number; the callers throw it away.)
PTX_HD uint32_t swar4(uint32_t v) { return dot4(v, 0x00010A64u, 0u) * 10u + (v >> 24); }
// n <= 8 digit bytes (already XORed with '0') at the low end of (hi:lo) -> right-aligned: byte 7 = last digit
PTX_HD void align8(uint32_t lo, uint32_t hi, uint32_t n, uint32_t& xl, uint32_t& xh) {
    const uint64_t x = (((uint64_t)hi << 32) | (uint64_t)lo) << (8u * (8u - n));
    xl = (uint32_t)x;
    xh = (uint32_t)(x >> 32);
}

// cursor over the set bits of a bitmap (bit i of word j = byte 32 j + i).  A word of all ones lies behind the
// bitmap of the window, so next() always terminates; positions at or beyond the window mean "not found".
struct BitCursor {
    Words w;
    uint32_t wo, m;  // byte offset of the current bitmap word, its remaining bits
    PTX_HD void seek(const Words& words, uint32_t pos) {
        w = words;
        wo = (pos >> 5) << 2;
        m = w.at(wo) & (0xFFFFFFFFu << (pos & 31u));
    }
    PTX_HD uint32_t next() {
        while (m == 0u) { wo += 4u; m = w.at(wo); }
        const uint32_t b = ffs32(m);
        m &= m - 1u;
        return (wo << 3) + b;
    }
    PTX_HD void skip() {  // consume one set bit without locating it
        while (m == 0u) { wo += 4u; m = w.at(wo); }
        m &= m - 1u;
    }
};

struct FastRec {
    IdHash h;
    uint32_t qlen, c7, c8, c9, mapq;  // valid unless the matching bit of `nulls` is set
    uint32_t nulls;                   // bit 0 qlen, 1 c7, 2 c8, 3 c9, 4 mapq
    uint32_t W, vmin, vmax;           // walk nodes (ids < 10^9), min / max id (W > 0)
    uint32_t path_pos, path_end;      // bytes [pos,end) of column 6
    bool path_null;
};
constexpr uint32_t FN_QLEN = 1u, FN_C7 = 2u, FN_C8 = 4u, FN_C9 = 8u, FN_MAPQ = 16u;

// One integer column [a,b): `[0-9]{1,8}` -> value; empty or non-numeric text -> null (parse_int_field: junk in an
// integer column is null); a sign or more than 8 bytes -> bit 0 of `slow` (the exact parser decides).  Branch-free.
PTX_HD uint32_t fast_int(const Words& W, uint32_t a, uint32_t b, uint32_t null_bit, uint32_t& nulls, uint32_t& slow) {
    const uint32_t n = b - a;
    uint32_t lo, hi;
    ld8(W, a, lo, hi);
    lo ^= 0x30303030u;
    hi ^= 0x30303030u;
    const uint32_t fits = (n - 1u <= 7u) ? 1u : 0u;  // 1..8 bytes
    uint32_t xl, xh;
    align8(lo, hi, fits ? n : 8u, xl, xh);
    const uint32_t bad = (((((xl + 0x76767676u) | xl) | ((xh + 0x76767676u) | xh)) & 0x80808080u) != 0u) ? 1u : 0u;  // some byte is not 0..9
    const uint32_t c0 = lo & 0xFFu;                                                  // '+' ^ '0' = 0x1b, '-' ^ '0' = 0x1d
    const uint32_t is_sign = (c0 == 0x1bu || c0 == 0x1du) ? 1u : 0u;
    slow |= (n > 8u ? 1u : 0u) | (fits & bad & is_sign);
    const uint32_t isnull = (fits ^ 1u) | bad;
    nulls |= isnull ? null_bit : 0u;
    return isnull ? 0u : swar4(xl) * 10000u + swar4(xh);
}

// Columns 1-5 and 7-12 of the line [s, e] of the window, e = position of the '\n' that ends it: id hash, the five integer
// columns, the extent of column 6.  lim = bytes of text in the window (the tab bitmap covers exactly those, then a sentinel
// word of all ones; the text is readable 128 bytes beyond).  Returns ok (12 columns inside the window, "\n" line end);
// `slow` gets bit 0 if an integer column needs the exact parser.  `mask` = lanes that call this together.
PTX_HD bool fast_cols(const Words& W, const Words& tabw, uint32_t s, uint32_t e, uint32_t lim, FastRec& r, uint32_t mask, uint32_t& slow) {
    BitCursor tc;
    tc.seek(tabw, s);
    const uint32_t t1 = tc.next(), t2 = tc.next();
    tc.skip();
    tc.skip();
    const uint32_t t5 = tc.next(), t6 = tc.next(), t7 = tc.next(), t8 = tc.next(), t9 = tc.next();
    tc.skip();
    const uint32_t t11 = tc.next(), t12 = tc.next();
    // 12 columns inside the window, and not a "\r\n" line (the '\r' would end the last column, term_at)
    bool ok = e < lim && t11 < e;
    if (ok) ok = ((W.at((e - 1u) & ~3u) >> (((e - 1u) & 3u) * 8u)) & 0xFFu) != (uint32_t)'\r';
    r.nulls = 0;
    r.qlen = r.c7 = r.c8 = r.c9 = r.mapq = 0;
    r.W = 0;
    r.vmin = 0xFFFFFFFFu;
    r.vmax = 0;
    r.path_pos = t5 + 1u;
    r.path_end = t6;
    r.path_null = false;
    r.h.lo = 1;
    r.h.hi = 0;
#if defined(__CUDA_ARCH__)
    const uint32_t okmask = __ballot_sync(mask, ok);  // the lanes that go through the columns together
#else
    const uint32_t okmask = mask;
#endif
    if (ok) {
        {   // column 1: the id hash over 4-byte words (same value as IdHasher::byte over the bytes)
            IdHasher H;
            const uint32_t n = t1 - s;
            uint32_t p = s;
            for (uint32_t k = n >> 2; k; --k, p += 4u) H.mix(ld4(W, p));
            if (n & 3u) H.mix(ld4(W, p) & (0xFFFFFFFFu >> (32u - 8u * (n & 3u))));
            H.nbytes = n;
            H.word = 0;
            r.h = H.finish_words();
        }
        PTX_RECONVERGE(okmask);
        r.qlen = fast_int(W, t1 + 1u, t2, FN_QLEN, r.nulls, slow);
        r.c7 = fast_int(W, t6 + 1u, t7, FN_C7, r.nulls, slow);
        r.c8 = fast_int(W, t7 + 1u, t8, FN_C8, r.nulls, slow);
        r.c9 = fast_int(W, t8 + 1u, t9, FN_C9, r.nulls, slow);
        r.mapq = fast_int(W, t11 + 1u, t12 < e ? t12 : e, FN_MAPQ, r.nulls, slow);
        r.path_null = (t6 - t5 - 1u == 1u) && ((ld4(W, t5 + 1u) & 0xFFu) == (uint32_t)'*');
    }
    PTX_RECONVERGE(mask);
    return ok;
}

// One walk id of n (1..9) digits starting at byte a -> its value (ids < 10^9)
PTX_HD uint32_t fast_node(const Words& W, uint32_t a, uint32_t n) {
    uint32_t top = 0, n8 = n;
    if (n > 8u) {  // rare
        top = ((ld4(W, a) & 0xFFu) - (uint32_t)'0') * 100000000u;
        ++a;
        n8 = 8u;
    }
    uint32_t lo, hi, xl, xh;
    ld8(W, a, lo, hi);
    align8(lo ^ 0x30303030u, hi ^ 0x30303030u, n8, xl, xh);
    return top + swar4(xl) * 10000u + swar4(xh);
}

// Columns 1..12 of the line [s, e] with the walk decoded by the calling thread (the short-read kernel: a handful of nodes
// per record).  Returns false if the record needs the exact parser.  Only the first stash_cap walk ids are stashed; W,
// vmin and vmax cover the whole walk (the caller decodes a longer walk again when it writes it out).
PTX_HD bool fast_parse(const Words& W, const Words& tabw, uint32_t s, uint32_t e, uint32_t lim, FastRec& r, uint32_t mask,
                       uint32_t* stash, uint32_t stash_stride, uint32_t stash_cap) {
    uint32_t slow = 0;
    const bool ok = fast_cols(W, tabw, s, e, lim, r, mask, slow);
    if (ok) {
        // column 6: every non-digit byte (the closing tab included) ends the digit run in front of it
        const uint32_t p6 = r.path_pos, e6 = r.path_end;
        uint32_t last_sep = p6 - 1u, Wn = 0, vmin = 0xFFFFFFFFu, vmax = 0;
        for (uint32_t p = p6; p <= e6; p += 32u) {
            const uint32_t nbits = (e6 - p + 1u) < 32u ? (e6 - p + 1u) : 32u;  // bytes of [p6, e6] in this segment
            uint32_t nd = 0;
            {
                const uint32_t sh = p << 3;
                uint32_t o = p & ~3u;
                uint32_t wa = W.at(o);
                for (uint32_t j = 0; j < nbits; j += 8u, o += 8u) {
                    const uint32_t wb = W.at(o + 4u), wc = W.at(o + 8u);
                    nd |= (dot4(nondigit_mask4(funnel_r(wb, wc, sh)), 0x80402010u, dot4(nondigit_mask4(funnel_r(wa, wb, sh)), 0x08040201u, 0u)) >> 7) << j;
                    wa = wc;
                }
            }
            if (nbits < 32u) nd &= (1u << nbits) - 1u;
            while (nd) {
                const uint32_t q = p + ffs32(nd);
                nd &= nd - 1u;
                const uint32_t n = q - last_sep - 1u;
                const uint32_t a = last_sep + 1u;
                last_sep = q;
                if (n) {
                    // ids of up to 9 digits; 10-18 digit ids are valid too: the exact parser reads those
                    slow |= n > 9u ? 1u : 0u;
                    const uint32_t v = fast_node(W, a, n < 9u ? n : 9u);
                    vmin = v < vmin ? v : vmin;
                    vmax = v > vmax ? v : vmax;
                    if (Wn < stash_cap) stash[Wn * stash_stride] = v;
                    ++Wn;
                }
            }
        }
        r.W = Wn;
        r.vmin = vmin;
        r.vmax = vmax;
    }
    PTX_RECONVERGE(mask);  // (never inside the loops: their trip counts differ from lane to lane)
    return ok && slow == 0u;
}

}  // namespace ptx
