// prep.cuh -- the reference's text normalisers (genz_tokenize/preprocess.py) as device byte filters: the optional
// step a user runs BEFORE Tokenize.  One warp walks one document 32 bytes at a time; every lane
// decides what its byte emits (0..3 bytes); ballots + carried "last position of ..." registers stand in for the
// regex engine's left-to-right state; lengths -> scan -> second pass writes.
//
//   OP_HTML     remove_html        preprocess.py:5-9      re.sub(r'<[^>]*>', '', txt)
//   OP_UNICODE  convert_unicode    preprocess.py:30-36    base letter + combining tone mark -> precomposed letter
//   OP_PUNCT    remove_punctuations preprocess.py:39-44   drop the 32 characters of string.punctuation
//   OP_EMOJI    remove_emoji       preprocess.py:47-72    drop the emoji class, then ' '.join(result.split())
//   OP_URL      remove_URL         preprocess.py:75-80    re.sub(r'http\S+', '', txt)
#pragma once
#include "device_common.cuh"
#include "encode.cuh"

namespace gzt {

enum PrepOp { OP_HTML = 0, OP_UNICODE = 1, OP_PUNCT = 2, OP_EMOJI = 3, OP_URL = 4 };

struct PrepArgs {
    const uint8_t* in;
    const int64_t* off;       // [n+1]
    int64_t n;
    int op;
    int64_t* out_len;         // pass 1
    const int64_t* out_off;   // pass 2
    uint8_t* out;
};

// base letter x tone mark (U+0300, 0301, 0303, 0309, 0323) -> precomposed letter: the Unicode canonical
// compositions of the 24 Vietnamese vowel bases (identical to the reference's 120-entry dicchar, preprocess.py:16-27)
__constant__ uint16_t PREP_COMPOSE[24][6] = {
    {0x0041, 0x00C0, 0x00C1, 0x00C3, 0x1EA2, 0x1EA0}, {0x0045, 0x00C8, 0x00C9, 0x1EBC, 0x1EBA, 0x1EB8}, {0x0049, 0x00CC, 0x00CD, 0x0128, 0x1EC8, 0x1ECA},
    {0x004F, 0x00D2, 0x00D3, 0x00D5, 0x1ECE, 0x1ECC}, {0x0055, 0x00D9, 0x00DA, 0x0168, 0x1EE6, 0x1EE4}, {0x0059, 0x1EF2, 0x00DD, 0x1EF8, 0x1EF6, 0x1EF4},
    {0x0061, 0x00E0, 0x00E1, 0x00E3, 0x1EA3, 0x1EA1}, {0x0065, 0x00E8, 0x00E9, 0x1EBD, 0x1EBB, 0x1EB9}, {0x0069, 0x00EC, 0x00ED, 0x0129, 0x1EC9, 0x1ECB},
    {0x006F, 0x00F2, 0x00F3, 0x00F5, 0x1ECF, 0x1ECD}, {0x0075, 0x00F9, 0x00FA, 0x0169, 0x1EE7, 0x1EE5}, {0x0079, 0x1EF3, 0x00FD, 0x1EF9, 0x1EF7, 0x1EF5},
    {0x00C2, 0x1EA6, 0x1EA4, 0x1EAA, 0x1EA8, 0x1EAC}, {0x00CA, 0x1EC0, 0x1EBE, 0x1EC4, 0x1EC2, 0x1EC6}, {0x00D4, 0x1ED2, 0x1ED0, 0x1ED6, 0x1ED4, 0x1ED8},
    {0x00E2, 0x1EA7, 0x1EA5, 0x1EAB, 0x1EA9, 0x1EAD}, {0x00EA, 0x1EC1, 0x1EBF, 0x1EC5, 0x1EC3, 0x1EC7}, {0x00F4, 0x1ED3, 0x1ED1, 0x1ED7, 0x1ED5, 0x1ED9},
    {0x0102, 0x1EB0, 0x1EAE, 0x1EB4, 0x1EB2, 0x1EB6}, {0x0103, 0x1EB1, 0x1EAF, 0x1EB5, 0x1EB3, 0x1EB7}, {0x01A0, 0x1EDC, 0x1EDA, 0x1EE0, 0x1EDE, 0x1EE2},
    {0x01A1, 0x1EDD, 0x1EDB, 0x1EE1, 0x1EDF, 0x1EE3}, {0x01AF, 0x1EEA, 0x1EE8, 0x1EEE, 0x1EEC, 0x1EF0}, {0x01B0, 0x1EEB, 0x1EE9, 0x1EEF, 0x1EED, 0x1EF1},
};

// ---- helpers on a document's bytes d[0..n) ------------------------------------------------------------------------
// does byte p belong to a whitespace code point (Python \s)?
__device__ __forceinline__ bool prep_is_ws(const uint8_t* d, int32_t p, int32_t n) {
    const uint32_t b = d[p];
    if (b <= 0x20) return (b >= 0x09 && b <= 0x0D) || b >= 0x1C;
    if (b < 0x80) return false;
    for (int back = 0; back < 3 && p - back >= 0; back++) {          // p may be the 1st, 2nd or 3rd byte of the sequence
        const uint32_t l = d[p - back];
        if (l >= 0xC2 && l <= 0xE3) { const int len = multibyte_ws(l, d, p - back, n); if (len > back) return true; }
        if ((l & 0xC0) != 0x80) break;                              // reached a lead byte that is not such a sequence
    }
    return false;
}
// start of the code point that holds byte p, and its value (malformed sequences give 0xFFFFFFFF)
__device__ __forceinline__ uint32_t prep_cp_at(const uint8_t* d, int32_t p, int32_t n, int32_t* start) {
    int32_t q = p;
    while (q > 0 && p - q < 3 && (d[q] & 0xC0) == 0x80) q--;
    *start = q;
    const uint32_t b = d[q];
    if (b < 0x80) return b;
    const int len = (b & 0xE0) == 0xC0 ? 2 : (b & 0xF0) == 0xE0 ? 3 : (b & 0xF8) == 0xF0 ? 4 : 0;
    if (len == 0 || q + len > n) return 0xFFFFFFFFu;
    uint32_t c = len == 2 ? (b & 0x1F) : len == 3 ? (b & 0x0F) : (b & 0x07);
    for (int k = 1; k < len; k++) { const uint32_t x = d[q + k]; if ((x & 0xC0) != 0x80) return 0xFFFFFFFFu; c = (c << 6) | (x & 0x3F); }
    return c;
}
// remove_emoji's character class (preprocess.py:51-69): its ranges add up to [U+24C2, U+10FFFF] plus four singles
__device__ __forceinline__ bool prep_is_emoji(uint32_t c) {
    return (c >= 0x24C2 && c <= 0x10FFFF) || c == 0x200D || c == 0x23CF || c == 0x23E9 || c == 0x231A;
}
// convert_unicode: if a base letter starts at q and a tone mark follows, the precomposed code point (else 0);
// *plen = bytes of base + mark
__device__ __forceinline__ uint32_t prep_compose_at(const uint8_t* d, int32_t q, int32_t n, int* plen) {
    if (q < 0 || q >= n) return 0;
    const uint32_t b = d[q];
    uint32_t base; int bl;
    if (b < 0x80) { base = b; bl = 1; }
    else if ((b == 0xC3 || b == 0xC4 || b == 0xC6) && q + 1 < n && (d[q + 1] & 0xC0) == 0x80) { base = ((b & 0x1F) << 6) | (d[q + 1] & 0x3F); bl = 2; }
    else return 0;
    if (q + bl + 1 >= n || d[q + bl] != 0xCC) return 0;
    const uint32_t m = d[q + bl + 1];
    const int mi = m == 0x80 ? 1 : m == 0x81 ? 2 : m == 0x83 ? 3 : m == 0x89 ? 4 : m == 0xA3 ? 5 : 0;
    if (!mi) return 0;
    for (int r = 0; r < 24; r++)
        if (PREP_COMPOSE[r][0] == base) { *plen = bl + 2; return PREP_COMPOSE[r][mi]; }
    return 0;
}

__device__ __forceinline__ int32_t last_bit_le(uint32_t mask, int lane) {   // highest set bit at or below lane, else -1
    const uint32_t m = mask & ((2u << lane) - 1);
    return m ? 31 - __clz(m) : -1;
}
__device__ __forceinline__ int32_t last_bit_lt(uint32_t mask, int lane) {   // highest set bit strictly below lane, else -1
    const uint32_t m = mask & ((1u << lane) - 1);
    return m ? 31 - __clz(m) : -1;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) k_prep(PrepArgs A) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < (uint64_t)A.n; r += nwarps) {
        const int64_t s64 = A.off[r];
        const int32_t n = (int32_t)(A.off[r + 1] - s64);
        const uint8_t* d = A.in + s64;
        uint8_t* o = WRITE ? A.out + A.out_off[r] : nullptr;
        int32_t lastgt_doc = -1;
        if (A.op == OP_HTML) {                                          // the last '>' of the document: no tag can open behind it
            for (int32_t base = 0; base < n; base += 32) {
                const int32_t p = base + lane;
                const uint32_t m = __ballot_sync(FULL_MASK, p < n && d[p] == '>');
                if (m) lastgt_doc = base + 31 - __clz(m);
            }
        }
        int32_t cA = -1, cB = -1;       // carried positions: HTML last '<' / last '>' ; URL last valid "http" / last ws ; EMOJI last kept / last ws
        int64_t run = 0;
        for (int32_t base = 0; base < n; base += 32) {
            const int32_t p = base + lane;
            const bool valid = p < n;
            const uint32_t b = valid ? d[p] : 0;
            uint32_t emit = 0, bytes = b;                                // bytes: up to 3 output bytes, low byte first
            if (A.op == OP_PUNCT) {
                const bool punct = (b >= 0x21 && b <= 0x2F) || (b >= 0x3A && b <= 0x40) || (b >= 0x5B && b <= 0x60) || (b >= 0x7B && b <= 0x7E);
                emit = valid && !punct;
            } else if (A.op == OP_HTML) {
                const uint32_t mlt = __ballot_sync(FULL_MASK, valid && b == '<'), mgt = __ballot_sync(FULL_MASK, valid && b == '>');
                const int32_t l1 = last_bit_le(mlt, lane), g1 = last_bit_lt(mgt, lane);
                const int32_t last_lt = l1 >= 0 ? base + l1 : cA, last_gt_before = g1 >= 0 ? base + g1 : cB;
                const bool del = p <= lastgt_doc && last_lt > last_gt_before;   // inside '<' ... first '>' behind it
                emit = valid && !del;
                if (mlt) cA = base + 31 - __clz(mlt);
                if (mgt) cB = base + 31 - __clz(mgt);
            } else if (A.op == OP_URL) {
                const bool ws = valid && prep_is_ws(d, p, n);
                const bool start = valid && p + 4 < n && b == 'h' && d[p + 1] == 't' && d[p + 2] == 't' && d[p + 3] == 'p' && !prep_is_ws(d, p + 4, n);
                const uint32_t ms = __ballot_sync(FULL_MASK, start), mw = __ballot_sync(FULL_MASK, ws);
                const int32_t s1 = last_bit_le(ms, lane), w1 = last_bit_le(mw, lane);
                const int32_t last_start = s1 >= 0 ? base + s1 : cA, last_ws = w1 >= 0 ? base + w1 : cB;
                emit = valid && !(last_start > last_ws);                 // "http" + at least one \S, up to the next whitespace
                if (ms) cA = base + 31 - __clz(ms);
                if (mw) cB = base + 31 - __clz(mw);
            } else if (A.op == OP_EMOJI) {
                int32_t q = p; uint32_t cp = 0x20;
                if (valid) cp = prep_cp_at(d, p, n, &q);
                const bool del = valid && prep_is_emoji(cp);
                const bool ws = valid && !del && prep_is_ws(d, p, n);
                const bool kept = valid && !del && !ws;
                const uint32_t mk = __ballot_sync(FULL_MASK, kept), mw = __ballot_sync(FULL_MASK, ws);
                const int32_t k1 = last_bit_lt(mk, lane), w1 = last_bit_lt(mw, lane);
                const int32_t last_kept = k1 >= 0 ? base + k1 : cA, last_ws = w1 >= 0 ? base + w1 : cB;
                // ' '.join(result.split()): a word that is not the first one is preceded by exactly one space
                const bool space = kept && q == p && last_kept >= 0 && last_ws > last_kept;
                emit = kept ? (space ? 2u : 1u) : 0u;
                bytes = space ? (0x20u | (b << 8)) : b;
                if (mk) cA = base + 31 - __clz(mk);
                if (mw) cB = base + 31 - __clz(mw);
            } else {   // OP_UNICODE
                emit = valid ? 1u : 0u;
                if (valid) {
                    int pl = 0;
                    const uint32_t c0 = prep_compose_at(d, p, n, &pl);
                    const bool cp_start = (b & 0xC0) != 0x80;
                    if (c0 && cp_start) {                                // I am the first byte of base + mark: emit the precomposed letter
                        if (c0 < 0x800) { emit = 2; bytes = (0xC0u | (c0 >> 6)) | ((0x80u | (c0 & 0x3F)) << 8); }
                        else { emit = 3; bytes = (0xE0u | (c0 >> 12)) | ((0x80u | ((c0 >> 6) & 0x3F)) << 8) | ((0x80u | (c0 & 0x3F)) << 16); }
                    } else {
                        for (int back = 1; back <= 3; back++) {          // am I a later byte of a composed pair that starts just before me?
                            const int32_t q = p - back;
                            if (q < 0) break;
                            if ((d[q] & 0xC0) == 0x80) continue;         // not a code point start
                            int pl2 = 0;
                            if (prep_compose_at(d, q, n, &pl2) && back < pl2) emit = 0;
                        }
                    }
                }
            }
            uint32_t incl = emit;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, incl, k); if (lane >= k) incl += t; }
            if (WRITE) {
                uint8_t* w = o + run + (incl - emit);
                for (uint32_t k = 0; k < emit; k++) w[k] = (uint8_t)(bytes >> (8 * k));
            }
            run += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (!WRITE && lane == 0) A.out_len[r] = run;
    }
}

}  // namespace gzt
