// encode.cuh -- the fused row kernel: UTF-8 whitespace pre-split, word-cache lookup, framing,
// truncation, padding, attention_mask and token types, one launch per chunk of documents.
//
// Replaces, per document (file:line in /root/reference/genz_tokenize/tokenize.py):
//   :106      re.findall(r"\S+\n?", text)                 -> classify16 / scan_side
//   :108-114  per-word bpe() + split                      -> word-cache hit (bpe.cuh fills misses)
//   :120-121  piece -> id with <unk> fallback             -> stored in the cache slot
//   :126-135  [bos] + ids + [eos]                         -> positions 0 / 1+nA
//   :222-246  pair framing  <s> A </s> </s> B </s>
//   :141-146  __padding (pad / truncate to max_len)
//   :148-152  get_atttention_mask
//   :154-182  get_sequence_id + get_token_type, :252-258 token_type_ids
//
// Work decomposition: a warp owns a tile of D = 32/G consecutive documents; G lanes walk one
// document in 16-byte pieces (one LDG.128 per lane), the D token rows are staged in shared memory
// and then written by the whole warp with 16-byte streaming stores.  Rows whose words are not all
// in the cache yet are queued for a second pass after k_bpe_pending has filled the new slots.
#pragma once
#include "device_common.cuh"

namespace gzt {

enum RowMode { MODE_FIXED = 0, MODE_COUNT = 1, MODE_RAGGED = 2 };

struct Side {
    const uint8_t* bytes;   // 16-byte aligned, readable up to round_up(nbytes, 16)
    const int64_t* off;     // [n+1]
    int64_t nbytes;
};

struct RowArgs {
    Side a, b;
    int32_t has_pair;
    int64_t n_rows;
    int32_t W;                 // FIXED: max_len
    uint32_t flags;
    // FIXED outputs
    int32_t* ids; uint8_t* mask; int8_t* tt; int8_t* seq;
    int32_t* row_len; int32_t* seq_len; uint8_t* status;
    // COUNT output / RAGGED input: framed length per row, and what to keep
    int32_t* L;                // [n_rows] framed length (COUNT writes)
    const int64_t* row_off;    // RAGGED: start of each row in ids
    const int32_t* keep;       // RAGGED: number of framed tokens to copy
    // second pass over a row list
    const uint32_t* row_list;  // NULL = all rows
    uint32_t* redo_list;       // rows with cache misses (filled in pass 1)
    uint32_t* fix_list;        // rows whose token types need the generic kernel
    int8_t eos_i8;
};

// ---- byte classification -----------------------------------------------------------------------
// bit i of the result = high bit of byte i of x (x has only 0x80 bits set)
__device__ __forceinline__ uint32_t gather_msb(uint32_t x) { return (((x >> 7) * 0x00204081u) >> 21) & 0xFu; }

// ASCII whitespace of Python \s: 0x09-0x0D, 0x1C-0x20, SWAR over 4 bytes -> 0x80 flags
__device__ __forceinline__ uint32_t ascii_ws4(uint32_t w) {
    uint32_t lo = w & 0x7F7F7F7Fu;
    uint32_t ge09 = lo + 0x77777777u;   // bit7 set iff (b&0x7f) >= 0x09
    uint32_t ge0e = lo + 0x72727272u;   // >= 0x0E
    uint32_t ge1c = lo + 0x64646464u;   // >= 0x1C
    uint32_t ge21 = lo + 0x5F5F5F5Fu;   // >= 0x21
    return ((ge09 & ~ge0e) | (ge1c & ~ge21)) & ~w & 0x80808080u;
}

__device__ __forceinline__ uint32_t byte_of(const uint4& w, int j) {
    uint32_t x = j < 8 ? (j < 4 ? w.x : w.y) : (j < 12 ? w.z : w.w);
    return (x >> ((j & 3) * 8)) & 0xFFu;
}

// Length (2 or 3) of the non-ASCII whitespace code point whose lead byte b0 sits at position p, else 0.
// The 19 non-ASCII members of \s: C2 85, C2 A0, E1 9A 80, E2 80 80..8A, E2 80 A8/A9/AF, E2 81 9F, E3 80 80.
__device__ __forceinline__ int multibyte_ws(uint32_t b0, const uint8_t* bytes, int64_t p, int64_t e) {
    if (b0 == 0xC2) {
        if (p + 1 < e) { uint32_t b1 = bytes[p + 1]; if (b1 == 0x85 || b1 == 0xA0) return 2; }
        return 0;
    }
    if (b0 < 0xE1 || b0 > 0xE3 || p + 2 >= e) return 0;
    uint32_t b1 = bytes[p + 1];
    if (b0 == 0xE1) return (b1 == 0x9A && bytes[p + 2] == 0x80) ? 3 : 0;
    if (b0 == 0xE3) return (b1 == 0x80 && bytes[p + 2] == 0x80) ? 3 : 0;
    if (b1 == 0x80) { uint32_t b2 = bytes[p + 2]; return ((b2 >= 0x80 && b2 <= 0x8A) || b2 == 0xA8 || b2 == 0xA9 || b2 == 0xAF) ? 3 : 0; }
    if (b1 == 0x81) return bytes[p + 2] == 0x9F ? 3 : 0;
    return 0;
}

// Whitespace bits of the 16 bytes at `a` (bits 0..15) plus spill into the next piece (bits 16,17).
// Bytes outside [s,e) count as whitespace.
__device__ __forceinline__ uint32_t classify16(const uint4& w, const uint8_t* bytes, int64_t a, int64_t s, int64_t e) {
    uint32_t inseg = 0xFFFFu;
    if (a < s) inseg &= 0xFFFFu << (int)(s - a);
    if (a + 16 > e) inseg &= 0xFFFFu >> (int)(a + 16 - e);
    uint32_t ws = gather_msb(ascii_ws4(w.x)) | (gather_msb(ascii_ws4(w.y)) << 4) | (gather_msb(ascii_ws4(w.z)) << 8) |
                  (gather_msb(ascii_ws4(w.w)) << 12);
    // lead bytes >= 0xC0: bit7 & bit6
    uint32_t lead = gather_msb(w.x & (w.x << 1) & 0x80808080u) | (gather_msb(w.y & (w.y << 1) & 0x80808080u) << 4) |
                    (gather_msb(w.z & (w.z << 1) & 0x80808080u) << 8) | (gather_msb(w.w & (w.w << 1) & 0x80808080u) << 12);
    lead &= inseg;
    while (lead) {
        int j = __ffs(lead) - 1;
        lead &= lead - 1;
        uint32_t b0 = byte_of(w, j);
        if (b0 != 0xC2 && (b0 < 0xE1 || b0 > 0xE3)) continue;
        int l = multibyte_ws(b0, bytes, a + j, e);
        if (l) ws |= ((1u << l) - 1) << j;
    }
    return (ws & 0x3FFFFu) | (~inseg & 0xFFFFu);
}

// First whitespace position at or after q (q is on a code point boundary or inside a non-ws one).
__device__ __noinline__ int64_t slow_word_end(const uint8_t* bytes, int64_t q, int64_t e) {
    while (q < e) {
        uint32_t b = bytes[q];
        if (b <= 0x20) { if ((b >= 0x09 && b <= 0x0D) || b >= 0x1C) return q; }
        else if (b >= 0xC2 && multibyte_ws(b, bytes, q, e)) return q;
        q++;
    }
    return e;
}

// ---- word -> cache slot --------------------------------------------------------------------------
// 16-byte zero-padded key of the word at [p, p+len), len <= 16, via aligned 16-byte loads.
__device__ __forceinline__ void load_key16(const uint8_t* bytes, int64_t p, uint32_t len, uint64_t* k0, uint64_t* k1) {
    int64_t A = p & ~(int64_t)15;
    int sh = (int)(p - A);
    uint4 lo = ldg128(bytes + A);
    uint64_t q0 = ((uint64_t)lo.y << 32) | lo.x, q1 = ((uint64_t)lo.w << 32) | lo.z, q2 = 0, q3 = 0;
    if (sh + (int)len > 16) {
        uint4 hi = ldg128(bytes + A + 16);
        q2 = ((uint64_t)hi.y << 32) | hi.x;
        q3 = ((uint64_t)hi.w << 32) | hi.z;
    }
    if (sh >= 8) { q0 = q1; q1 = q2; q2 = q3; sh -= 8; }
    uint64_t r0 = q0, r1 = q1;
    if (sh) {
        int s8 = sh * 8;
        r0 = (q0 >> s8) | (q1 << (64 - s8));
        r1 = (q1 >> s8) | (q2 << (64 - s8));
    }
    if (len < 8) { r0 &= (1ULL << (len * 8)) - 1; r1 = 0; }
    else if (len < 16) { r1 &= (1ULL << ((len - 8) * 8)) - 1; }
    *k0 = r0; *k1 = r1;
}

// Find the word in the cache or insert it (BPE pending).  Returns the slot index.
__device__ __forceinline__ uint32_t cache_find_or_insert(const WordCache& C, const uint8_t* wptr, uint32_t len, uint64_t k0, uint64_t k1,
                                                         uint32_t h) {
    uint32_t idx = h & C.mask;
    bool fresh = false;   // false: first look at a slot may come from L1
    for (uint32_t guard = 0;; guard++) {
        if (guard > (1u << 24)) { atomicAdd(&C.ctr[C_ERR], 1ULL); return idx; }   // never spin forever: report instead
        Slot* s = &C.slots[idx];
        uint4 a = fresh ? ld_cg128(s) : *reinterpret_cast<const uint4*>(s);
        if (a.x == len) {
            uint4 b = fresh ? ld_cg128(reinterpret_cast<const uint4*>(s) + 1) : *(reinterpret_cast<const uint4*>(s) + 1);
            uint64_t s0 = ((uint64_t)b.y << 32) | b.x, s1 = ((uint64_t)b.w << 32) | b.z;
            bool eq;
            if (len <= 16) eq = (s0 == k0) && (s1 == k1);
            else {
                eq = s1 == k1;
                if (eq) {
                    const uint8_t* kp = C.key_arena + s0;
                    for (uint32_t i = 0; i < len && eq; i++) eq = kp[i] == wptr[i];
                }
            }
            if (eq) return idx;
        } else if (a.x == SLOT_EMPTY || a.x == SLOT_LOCKED) {
            if (!fresh) { fresh = true; continue; }          // L1 may be stale: look again in L2
            if (a.x == SLOT_LOCKED) continue;                // another thread is writing this slot
            uint32_t old = atomicCAS(&s->len, SLOT_EMPTY, SLOT_LOCKED);
            if (old != SLOT_EMPTY) continue;                 // lost the race: re-examine the same slot
            uint64_t v0 = k0, v1 = k1;
            if (len > 16) {
                uint64_t off = atomicAdd(&C.ctr[C_KEYS], (unsigned long long)len);
                uint8_t* kp = C.key_arena + off;
                for (uint32_t i = 0; i < len; i++) kp[i] = wptr[i];
                v0 = off;
            }
            s->ntok = 0; s->t0 = 0; s->t1 = 0; s->k0 = v0; s->k1 = v1;
            __threadfence();
            st_release_u32(&s->len, len);
            atomicAdd(&C.ctr[C_SLOTS], 1ULL);
            uint64_t pi = atomicAdd(&C.ctr[C_PENDING], 1ULL);
            if (pi < C.pending_cap) C.pending[pi] = idx;
            else atomicAdd(&C.ctr[C_ERR], 1ULL);
            return idx;
        }
        idx = (idx + 1) & C.mask;
        fresh = false;
    }
}

// ---- token types for a pair row in closed form ---------------------------------------------------
// get_sequence_id + get_token_type (tokenize.py:154-182) evaluated on the known positions of
// </s> in a row without interior special ids.  Generic rows go to k_post_rows.
struct SeqDesc { int32_t p1, m, f1, f2, r1, r2, err; };

__device__ __forceinline__ SeqDesc seq_describe(int32_t nA, int32_t L, int32_t W) {
    // eos positions of the final row T (length W): framed a, a+1, c=L-1 if they survive truncation, W-1 if truncated
    int32_t E[4]; int ne = 0;
    const int32_t a = 1 + nA, c = L - 1;
    const bool trunc = L >= W;
    if (a <= W - 2) E[ne++] = a;
    if (a + 1 <= W - 2) E[ne++] = a + 1;
    if (c <= W - 2 && c > a + 1) E[ne++] = c;
    if (trunc) E[ne++] = W - 1;
    SeqDesc d;
    d.p1 = E[0];
    int32_t e = -1;
    for (int k = 1; k < ne; k++)
        if (E[k] >= d.p1 + 2 && E[k - 1] != E[k] - 1) { e = E[k]; break; }
    d.m = e >= 0 ? e + 1 : W;
    int32_t N[4]; int nn = 0;   // None positions after S[0]=0, S[m-1]=1
    for (int k = 0; k < ne; k++)
        if (E[k] < d.m && E[k] != 0 && E[k] != d.m - 1) N[nn++] = E[k];
    d.f1 = nn > 0 ? N[0] : -1; d.f2 = nn > 1 ? N[1] : -1;
    d.r1 = nn > 2 ? N[2] : -1; d.r2 = nn > 3 ? N[3] : -1;
    d.err = nn < 2;
    return d;
}
__device__ __forceinline__ int32_t seq_value(const SeqDesc& d, int32_t i) {
    int32_t v = i < d.p1 ? 0 : 1;
    if (i == d.r1 || i == d.r2) v = -1;
    if (i == d.f1) v = 0;
    if (i == d.f2) v = 1;
    if (i == 0) v = 0;
    if (i == d.m - 1) v = 1;
    return v;
}

// ---- the row kernel --------------------------------------------------------------------------------
template <int G, int MODE>
struct RowKernel {
    static constexpr int D = 32 / G;
    static constexpr int MAXW = 8;   // word starts per 16-byte piece

    // Walk one side of the tile's documents.  All 32 lanes call this; lanes [g*G, g*G+G) own document g.
    // pos (group-uniform) is the next token position of the row; tokens at positions < limit are delivered.
    __device__ static __forceinline__ void scan_side(const DevTables& T, const WordCache& C, const Side& sd, int64_t s, int64_t e, bool active,
                                                    int gl, uint32_t* scratch /*[MAXW][32] for this warp*/, int lane, int32_t& pos, int32_t limit,
                                                    uint32_t& flags, int32_t* rowbuf, int32_t* gout, int32_t glimit) {
        const uint8_t* bytes = sd.bytes;
        const int64_t base = s & ~(int64_t)15;
        int32_t n_it = (active && e > s) ? (int32_t)((e - base + 16 * G - 1) / (16 * G)) : 0;
        const int32_t max_it = __reduce_max_sync(FULL_MASK, n_it);
        uint32_t carry = 1u << 15;    // the byte before the document is whitespace, nothing spills in
        for (int32_t it = 0; it < max_it; ++it) {
            const bool g_on = it < n_it && pos < limit;
            const int64_t a = base + ((int64_t)it * G + gl) * 16;
            const bool l_on = g_on && a < e;
            uint32_t ws18 = 0xFFFFu;
            if (l_on) {
                uint4 w = ldg128(bytes + a);
                ws18 = classify16(w, bytes, a, s, e);
            }
            uint32_t prev = __shfl_up_sync(FULL_MASK, ws18, 1, G);
            if (gl == 0) prev = carry;
            carry = __shfl_sync(FULL_MASK, ws18, G - 1, G);
            const uint32_t nw = ~(ws18 | (prev >> 16)) & 0xFFFFu;
            uint32_t st = nw & ~((nw << 1) | ((~prev >> 15) & 1u));
            // non-whitespace bits of the following pieces, to find word ends without touching memory
            uint64_t win = nw;
            {
                uint32_t n1 = __shfl_down_sync(FULL_MASK, nw, 1, G), n2 = __shfl_down_sync(FULL_MASK, nw, 2, G),
                         n3 = __shfl_down_sync(FULL_MASK, nw, 3, G);
                if (gl + 1 < G) win |= (uint64_t)n1 << 16;
                if (gl + 2 < G) win |= (uint64_t)n2 << 32;
                if (gl + 3 < G) win |= (uint64_t)n3 << 48;
            }
            const int known = 16 * ((G - gl) < 4 ? (G - gl) : 4);
            // ---- my word starts: find end, look up
            int32_t cnt = 0; int nwords = 0;
            if (!l_on) st = 0;
            while (st) {
                const int b = __ffs(st) - 1;
                st &= st - 1;
                const int64_t p = a + b;
                uint64_t z = (~win) >> (b + 1);
                int run = z ? __ffsll((long long)z) : 65;     // bytes after p up to the first whitespace
                int64_t end;
                if (b + run < known) end = p + run;
                else end = slow_word_end(bytes, a + known, e);
                if (end < e && bytes[end] == 0x0A) end++;          // \S+\n?  (tokenize.py:106)
                const uint32_t len = (uint32_t)(end - p);
                uint64_t k0, k1; uint32_t h;
                if (len <= 16) { load_key16(bytes, p, len, &k0, &k1); h = hash_key16(k0, k1, len); }
                else { k1 = hash_long(bytes + p, len); k0 = 0; h = fmix32((uint32_t)k1 ^ (uint32_t)(k1 >> 32)); }
                const uint32_t si = cache_find_or_insert(C, bytes + p, len, k0, k1, h);
                const Slot* sl = &C.slots[si];
                const uint32_t nt = sl->ntok;
                uint32_t enc = si;
                if (nt == 1) enc = 0x80000000u | sl->t0;
                else if (nt == 0) flags |= 1u;                      // BPE pending: the row needs the second pass
                cnt += (int32_t)nt;
                if (nwords < MAXW) scratch[nwords * 32 + lane] = enc;
                nwords++;
            }
            // ---- positions: exclusive scan of token counts over the group
            int32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                int32_t t = __shfl_up_sync(FULL_MASK, incl, o, G);
                if (gl >= o) incl += t;
            }
            const int32_t total = __shfl_sync(FULL_MASK, incl, G - 1, G);
            if (MODE != MODE_COUNT) {
                int32_t q = pos + incl - cnt;
                for (int j = 0; j < nwords; j++) {
                    const uint32_t enc = scratch[j * 32 + lane];
                    if (enc & 0x80000000u) {
                        const int32_t t = (int32_t)(enc & 0x7FFFFFFFu);
                        if (t == T.eos || t == T.bos) flags |= 2u;
                        if (MODE == MODE_FIXED) { if (q < limit) rowbuf[q] = t; }
                        else if (q < glimit) gout[q] = t;
                        q++;
                    } else {
                        const Slot* sl = &C.slots[enc];
                        const uint32_t nt = sl->ntok;
                        const uint32_t t0 = sl->t0, t1 = sl->t1;
                        for (uint32_t k = 0; k < nt; k++) {
                            const int32_t t = (int32_t)(nt <= 2 ? (k == 0 ? t0 : t1) : C.tok_arena[t0 + k]);
                            if (t == T.eos || t == T.bos) flags |= 2u;
                            if (MODE == MODE_FIXED) { if (q < limit) rowbuf[q] = t; }
                            else if (q < glimit) gout[q] = t;
                            q++;
                        }
                    }
                }
            }
            pos += total;
        }
    }
};

template <int G, int MODE>
__global__ void __launch_bounds__(256) k_rows(DevTables T, WordCache C, RowArgs A) {
    using K = RowKernel<G, MODE>;
    constexpr int D = K::D;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gl = lane % G, g = lane / G;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)wib * K::MAXW * 32;
    int32_t* rowbufs = reinterpret_cast<int32_t*>(smem_raw + (size_t)wpb * K::MAXW * 32 * 4);
    const int32_t W = A.W;
    const int32_t Wp = (W + 3) & ~3;
    int32_t* rowbuf = MODE == MODE_FIXED ? rowbufs + ((size_t)wib * D + g) * Wp : nullptr;

    const unsigned long long n_items = A.row_list ? C.ctr[C_REDO] : (unsigned long long)A.n_rows;
    const unsigned long long n_tiles = (n_items + D - 1) / D;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * wpb;
    unsigned long long tok_total = 0;
    for (unsigned long long tile = (unsigned long long)blockIdx.x * wpb + wib; tile < n_tiles; tile += n_warps) {
        const unsigned long long item = tile * D + g;
        const bool active = item < n_items;
        const int64_t r = active ? (A.row_list ? (int64_t)A.row_list[item] : (int64_t)item) : 0;
        int64_t sa = 0, ea = 0, sb = 0, eb = 0;
        if (active) {
            sa = A.a.off[r]; ea = A.a.off[r + 1];
            if (A.has_pair) { sb = A.b.off[r]; eb = A.b.off[r + 1]; }
        }
        // where the row's tokens go
        int32_t limit, glimit = 0; int32_t* gout = nullptr;
        if (MODE == MODE_FIXED) limit = W - 1;
        else if (MODE == MODE_COUNT) limit = 0x7FFFFFFF;
        else { limit = 0x7FFFFFFF; if (active) { gout = A.ids + A.row_off[r]; glimit = A.keep[r]; } }

        uint32_t flags = 0;
        int32_t pos = 1;                                           // position 0 is <s> (tokenize.py:135)
        if (MODE == MODE_FIXED) { if (gl == 0 && active && limit > 0) rowbuf[0] = T.bos; }
        else if (MODE == MODE_RAGGED) { if (gl == 0 && active && glimit > 0) gout[0] = T.bos; }
        __syncwarp();
        K::scan_side(T, C, A.a, sa, ea, active, gl, scratch, lane, pos, limit, flags, rowbuf, gout, glimit);
        const int32_t nA = pos - 1;
        if (A.has_pair) {
            // ... </s> </s> B   (tokenize.py:237-239)
            if (gl == 0 && active) {
                if (MODE == MODE_FIXED) { if (pos < limit) rowbuf[pos] = T.eos; if (pos + 1 < limit) rowbuf[pos + 1] = T.eos; }
                else if (MODE == MODE_RAGGED) { if (pos < glimit) gout[pos] = T.eos; if (pos + 1 < glimit) gout[pos + 1] = T.eos; }
            }
            pos += 2;
            K::scan_side(T, C, A.b, sb, eb, active, gl, scratch, lane, pos, limit, flags, rowbuf, gout, glimit);
        }
        if (gl == 0 && active) {                                   // closing </s>
            if (MODE == MODE_FIXED) { if (pos < limit) rowbuf[pos] = T.eos; }
            else if (MODE == MODE_RAGGED) { if (pos < glimit) gout[pos] = T.eos; }
        }
        const int32_t L = pos + 1;                                 // framed length (>= W when the walk stopped early)
        // merge per-lane flags over the group
#pragma unroll
        for (int o = 1; o < G; o <<= 1) flags |= __shfl_xor_sync(FULL_MASK, flags, o, G);
        const bool dirty = flags & 1u;
        if (active && gl == 0 && dirty) {
            if (!A.row_list) {
                unsigned long long k = atomicAdd(&C.ctr[C_REDO], 1ULL);
                A.redo_list[k] = (uint32_t)r;
            } else {
                atomicAdd(&C.ctr[C_ERR], 1ULL);   // cannot happen: every word of a redo row was inserted in pass 1
            }
        }
        if (MODE == MODE_COUNT) {
            if (active && gl == 0) A.L[r] = L;
            continue;
        }
        if (MODE == MODE_RAGGED) continue;
        __syncwarp();
        // ---- FIXED: the warp writes its D rows --------------------------------------------------------
        const bool need_fix = A.has_pair && ((flags & 2u) || !T.specials_distinct);
        for (int d = 0; d < D; d++) {
            const int src = d * G;
            const bool d_active = __shfl_sync(FULL_MASK, (int)active, src);
            const bool d_dirty = __shfl_sync(FULL_MASK, (int)dirty, src);
            if (!d_active || d_dirty) continue;
            const int64_t dr = __shfl_sync(FULL_MASK, (long long)r, src);
            const int32_t dL = __shfl_sync(FULL_MASK, L, src);
            const int32_t dnA = __shfl_sync(FULL_MASK, nA, src);
            const bool d_fix = __shfl_sync(FULL_MASK, (int)need_fix, src);
            const int32_t Lr = dL < W ? dL : W;
            const bool trunc = dL >= W;
            const int32_t* rb = rowbufs + ((size_t)wib * D + d) * Wp;
            SeqDesc sd;
            if (A.has_pair) sd = seq_describe(dnA, dL, W);
            int32_t ntok_row = 0;
            if ((W & 15) == 0) {
                for (int32_t i0 = lane * 4; i0 < W; i0 += 128) {
                    int4 v = *reinterpret_cast<const int4*>(rb + i0);
                    int32_t x[4] = {v.x, v.y, v.z, v.w};
                    uint32_t mk = 0, ttw = 0, sqw = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int32_t i = i0 + k;
                        int32_t t = i < Lr ? ((trunc && i == W - 1) ? T.eos : x[k]) : T.pad;
                        x[k] = t;
                        const uint32_t on = t != T.pad;
                        mk |= on << (8 * k);
                        ntok_row += (int32_t)on;
                        if (A.has_pair) {
                            int32_t sv = i < sd.m ? seq_value(sd, i) : 0;
                            int32_t tv = (sd.m == W && i == W - 1) ? (int32_t)A.eos_i8 : sv;
                            ttw |= ((uint32_t)tv & 0xFFu) << (8 * k);
                            sqw |= ((uint32_t)(i < sd.m ? sv : -2) & 0xFFu) << (8 * k);
                        }
                    }
                    st_cs128(A.ids + dr * W + i0, make_uint4((uint32_t)x[0], (uint32_t)x[1], (uint32_t)x[2], (uint32_t)x[3]));
                    if (A.mask) st_cs32(A.mask + dr * W + i0, mk);
                    if (A.has_pair && A.tt) st_cs32(A.tt + dr * W + i0, ttw);
                    if (A.has_pair && A.seq) st_cs32(A.seq + dr * W + i0, sqw);
                }
            } else {
                for (int32_t i = lane; i < W; i += 32) {
                    int32_t t = i < Lr ? ((trunc && i == W - 1) ? T.eos : rb[i]) : T.pad;
                    A.ids[dr * W + i] = t;
                    const uint32_t on = t != T.pad;
                    ntok_row += (int32_t)on;
                    if (A.mask) A.mask[dr * W + i] = (uint8_t)on;
                    if (A.has_pair) {
                        int32_t sv = i < sd.m ? seq_value(sd, i) : 0;
                        int32_t tv = (sd.m == W && i == W - 1) ? (int32_t)A.eos_i8 : sv;
                        if (A.tt) A.tt[dr * W + i] = (int8_t)tv;
                        if (A.seq) A.seq[dr * W + i] = (int8_t)(i < sd.m ? sv : -2);
                    }
                }
            }
            ntok_row = __reduce_add_sync(FULL_MASK, ntok_row);
            if (lane == 0) {
                if (A.row_len) A.row_len[dr] = Lr;
                if (A.has_pair) {
                    if (A.seq_len) A.seq_len[dr] = sd.m;
                    if (A.status) A.status[dr] = (uint8_t)sd.err;
                    if (d_fix) { unsigned long long k = atomicAdd(&C.ctr[C_FIX], 1ULL); A.fix_list[k] = (uint32_t)dr; }
                }
                tok_total += (unsigned long long)ntok_row;
            }
        }
        __syncwarp();
    }
    if (MODE == MODE_FIXED && lane == 0 && tok_total) atomicAdd(&C.ctr[C_TOKENS], tok_total);
}

}  // namespace gzt
