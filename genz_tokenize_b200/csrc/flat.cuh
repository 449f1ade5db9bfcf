// flat.cuh -- the byte-parallel encode pipeline for the fixed layout.
//
// The fused row kernel (encode.cuh) gives every warp a tile of documents; its lanes idle whenever documents, words
// and 16-byte pieces do not line up, and each tile is one long dependent chain.  Here the work is cut by BYTES:
//
//   k_flat_doc_starts   one bit per byte position where a document starts (the only place documents matter before
//                       the rows are assembled: a word neither continues into nor a whitespace straddles a new document)
//   k_flat_words        a block owns 8 KiB of the packed text, whatever documents it holds: every thread classifies 32
//                       bytes (tokenize.py:106  \S+\n?), the block compacts its word starts, then looks the words up in
//                       the word cache one word per thread (tokenize.py:108-121; misses are inserted for k_bpe_pending)
//                       and stores the cache value of every word at  block * 8192 + (index of the word in the block)
//   k_flat_fix          after k_bpe_pending: the values of the words that were pending
//   k_flat_rows         a warp owns D rows: word ranges from the start bitmap (rank of the document offsets), framing,
//                       truncation, padding, mask, token types (tokenize.py:126-182,222-258) staged in final form and
//                       written by the TMA unit (TmaPlanes); rows it cannot finish go to the generic second pass
//
// Replaces the same reference lines as encode.cuh; outputs are bit-identical to the fused kernel's.
#pragma once
#include "encode.cuh"

namespace gzt {

static const int FC_OWN = 30;                  // 32-byte granules of the text a warp of k_flat_words owns per chunk (30 of the 32 it classifies)
static const int FC_BYTES = FC_OWN * 32;       // 960 text bytes per chunk
static const int FW_WARPS = 8;                 // warps per block of k_flat_words (each warp is on its own)
static const int FC_SHIFT = 10;                // wtok holds up to 1024 words per chunk (>= 960, one per byte at most)

struct FlatSide {
    const uint8_t* bytes;     // Side.bytes (absolute offsets index it)
    const int64_t* off;       // [n+1]
    int64_t n;
    uint32_t* dsb;            // [nB*30 + 2] document-start bits (bit q = position P0 + q, P0 = off[0] & ~15)
    uint32_t* st;             // [nB*30 + 2] word-start bits per 32-byte granule
    uint16_t* tpref;          // [nB*30 + 2] words of the chunk before the granule
    uint32_t* cnt;            // [nB]        words per chunk
    uint32_t* wtok;           // [nB*1024]   cache value (VAL_*) per word
    uint32_t* fixa;           // pending words: index into wtok ...
    uint32_t* fixp;           // ... and position (relative to P0)
    uint32_t fix_cap;
    int ctr_fix;              // counter index in WordCache::ctr
    uint32_t nB;              // chunks of FC_BYTES that cover the text (from the caller's byte count)
};
// end of the text relative to P0, never beyond what the work arrays cover (a caller that under-reports the byte
// count gets an error from k_flat_doc_starts instead of an out-of-bounds access)
__device__ __forceinline__ uint32_t flat_hi(const FlatSide& S, int64_t P0) {
    const int64_t hi = S.off[S.n] - P0, capq = (int64_t)S.nB * FC_BYTES - 64;
    return (uint32_t)(hi < capq ? hi : capq);
}

// ---- helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ascii_ws_byte(uint32_t b) { return (b >= 0x09 && b <= 0x0D) || (b >= 0x1C && b <= 0x20); }
__device__ __forceinline__ bool dsb_bit(const uint32_t* dsb, uint32_t q) { return (dsb[q >> 5] >> (q & 31)) & 1u; }
// length of the multi-byte whitespace code point starting at q (0 = none): inside [.., hi) and not across a document start
__device__ __noinline__ int flat_mb_ws(const uint8_t* tb, const uint32_t* dsb, uint32_t q, uint32_t hi) {
    const uint32_t b0 = tb[q];
    if (b0 < 0xC2) return 0;
    const int l = multibyte_ws(b0, tb, (int32_t)q, (int32_t)hi);
    if (!l) return 0;
    for (int k = 1; k < l; k++) if (dsb_bit(dsb, q + k)) return 0;
    return l;
}
// is byte q (lo <= q < hi) part of a whitespace code point?
__device__ __noinline__ bool flat_is_ws(const uint8_t* tb, const uint32_t* dsb, uint32_t q, uint32_t lo, uint32_t hi) {
    const uint32_t b = tb[q];
    if (b < 0x80) return ascii_ws_byte(b);
    for (uint32_t back = 0; back < 3 && q >= lo + back; back++) {
        const int l = flat_mb_ws(tb, dsb, q - back, hi);
        if (l > (int)back) return true;
    }
    return false;
}
// length of the word (\S+\n?) that starts at q: up to whitespace, a document start or the end of the text
__device__ __noinline__ uint32_t flat_word_len(const uint8_t* tb, const uint32_t* dsb, uint32_t q, uint32_t hi) {
    uint32_t e = q + 1;
    while (e < hi) {
        if (dsb_bit(dsb, e)) return e - q;
        const uint32_t b = tb[e];
        if (b <= 0x20) { if (ascii_ws_byte(b)) break; }
        else if (b >= 0xC2 && flat_mb_ws(tb, dsb, e, hi)) break;
        e++;
    }
    if (e < hi && tb[e] == 0x0A && !dsb_bit(dsb, e)) e++;
    return e - q;
}
// lookup with a known length (the tail of lookup_slow)
__device__ __noinline__ uint32_t lookup_len(const WordCache& C, const uint8_t* wptr_base, uint32_t q, uint32_t len, int insert_ok) {
    uint64_t k0, k1, k2; uint32_t h;
    if (len <= KEY_INLINE) { load_key24(wptr_base, (int32_t)q, len, &k0, &k1, &k2); h = hash_key24(k0, k1, k2, len); }
    else { k1 = hash_long(wptr_base + q, len); k0 = 0; k2 = 0; h = fmix32((uint32_t)k1 ^ (uint32_t)(k1 >> 32)); }
    return cache_find_or_insert(C, wptr_base + q, len, k0, k1, k2, h, insert_ok != 0);
}

// Terminator and newline bits of the granule behind a block, so that the block's last words end inside known bits:
// only its plain-ASCII case (a high byte there leaves the bits unknown and such a word is measured bytewise).
__device__ __noinline__ void flat_lookahead(const uint8_t* tb, const uint32_t* dsb, uint32_t qn, uint32_t hi, uint32_t* term, uint32_t* nl) {
    uint32_t tn = 0u, nn = 0u;
    if (qn + 32 <= hi) {
        const uint4 a = ldg128(tb + qn), b = ldg128(tb + qn + 16);
        const uint32_t x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t wsn = 0, hn = 0, nln = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            wsn |= gather_msb(ascii_ws4(x[k])) << (4 * k);
            hn |= x[k];
            const uint32_t y = x[k] ^ 0x0A0A0A0Au;
            nln |= gather_msb(~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u) << (4 * k);
        }
        if (!(hn & 0x80808080u)) { const uint32_t dn = dsb[qn >> 5]; tn = wsn | dn; nn = nln & ~dn; }
    }
    *term = tn; *nl = nn;
}

// ---- document starts ---------------------------------------------------------------------------------------
__global__ void k_flat_doc_starts(WordCache C, FlatSide S) {
    pdl_wait(); pdl_trigger();
    const int64_t o0 = S.off[0];
    const int64_t P0 = o0 & ~(int64_t)15;
    const int64_t capq = (int64_t)S.nB * FC_BYTES - 64;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d <= S.n; d += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = S.off[d] - P0;
        if (q < 0 || q > capq || (d > 0 && S.off[d] < S.off[d - 1])) { atomicAdd(&C.ctr[C_ERR], 1ULL); continue; }   // offsets outside the stated text
        if (d < S.n) atomicOr(&S.dsb[q >> 5], 1u << (q & 31));
    }
}

__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* m, uint32_t smem_addr, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_addr), "r"(c0), "r"(c1) : "memory");
}
// the same with an L2 eviction policy: the planes are written once and never read here, they must not push the word
// arrays (read right after they were written) out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// The slots of the frequent words are read again and again while hundreds of MB of planes stream through L2: the probes ask
// L2 to evict those lines last, so that a batch of probes does not wait for the one that had to go to DRAM.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld_keep128(const uint4* p, uint64_t policy) {
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, uint32_t smem_addr, int32_t c0, int32_t c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(m), "r"(smem_addr), "r"(c0), "r"(c1), "l"(policy)
                 : "memory");
}

// The pad columns KR..W-1 of the output planes depend on nothing: k_flat_words writes them on the side (TMA tensor stores
// from a block-wide constant buffer, issued by one lane per chunk), while its warps wait for their probes.
struct PadJob {
    int32_t on;          // 0: nothing to do
    int32_t n_tiles;     // tiles of D rows this launch writes ...
    int32_t tile0;       // ... starting with this one
    int32_t W, D, KR, PB;
    int32_t want_tt;
    int32_t pad_id;
};

// tiles [c * n_tiles / n_chunks, (c + 1) * n_tiles / n_chunks) of the pad columns; one thread.  Out of line: k_flat_words
// is short of registers.
__device__ __noinline__ void flat_pad_tiles(const PadJob& J, const TmaPlanes& M, uint32_t c, uint32_t n_chunks, uint32_t pad_s, uint32_t zeros_s, uint64_t l2_first) {
    const uint32_t t0 = (uint32_t)((uint64_t)c * (uint32_t)J.n_tiles / n_chunks), t1 = (uint32_t)(((uint64_t)c + 1) * (uint32_t)J.n_tiles / n_chunks);
    for (uint32_t t = t0; t < t1; t++) {
        const int32_t r = (int32_t)((t + (uint32_t)J.tile0) * (uint32_t)J.D);
        for (int32_t c0 = J.KR; c0 < J.W; c0 += J.PB) {
            tma_store_2d_hint(&M.ids_pad, pad_s, c0, r, l2_first);
            tma_store_2d_hint(&M.mask_pad, zeros_s, c0, r, l2_first);
            if (J.want_tt) tma_store_2d_hint(&M.tt_pad, zeros_s, c0, r, l2_first);
        }
    }
    if (t1 > t0) bulk_commit();
}

// ---- words -------------------------------------------------------------------------------------------------
// One warp per chunk, no block-wide synchronisation: lane l classifies granule 30c - 1 + l of chunk c, so lanes 1..30 own
// their granules' words while lane 0 (the granule before) and lane 31 (the granule after) only supply what the neighbours
// need: whether the byte before a granule belongs to a word, whitespace spilling over, where a word of the last owned
// granule ends.  Warps walk the chunks with a grid stride and load the next chunk's text while they look up this one's words.
struct FlatWarpSmem {
    uint8_t text[1024 + 16];       // the 32 classified granules and 16 bytes more (key gathers)
    uint16_t wl[1024];             // words: position in the classified window | length << 10
};
template <int MINB, int ILP>
__global__ void __launch_bounds__(FW_WARPS * 32, MINB * (8 / FW_WARPS)) k_flat_words(DevTables T, WordCache C, FlatSide S, int insert_ok, PadJob J, const __grid_constant__ TmaPlanes M) {
    pdl_wait(); pdl_trigger();
    __shared__ __align__(16) FlatWarpSmem s_warp[FW_WARPS];
    extern __shared__ __align__(1024) uint8_t pad_smem[];                   // J.on: [D x PB] pad ids, [D x PB] zero bytes
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    FlatWarpSmem& sm = s_warp[wib];
    uint32_t pad_s = 0, zeros_s = 0; uint64_t l2_first = 0;
    const uint64_t l2_last = l2_policy_evict_last();
    if (J.on) {
        const uint4 pad4 = make_uint4((uint32_t)J.pad_id, (uint32_t)J.pad_id, (uint32_t)J.pad_id, (uint32_t)J.pad_id);
        uint4* cp = reinterpret_cast<uint4*>(pad_smem);
        const int n_pad = (int)(r128((size_t)J.D * J.PB * 4) >> 4), n_all = (int)(tma_const_bytes(J.D, J.PB) >> 4);
        for (int i = tid; i < n_all; i += blockDim.x) cp[i] = i < n_pad ? pad4 : make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        __syncthreads();
        pad_s = (uint32_t)__cvta_generic_to_shared(pad_smem);
        zeros_s = pad_s + (uint32_t)r128((size_t)J.D * J.PB * 4);
        l2_first = l2_policy_evict_first();
    }
    const int64_t o0 = S.off[0];
    const int64_t P0 = o0 & ~(int64_t)15;
    const uint8_t* __restrict__ tb = S.bytes + P0;
    const uint32_t lo = (uint32_t)(o0 - P0), hi = flat_hi(S, P0);   // the text is [lo, hi)
    const uint32_t n_warps = gridDim.x * (uint32_t)FW_WARPS;
    // granule of this lane in chunk c is 30c - 1 + lane; its bytes start at q0 (may be "negative" for chunk 0, lane 0)
    const uint4 sp = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
    uint4 nv0 = sp, nv1 = sp, nv2 = sp; uint32_t ndsb = 0;
    auto load_chunk = [&](uint32_t c) {
        nv0 = sp; nv1 = sp; nv2 = sp; ndsb = 0;
        if (c >= S.nB) return;
        const int64_t gq = ((int64_t)c * FC_OWN - 1 + lane) * 32;
        if (gq >= 0) {
            const uint32_t q0 = (uint32_t)gq;
            if (q0 < hi) nv0 = ldg128(tb + q0);
            if (q0 + 16 < hi) nv1 = ldg128(tb + q0 + 16);
            if (lane == 31 && q0 + 32 < hi) nv2 = ldg128(tb + q0 + 32);
            ndsb = S.dsb[q0 >> 5];
        }
    };
    uint32_t c = blockIdx.x * (uint32_t)FW_WARPS + wib;
    load_chunk(c);
    for (; c < S.nB; c += n_warps) {
        const uint4 v0 = nv0, v1 = nv1, v2 = nv2;
        const uint32_t dsbw = ndsb;
        const int64_t gq = ((int64_t)c * FC_OWN - 1 + lane) * 32;
        const bool exists = gq >= 0;
        const uint32_t q0 = exists ? (uint32_t)gq : 0u;
        const uint32_t wq = c * (uint32_t)FC_BYTES - 32u;                     // position of the window's byte 0 (wraps for chunk 0: only used + p >= 32)
        __syncwarp();                                                         // the previous chunk's lookups have read the staging area
        uint8_t* const text = sm.text;
        reinterpret_cast<uint4*>(text)[lane * 2] = v0;
        reinterpret_cast<uint4*>(text)[lane * 2 + 1] = v1;
        if (lane == 31) reinterpret_cast<uint4*>(text)[64] = v2;
        load_chunk(c + n_warps);                                              // next chunk's text: in flight during this chunk's work
        if (J.on && lane == 0) flat_pad_tiles(J, M, c, S.nB, pad_s, zeros_s, l2_first);   // this chunk's share of the pad columns

        // whitespace bits of my bytes; bytes outside [lo, hi) are not text.  Usual text: a byte is whitespace iff it is
        // <= 0x20 (exact set only when a control character is around); a lead of a multi-byte whitespace (C2, E1..E3) is
        // only looked at when the byte behind it is in 0x80..0xA0 (rules out the Vietnamese E1 BA / E1 BB and C3 letters)
        uint32_t vmask = exists ? 0xFFFFFFFFu : 0u;
        if (q0 < lo) vmask &= lo - q0 >= 32 ? 0u : (0xFFFFFFFFu << (lo - q0));
        if (q0 + 32 > hi) vmask &= q0 >= hi ? 0u : (0xFFFFFFFFu >> (q0 + 32 - hi));
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t ws = 0, hib = 0, low = 0, cand_any = 0;
        uint32_t secf[9], candf[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t t = ~((w[k] & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) & 0x80808080u;   // low seven bits <= 0x20
            ws |= gather_msb(t & ~w[k]) << (4 * k);
            secf[k] = t & w[k];                                                          // 0x80..0xA0
            hib |= w[k];
            low |= (w[k] - 0x20202020u) & ~w[k];
        }
        secf[8] = 0x80u;                                                                 // the byte behind my last one: not known
        if (low & 0x80808080u) {                                                         // some byte below 0x20: the exact set
            ws = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) ws |= gather_msb(ascii_ws4(w[k])) << (4 * k);
        }
        uint32_t spill = 0;       // whitespace bits that a code point starting in my granule puts into the next one
        if ((hib & 0x80808080u) && vmask) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                candf[k] = ws_lead4(w[k]) & __funnelshift_r(secf[k], secf[k + 1], 8);     // C2, E1..E3 and then 0x80..0xA0
                cand_any |= candf[k];
            }
            if (cand_any & 0x80808080u) {
                uint32_t lead = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) lead |= gather_msb(candf[k] & 0x80808080u) << (4 * k);
                lead &= vmask;
                while (lead) {
                    const int j = __ffs(lead) - 1;
                    lead &= lead - 1;
                    const int l = flat_mb_ws(tb, S.dsb, q0 + j, hi);
                    if (l) {
                        const uint64_t m = (uint64_t)((1u << l) - 1) << j;
                        ws |= (uint32_t)m;
                        spill |= (uint32_t)(m >> 32);
                    }
                }
            }
        }
        uint32_t nonws = ~ws & vmask;
        // bytes 0x0A that the word before them takes along (\S+\n?): only looked for when some byte is below 0x20
        uint32_t nlb = 0;
        if (low & 0x80808080u) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t x = w[k] ^ 0x0A0A0A0Au;
                nlb |= gather_msb(~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u) << (4 * k);
            }
            nlb &= vmask & ~dsbw;
        }
        // what the previous granule hands over: spilled whitespace bits, and whether its last byte is part of a word
        const uint32_t x_prev = __shfl_up_sync(FULL_MASK, spill | ((nonws >> 31) << 2), 1);
        uint32_t prevnon = 0;
        if (lane > 0) { nonws &= ~(x_prev & 3u); prevnon = x_prev >> 2; }
        const uint32_t term = ~nonws | dsbw;                    // a word cannot continue into these bytes
        const bool own = lane >= 1 && lane <= FC_OWN;
        const uint32_t st = own ? (nonws & (~((nonws << 1) | prevnon) | dsbw)) : 0u;
        // words before mine in the chunk
        const uint32_t cnt = __popc(st);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
        const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
        const uint32_t pref = incl - cnt;
        if (own) { S.st[q0 >> 5] = st; S.tpref[q0 >> 5] = (uint16_t)pref; }
        if (lane == 0) S.cnt[c] = total;
        // my words: position and length (terminator bits of my granule and the next one; the newline a word takes along
        // is the terminator itself).  Length 0: the end is further away, the word is measured bytewise.
        {
            const uint32_t term_n = __shfl_down_sync(FULL_MASK, term, 1);
            const bool any_nl = __any_sync(FULL_MASK, nlb != 0);
            const uint32_t nl_n = any_nl ? __shfl_down_sync(FULL_MASK, nlb, 1) : 0u;
            uint32_t m = st, k = pref;
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                const uint32_t z = __funnelshift_rc(term, term_n, b + 1);    // term bits b+1 .. b+32
                uint32_t len = 0;
                if (z) {
                    const int run = __ffs(z);
                    len = (uint32_t)run;
                    if (any_nl) len += (__funnelshift_rc(nlb, nl_n, b + 1) >> (run - 1)) & 1u;
                }
                sm.wl[k++] = (uint16_t)((lane * 32 + b) | (len << 10));
            }
        }
        __syncwarp();
        // one word per lane: its key from the staged text (aligned 16-byte pieces + funnel shifts), one probe of the word's
        // home slot -- the second half of the slot only for words longer than 8 bytes
        uint32_t* const wtok = S.wtok + ((size_t)c << FC_SHIFT);
        // (two words per lane and step: both probes are in flight before either is examined)
        auto prep = [&](uint32_t i, uint32_t& p, uint32_t& len, uint64_t& k0, uint64_t& k1, uint64_t& k2, uint32_t& h, uint4& a, uint4& b) {
            // returns true when the word takes the fast path (key in k0..k2, slot halves being loaded into a / b)
            a = make_uint4(0u, 0u, 0u, 0u); b = a; p = 0; len = 0; k0 = k1 = k2 = 0; h = 0;
            if (i >= total) return false;
            const uint32_t e = sm.wl[i];
            p = e & 1023u; len = e >> 10;
            if (len == 0 || len > KEY_INLINE) return false;
            const uint32_t a16 = p & ~15u;
            const int s16 = (int)(p & 15u);
            const uint4 x0 = *reinterpret_cast<const uint4*>(text + a16);
            const uint4 x1 = *reinterpret_cast<const uint4*>(text + a16 + 16);
            uint2 x2 = make_uint2(0u, 0u);
            if (s16 + (int)len > 32) x2 = *reinterpret_cast<const uint2*>(text + a16 + 32);
            key_from_pieces40_nomask(x0, x1, x2, s16, &k0, &k1, &k2);
            key_mask24(len, &k0, &k1, &k2);
            h = hash_key24(k0, k1, k2, len);
            const uint4* slot = reinterpret_cast<const uint4*>(&C.slots[h & C.mask]);
            a = ld_keep128(slot, l2_last);
            if (len > 8) b = ld_keep128(slot + 1, l2_last);          // (equal lengths <= 8: the rest of both keys is zero)
            return true;
        };
        auto done = [&](uint32_t i, bool fast, uint32_t p, uint32_t len, uint64_t k0, uint64_t k1, uint64_t k2, uint32_t h, const uint4& a, const uint4& b) {
            if (i >= total) return;
            uint32_t val;
            if (fast) {
                const uint64_t s0 = ((uint64_t)a.w << 32) | a.z, s1 = ((uint64_t)b.y << 32) | b.x, s2 = ((uint64_t)b.w << 32) | b.z;
                val = a.y;
                if (!((a.x == len) & (s0 == k0) & (s1 == k1) & (s2 == k2))) val = cache_find_or_insert(C, tb + wq + p, len, k0, k1, k2, h, insert_ok != 0);
            } else val = lookup_len(C, tb, wq + p, len ? len : flat_word_len(tb, S.dsb, wq + p, hi), insert_ok);
            if ((val & VAL_KIND) == VAL_PENDING) {
                const unsigned long long k = atomicAdd(&C.ctr[S.ctr_fix], 1ULL);
                if (k < S.fix_cap) { S.fixa[k] = (c << FC_SHIFT) + i; S.fixp[k] = wq + p; }
                else atomicAdd(&C.ctr[C_ERR], 1ULL);
            }
            wtok[i] = val;
        };
        if (ILP == 2) {
            for (uint32_t base = 0; base < total; base += 64) {
                const uint32_t i1 = base + lane, i2 = i1 + 32;
                uint32_t p1, l1, h1, p2, l2, h2; uint64_t ka0, ka1, ka2, kb0, kb1, kb2; uint4 a1, b1, a2, b2;
                const bool f1 = prep(i1, p1, l1, ka0, ka1, ka2, h1, a1, b1);
                const bool f2 = prep(i2, p2, l2, kb0, kb1, kb2, h2, a2, b2);
                done(i1, f1, p1, l1, ka0, ka1, ka2, h1, a1, b1);
                done(i2, f2, p2, l2, kb0, kb1, kb2, h2, a2, b2);
            }
        } else {
            for (uint32_t i = lane; i < total; i += 32) {
                uint32_t p1, l1, h1; uint64_t ka0, ka1, ka2; uint4 a1, b1;
                const bool f1 = prep(i, p1, l1, ka0, ka1, ka2, h1, a1, b1);
                done(i, f1, p1, l1, ka0, ka1, ka2, h1, a1, b1);
            }
        }
    }
    if (J.on && lane == 0) bulk_wait<0>();      // the constant buffer must outlive the tensor stores that read it
    __syncwarp();
}

// after k_bpe_pending: every pending word has its tokens now
__global__ void k_flat_fix(WordCache C, FlatSide S) {
    pdl_wait(); pdl_trigger();
    const int64_t o0 = S.off[0];
    const int64_t P0 = o0 & ~(int64_t)15;
    const uint8_t* tb = S.bytes + P0;
    const uint32_t hi = flat_hi(S, P0);
    const unsigned long long n = min(C.ctr[S.ctr_fix], (unsigned long long)S.fix_cap);
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t q = S.fixp[k];
        const uint32_t val = lookup_len(C, tb, q, flat_word_len(tb, S.dsb, q, hi), 0);
        if ((val & VAL_KIND) == VAL_PENDING) atomicAdd(&C.ctr[C_ERR], 1ULL);
        S.wtok[S.fixa[k]] = val;
    }
}

// ---- rows ----------------------------------------------------------------------------------------------------
struct FlatRowsArgs {
    FlatSide a, b;
    int32_t has_pair;
    int64_t n_rows;
    int32_t W, D;
    int32_t* ids; uint8_t* mask;
    int8_t* tt;                    // non-null: token types wanted
    int32_t* row_len; int32_t* seq_len; uint8_t* status;
    uint32_t* redo_list; uint32_t* fix_list;
    int8_t eos_i8;
};
struct __align__(16) FlatTile {
    uint32_t base[2][32];          // wtok index of the row's first word, per side
    uint32_t avail[2][32];         // words of the row that lie in the block of the first one
    uint32_t nw[2][32];            // words of the row
    int32_t dnA[32], dL[32];       // tokens of side A, framed length
    uint32_t dflag[32];
    uint32_t demit[32];
    SeqDesc dsd[32];
    int32_t tlo[32];               // pairs, the usual row: token types are 1 exactly on [tlo, m) -- else -1 (closed form per quad)
};
static const uint32_t FF_AGAIN = 8u;   // the row goes to the generic second pass

__device__ __forceinline__ uint32_t flat_rank(const FlatSide& S, uint32_t q, uint32_t* blk) {
    const uint32_t gi = q >> 5;
    *blk = q / (uint32_t)FC_BYTES;
    return (uint32_t)S.tpref[gi] + __popc(S.st[gi] & ((1u << (q & 31)) - 1u));
}
// index into wtok of word i of a row (first word at `base` in block base >> 13, `avail` words in that block)
__device__ __forceinline__ uint32_t flat_word_index(const FlatSide& S, uint32_t base, uint32_t avail, uint32_t i) {
    if (i < avail) return base + i;
    i -= avail;
    uint32_t b = (base >> FC_SHIFT) + 1;
    for (;;) {
        const uint32_t c = S.cnt[b];
        if (i < c) return (b << FC_SHIFT) + i;
        i -= c; b++;
    }
}
// Tokens of one side of one row, whole warp: words -> token counts -> positions (warp scan) -> staged row.
// Returns the next token position; stops once the row is full (pos >= limit).
__device__ __forceinline__ int32_t flat_side_tokens(const DevTables& T, const WordCache& C, const FlatSide& S, uint32_t base, uint32_t avail, uint32_t nw,
                                                    int32_t pos, int32_t limit, int32_t cap, int32_t* row, int lane, uint32_t* flags) {
    const int32_t spec_max = max(T.pad, max(T.bos, T.eos));
    for (uint32_t i0 = 0; i0 < nw && pos < limit; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t val = 0, nt = 0;
        if (i < nw) {
            val = S.wtok[flat_word_index(S, base, avail, i)];
            nt = (val & VAL_KIND) == VAL_SINGLE ? 1u : ((val & VAL_KIND) == VAL_MULTI ? C.tok_arena[val & VAL_PAYLOAD] : 0u);
        }
        uint32_t sc = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, sc, o); if (lane >= o) sc += t; }
        int32_t q = pos + (int32_t)(sc - nt);
        uint32_t fl = 0;
        if (nt == 1) {
            const int32_t t = (int32_t)(val & VAL_PAYLOAD);
            if (t <= spec_max && (t == T.eos || t == T.bos || t == T.pad)) fl = FF_AGAIN;
            if (q < cap) row[q] = t;
        } else if (nt) {
            const uint32_t* src = C.tok_arena + (val & VAL_PAYLOAD) + 1;
            for (uint32_t k = 0; k < nt && q < cap; k++, q++) {
                const int32_t t = (int32_t)src[k];
                if (t <= spec_max && (t == T.eos || t == T.bos || t == T.pad)) fl = FF_AGAIN;
                row[q] = t;
            }
        }
        *flags |= fl;
        pos += (int32_t)__shfl_sync(FULL_MASK, sc, 31);
    }
    return pos;
}

// Rows.  The staged ("real") columns 0..KR-1 of a row are computed four positions per lane and stored straight from
// registers (16-byte streaming stores); the pad columns KR..W-1 -- most of the bytes -- are written by the TMA unit from
// a block-wide constant buffer, two tensor stores per tile of D rows, with nothing to wait for.
template <int MINB, bool PAIR>
__global__ void __launch_bounds__(256, MINB) k_flat_rows(DevTables T, WordCache C, FlatRowsArgs A, const __grid_constant__ TmaPlanes M) {
    pdl_wait(); pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int32_t W = A.W, D = A.D, KR = M.KR, PB = M.PB;
    constexpr bool pair = PAIR;
    const bool want_tt = pair && A.tt;
    const uint32_t const_b = PB ? (uint32_t)tma_const_bytes(D, PB) : 0u;
    const uint32_t per_warp = (uint32_t)r128(sizeof(FlatTile)) + (uint32_t)r128((size_t)KR * 4);
    FlatTile* ts = reinterpret_cast<FlatTile*>(smem_raw + const_b + (size_t)wib * per_warp);
    int32_t* const rebuild = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(ts) + r128(sizeof(FlatTile)));
    const uint32_t const_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t zeros_s = const_s + (uint32_t)r128((size_t)D * PB * 4);
    if (PB) {
        const uint4 pad4 = make_uint4((uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad);
        uint4* cp = reinterpret_cast<uint4*>(smem_raw);
        const int n_pad = (int)(r128((size_t)D * PB * 4) >> 4), n_all = (int)(const_b >> 4);
        for (int i = threadIdx.x; i < n_all; i += blockDim.x) cp[i] = i < n_pad ? pad4 : make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        __syncthreads();
    }
    const int64_t P0a = A.a.off[0] & ~(int64_t)15;
    const int64_t P0b = pair ? (A.b.off[0] & ~(int64_t)15) : 0;
    const int32_t limit = W - 1;
    const int32_t cap = min(limit, KR);
    const int32_t spec_max = max(T.pad, max(T.bos, T.eos));
    const int32_t qpr = KR >> 2;                                       // quads per row in the real columns
    const uint32_t nw_clamp = (uint32_t)W + 8u;                          // more words than this cannot matter
    const uint32_t n_tiles = (uint32_t)((A.n_rows + D - 1) / D);
    const uint32_t n_warps = gridDim.x * (uint32_t)wpb;
    uint32_t tok_total = 0;
    const uint64_t l2_first = l2_policy_evict_first();
    // Two loads head a tile's chain of dependent loads: the document offsets of my row, then the start bits / prefix of the
    // two granules they point into.  Both are issued ahead: the offsets two tiles ahead, the rank data one tile ahead.
    int64_t pre[2][2] = {{0, 0}, {0, 0}};
    uint32_t nq[2][2] = {{0, 0}, {0, 0}}, ntp[2][2] = {{0, 0}, {0, 0}}, nst[2][2] = {{0, 0}, {0, 0}};
    auto prefetch_offsets = [&](uint32_t tile) {
        if (tile >= n_tiles) return;
        const int64_t r = (int64_t)tile * D + lane;
        if (lane < D && r < A.n_rows) {
            pre[0][0] = A.a.off[r]; pre[0][1] = A.a.off[r + 1];
            if (pair) { pre[1][0] = A.b.off[r]; pre[1][1] = A.b.off[r + 1]; }
        }
    };
    auto prefetch_ranks = [&](uint32_t tile) {          // from the offsets in `pre` (which belong to this tile)
        if (tile >= n_tiles) return;
        const int64_t r = (int64_t)tile * D + lane;
        if (lane < D && r < A.n_rows) {
#pragma unroll
            for (int s = 0; s < (PAIR ? 2 : 1); s++) {
                const FlatSide& S = s ? A.b : A.a;
                const int64_t P0 = s ? P0b : P0a;
                const uint32_t qmax = S.nB * (uint32_t)FC_BYTES - 64u;      // (offsets beyond the stated text were reported by k_flat_doc_starts)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t q = (uint32_t)min((uint64_t)(pre[s][e] - P0), (uint64_t)qmax);
                    nq[s][e] = q;
                    ntp[s][e] = S.tpref[q >> 5];
                    nst[s][e] = S.st[q >> 5];
                }
            }
        }
    };
    auto mask_word = [](int32_t c) -> uint32_t { return c >= 4 ? 0x01010101u : (c <= 0 ? 0u : (0x01010101u & ((1u << (8 * c)) - 1))); };
    {
        const uint32_t t0 = blockIdx.x * (uint32_t)wpb + wib;
        prefetch_offsets(t0);
        prefetch_ranks(t0);
        prefetch_offsets(t0 + n_warps);
    }
    for (uint32_t tile = blockIdx.x * (uint32_t)wpb + wib; tile < n_tiles; tile += n_warps) {
        const int64_t r0 = (int64_t)tile * D;
        const int nd = (int)min((int64_t)D, A.n_rows - r0);
        const uint32_t cq[2][2] = {{nq[0][0], nq[0][1]}, {nq[1][0], nq[1][1]}};
        const uint32_t ctp[2][2] = {{ntp[0][0], ntp[0][1]}, {ntp[1][0], ntp[1][1]}};
        const uint32_t cst[2][2] = {{nst[0][0], nst[0][1]}, {nst[1][0], nst[1][1]}};
        prefetch_ranks(tile + n_warps);
        prefetch_offsets(tile + 2 * n_warps);
        // ---- the pad columns of the tile: nothing to compute (rows that turn out longer are redone as a whole later)
        if (PB && lane == 0) {
            const int32_t r = (int32_t)r0;
            for (int32_t c0 = KR; c0 < W; c0 += PB) {
                tma_store_2d_hint(&M.ids_pad, const_s, c0, r, l2_first);
                tma_store_2d_hint(&M.mask_pad, zeros_s, c0, r, l2_first);
                if (want_tt) tma_store_2d_hint(&M.tt_pad, zeros_s, c0, r, l2_first);
            }
            bulk_commit();
        }
        // ---- 1. word ranges of my rows (one lane per row) from the ranks of the document offsets; framed length and
        //         token types as if every word were one token (the usual case)
        if (lane < nd) {
            uint32_t nws[2] = {0u, 0u};
            for (int s = 0; s < (pair ? 2 : 1); s++) {
                const FlatSide& S = s ? A.b : A.a;
                const uint32_t q0 = cq[s][0], q1 = cq[s][1];
                const uint32_t b0 = q0 / (uint32_t)FC_BYTES, b1 = q1 / (uint32_t)FC_BYTES;
                const uint32_t rl0 = ctp[s][0] + __popc(cst[s][0] & ((1u << (q0 & 31)) - 1u)), rl1 = ctp[s][1] + __popc(cst[s][1] & ((1u << (q1 & 31)) - 1u));
                uint32_t nw, avail;
                if (b0 == b1) { nw = rl1 - rl0; avail = nw; }
                else {
                    avail = S.cnt[b0] - rl0;
                    nw = avail + rl1;
                    for (uint32_t b = b0 + 1; b < b1 && nw < nw_clamp; b++) nw += S.cnt[b];
                }
                nw = min(nw, nw_clamp);
                ts->base[s][lane] = (b0 << FC_SHIFT) + rl0;
                ts->avail[s][lane] = avail;
                ts->nw[s][lane] = nw;
                nws[s] = nw;
            }
            const int32_t dL = (int32_t)(nws[0] + 2u + (pair ? nws[1] + 2u : 0u));
            ts->dnA[lane] = (int32_t)nws[0];
            ts->dL[lane] = dL;
            ts->dflag[lane] = 0u;
            if (pair) {
                const SeqDesc sd = seq_describe((int32_t)nws[0], dL, W);
                ts->dsd[lane] = sd;
                // no residual None, the two None of the framing become 0 / 1 right behind A, no trailing eos id: 0..0 1..1 0..0
                ts->tlo[lane] = (!sd.err && sd.r1 < 0 && sd.r2 < 0 && sd.f1 == sd.p1 && sd.f2 == sd.p1 + 1 && sd.m < W && sd.p1 > 0) ? sd.p1 + 1 : -1;
            }
        }
        __syncwarp();
        // ---- 2. the real columns of every row, four lanes per row, four positions per lane and step: ids, mask, token types
        for (int dg = 0; dg < D; dg += 8) {                              // (uniform trip count: the ballots below need every lane)
            const int d = dg + (lane >> 2);
            const bool live = d < nd;
            const uint32_t nwA = live ? ts->nw[0][d] : 0u, baseA = ts->base[0][d], availA = ts->avail[0][d];
            const uint32_t nwB = pair && live ? ts->nw[1][d] : 0u, baseB = pair ? ts->base[1][d] : 0u, availB = pair ? ts->avail[1][d] : 0u;
            const int32_t L = live ? ts->dL[d] : 0;
            const int32_t Lr = min(L, W);
            const uint32_t b0 = nwA + 3u, e2 = nwA + nwB + 3u;          // pairs: first position of B, position of the closing </s>
            const size_t grow = (size_t)(r0 + d) * (size_t)W;
            bool odd = false, special = false;
            const uint32_t a1 = nwA + 1u;                                  // position of the </s> behind A
            const bool cut = L >= W;                                       // truncated: the last column is </s> (tokenize.py:145)
            // the longest row of the group: behind it every quad of every row is padding (a warp-uniform test, no divergence)
            const int32_t gmax = __reduce_max_sync(FULL_MASK, live ? Lr : 0);
            const bool tt_plain = !(PAIR && want_tt) || __all_sync(FULL_MASK, !live || ts->tlo[d] >= 0);
            for (int32_t q = lane & 3; q < qpr && live; q += 4) {
                const int32_t j0 = q * 4;
                if ((q & ~3) * 4 >= gmax && tt_plain) {
                    st_cs128(A.ids + grow + j0, make_uint4((uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad));
                    st_cs32(A.mask + grow + j0, 0u);
                    if (PAIR && want_tt) st_cs32(A.tt + grow + j0, 0u);
                    continue;
                }
                // the four words (if any) first: independent loads, one round trip
                uint32_t val[4]; bool inA[4], inB[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t j = (uint32_t)(j0 + k), u = j - 1u;     // (j == 0: u wraps, not a word)
                    inA[k] = u < nwA;
                    inB[k] = PAIR && j >= b0 && j < e2;
                    val[k] = 0u;
                    if (inA[k]) val[k] = A.a.wtok[flat_word_index(A.a, baseA, availA, u)];
                    if (PAIR && inB[k]) val[k] = A.b.wtok[flat_word_index(A.b, baseB, availB, j - b0)];
                }
                int32_t t[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int32_t j = j0 + k;
                    // <s> A </s> [</s> B </s>] pad...   (tokenize.py:135,237-239,141-146)
                    int32_t v = j == 0 ? T.bos : (j < Lr ? T.eos : T.pad);
                    if (inA[k] || (PAIR && inB[k])) {
                        v = (int32_t)(val[k] & VAL_PAYLOAD);
                        if ((val[k] & VAL_KIND) != VAL_SINGLE) odd = true;
                        else if (v <= spec_max && (v == T.eos || v == T.bos || v == T.pad)) special = true;
                    }
                    t[k] = v;
                }
                if (cut && j0 + 4 == W) t[3] = T.eos;
                st_cs128(A.ids + grow + j0, make_uint4((uint32_t)t[0], (uint32_t)t[1], (uint32_t)t[2], (uint32_t)t[3]));
                st_cs32(A.mask + grow + j0, mask_word(Lr - j0));
                if (PAIR && want_tt) {
                    uint32_t ttw, sqw;
                    const int32_t lo1 = ts->tlo[d];
                    if (lo1 >= 0) ttw = mask_word(ts->dsd[d].m - j0) & ~mask_word(lo1 - j0);
                    else seq_words4(ts->dsd[d], j0, W, A.eos_i8, &ttw, &sqw);
                    st_cs32(A.tt + grow + j0, ttw);
                }
            }
            (void)a1;
            // a word that is not a single token: the positions need the token counts -- the whole warp rebuilds such rows
            uint32_t odd_rows = __ballot_sync(FULL_MASK, odd);
            const uint32_t special_rows = __ballot_sync(FULL_MASK, special);
            if (special_rows && lane < 8) {
                const int dd = dg + lane;
                if (dd < nd && ((special_rows >> (4 * lane)) & 0xFu)) ts->dflag[dd] = FF_AGAIN;
            }
            while (odd_rows) {
                const int dr = ((__ffs(odd_rows) - 1) >> 2);
                odd_rows &= ~(0xFu << (4 * dr));
                const int dd = dg + dr;
                __syncwarp();
                for (int32_t j = lane; j < KR; j += 32) rebuild[j] = T.pad;
                __syncwarp();
                if (lane == 0 && cap > 0) rebuild[0] = T.bos;
                uint32_t flags = 0;
                int32_t pos = flat_side_tokens(T, C, A.a, ts->base[0][dd], ts->avail[0][dd], ts->nw[0][dd], 1, limit, cap, rebuild, lane, &flags);
                const int32_t nA = pos - 1;
                if (pair) {
                    if (lane == 0) { if (pos < cap) rebuild[pos] = T.eos; if (pos + 1 < cap) rebuild[pos + 1] = T.eos; }
                    pos = flat_side_tokens(T, C, A.b, ts->base[1][dd], ts->avail[1][dd], ts->nw[1][dd], pos + 2, limit, cap, rebuild, lane, &flags);
                }
                const int32_t dL = pos + 1;
                if (lane == 0) { if (pos < cap) rebuild[pos] = T.eos; if (dL >= W && W - 1 < KR) rebuild[W - 1] = T.eos; }
                flags = __reduce_or_sync(FULL_MASK, flags);
                SeqDesc sd;
                if (pair) sd = seq_describe(nA, dL, W);
                if (lane == 0) { ts->dnA[dd] = nA; ts->dL[dd] = dL; ts->dflag[dd] |= flags; if (pair) { ts->dsd[dd] = sd; ts->tlo[dd] = -1; } }
                __syncwarp();
                const int32_t Lr2 = min(dL, W);
                const size_t g2 = (size_t)(r0 + dd) * (size_t)W;
                for (int32_t q = lane; q < qpr; q += 32) {
                    const int4 v = *reinterpret_cast<const int4*>(rebuild + q * 4);
                    st_cs128(A.ids + g2 + q * 4, make_uint4((uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w));
                    st_cs32(A.mask + g2 + q * 4, mask_word(Lr2 - q * 4));
                    if (want_tt) { uint32_t ttw, sqw; seq_words4(sd, q * 4, W, A.eos_i8, &ttw, &sqw); st_cs32(A.tt + g2 + q * 4, ttw); }
                }
            }
        }
        __syncwarp();
        // ---- 3. row bookkeeping (one lane per row)
        if (lane < nd) {
            const int32_t dL = ts->dL[lane];
            const uint32_t fl = ts->dflag[lane];
            const bool again = (fl & FF_AGAIN) || (KR < W && (dL > KR || (pair && ts->dsd[lane].m > KR)));
            if (again) {
                const unsigned long long k = atomicAdd(&C.ctr[C_REDO], 1ULL);
                A.redo_list[k] = (uint32_t)(r0 + lane);
            } else {
                const int64_t dr = r0 + lane;
                const int32_t Lr = dL < W ? dL : W;
                if (A.row_len) A.row_len[dr] = Lr;
                tok_total += (uint32_t)Lr;
                if (pair) {
                    if (A.seq_len) A.seq_len[dr] = ts->dsd[lane].m;
                    if (A.status) A.status[dr] = (uint8_t)ts->dsd[lane].err;
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait<0>();          // the constant buffer must outlive the tensor stores that read it
    __syncwarp();
    tok_total = __reduce_add_sync(FULL_MASK, tok_total);
    if (lane == 0 && tok_total) atomicAdd(&C.ctr[C_TOKENS], (unsigned long long)tok_total);
}

}  // namespace gzt
