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
//                       (a word whose BPE is pending is stored as its slot index: k_flat_rows fetches the value after k_bpe_pending)
//   k_flat_rows         a warp owns D rows: word ranges from the start bitmap (rank of the document offsets), framing,
//                       truncation, padding, mask, token types (tokenize.py:126-182,222-258) staged in final form and
//                       written by the TMA unit (TmaPlanes); rows it cannot finish go to the generic second pass
//
// Replaces the same reference lines as encode.cuh; outputs are bit-identical to the fused kernel's.
#pragma once
#include <type_traits>

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
    uint32_t* wtok;           // [nB*1024]   cache value (VAL_*) per word; VAL_PENDING | slot index while the word's BPE has not run
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
    bool ended = false;
    while (e < hi && !ended) {
        // eight bytes a step while none of them is ASCII whitespace, can start a multi-byte one, or starts a document (long tokens)
        if (e + 8 <= hi) {
            const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(tb + e) & ~(uintptr_t)7);
            const uint64_t v = bytes8(a[0], a[1], (uint32_t)(reinterpret_cast<uintptr_t>(tb + e) & 7) * 8u);
            const uint32_t d8 = __funnelshift_r(dsb[e >> 5], dsb[(e >> 5) + 1], e & 31u) & 0xFFu;
            if (!may_end_word8(v) && !d8) { e += 8; continue; }
        }
        for (int k = 0; k < 8 && e < hi; k++, e++) {
            if (dsb_bit(dsb, e)) return e - q;
            const uint32_t b = tb[e];
            if (b <= 0x20) { if (ascii_ws_byte(b)) { ended = true; break; } }
            else if (b >= 0xC2 && flat_mb_ws(tb, dsb, e, hi)) { ended = true; break; }
        }
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

// Everything the fast path of k_flat_words does not take: words longer than 16 bytes, words whose end lies beyond the bits the
// warp holds (len 0: measured bytewise).  Out of line: the kernel is short of registers.
__device__ __noinline__ uint32_t flat_lookup_general(const WordCache& C, const uint8_t* tb, const uint32_t* dsb, uint32_t q, uint32_t len, uint32_t hi, int insert_ok) {
    return lookup_len(C, tb, q, len ? len : flat_word_len(tb, dsb, q, hi), insert_ok);
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
// (both sides of a pair batch in one launch: blocks with an odd index take side B)
__global__ void k_flat_doc_starts(WordCache C, FlatSide Sa, FlatSide Sb, int two) {
    pdl_wait(); pdl_trigger();
    const bool second = two && (blockIdx.x & 1);
    const FlatSide& S = second ? Sb : Sa;
    const int64_t bid = two ? (blockIdx.x >> 1) : blockIdx.x, nblk = two ? ((gridDim.x + (second ? 0 : 1)) >> 1) : gridDim.x;
    const int64_t o0 = S.off[0];
    const int64_t P0 = o0 & ~(int64_t)15;
    const int64_t capq = (int64_t)S.nB * FC_BYTES - 64;
    for (int64_t d = bid * blockDim.x + threadIdx.x; d <= S.n; d += nblk * blockDim.x) {
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
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
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
// streaming reads (text, offsets: read once) must not push the word arrays out of L2 either
__device__ __forceinline__ uint4 ld_stream128(const void* p, uint64_t policy) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_keep32(void* p, uint32_t v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_keep16(void* p, uint16_t v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(p), "h"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint2 ld_keep64(const void* p, uint64_t policy) {
    uint2 v;
    asm volatile("ld.global.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
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
    int32_t l2_policy;   // bit 0: the text is read with the L2 evict-first policy; bit 1: the word arrays are stored with evict-last (experiments)
    uint32_t ratio;      // ceil(n_tiles * 2^20 / chunks of the launch): chunk c takes tiles [c * ratio >> 20, (c + 1) * ratio >> 20), clipped to n_tiles
};

// tiles [c * n_tiles / n_chunks, (c + 1) * n_tiles / n_chunks) of the pad columns; one thread.  Out of line: k_flat_words
// is short of registers.
__device__ __noinline__ void flat_pad_tiles(const PadJob& J, const TmaPlanes& M, uint32_t c, uint32_t n_chunks, uint32_t pad_s, uint32_t zeros_s, uint64_t l2_first) {
    (void)n_chunks;
    const uint32_t t0 = min((uint32_t)J.n_tiles, (uint32_t)(((uint64_t)c * J.ratio) >> 20)), t1 = min((uint32_t)J.n_tiles, (uint32_t)((((uint64_t)c + 1) * J.ratio) >> 20));
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
// Both sides of a pair batch in one launch: blocks [0, blocks_a) walk side 0, the others side 1.
struct FlatWordsArgs {
    FlatSide side[2];
    PadJob job[2];
    uint32_t blocks_a;
    int32_t insert_ok;
};
// A block is FW_WARPS word warps and one pad warp: lane 0 of the pad warp issues the block's share of the pad boxes, paced by the
// word warps' progress (at most FW_AHEAD chunk rounds ahead), so that no word warp ever waits for the TMA unit to take a store.
static const int FW_THREADS = (FW_WARPS + 1) * 32;
static const uint32_t FW_AHEAD = 1;
template <int MINB, int ILP>
__global__ void __launch_bounds__(FW_THREADS, MINB) k_flat_words(DevTables T, WordCache C, const __grid_constant__ FlatWordsArgs P, const __grid_constant__ TmaPlanes M) {
    pdl_wait(); pdl_trigger();
    const int which = blockIdx.x >= P.blocks_a ? 1 : 0;
    const FlatSide& S = P.side[which];
    const PadJob& J = P.job[which];
    const int insert_ok = P.insert_ok;
    const uint32_t block_id = blockIdx.x - (which ? P.blocks_a : 0u), n_blocks = which ? gridDim.x - P.blocks_a : P.blocks_a;
    __shared__ __align__(16) FlatWarpSmem s_warp[FW_WARPS];
    __shared__ uint4 s_keymask[17];                                          // byte masks of a 16-byte key by length
    __shared__ uint32_t s_prog[FW_WARPS];                                    // chunk rounds each word warp has started
    extern __shared__ __align__(1024) uint8_t pad_smem[];                   // J.on: [D x PB] pad ids, [D x PB] zero bytes
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    if (tid < 17) {
        auto m32 = [&](int i) -> uint32_t { const int nb = min(max(tid - 4 * i, 0), 4); return nb ? (0xFFFFFFFFu >> (8 * (4 - nb))) : 0u; };
        s_keymask[tid] = make_uint4(m32(0), m32(1), m32(2), m32(3));
    }
    if (tid < FW_WARPS) s_prog[tid] = 0u;
    __syncthreads();
    FlatWarpSmem& sm = s_warp[wib < FW_WARPS ? wib : 0];
    uint32_t pad_s = 0, zeros_s = 0; uint64_t l2_first = 0;
    const uint64_t l2_last = l2_policy_evict_last();
    const uint64_t l2_stream = (J.l2_policy & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t l2_keep = (J.l2_policy & 2) ? l2_last : l2_policy_evict_normal();
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
    const uint32_t n_warps = n_blocks * (uint32_t)FW_WARPS;
    // granule of this lane in chunk c is 30c - 1 + lane; its bytes start at q0 (may be "negative" for chunk 0, lane 0)
    const uint4 sp = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
    uint4 nv0 = sp, nv1 = sp, nv2 = sp; uint32_t ndsb = 0;
    auto load_chunk = [&](uint32_t c) {
        nv0 = sp; nv1 = sp; nv2 = sp; ndsb = 0;
        if (c >= S.nB) return;
        const int64_t gq = ((int64_t)c * FC_OWN - 1 + lane) * 32;
        if (gq >= 0) {
            const uint32_t q0 = (uint32_t)gq;
            if (q0 < hi) nv0 = ld_stream128(tb + q0, l2_stream);
            if (q0 + 16 < hi) nv1 = ld_stream128(tb + q0 + 16, l2_stream);
            if (lane == 31 && q0 + 32 < hi) nv2 = ld_stream128(tb + q0 + 32, l2_stream);
            ndsb = S.dsb[q0 >> 5];
        }
    };
    if (wib == FW_WARPS) {                                                    // the pad warp
        if (J.on && lane == 0) {
            volatile uint32_t* prog = s_prog;
            uint32_t round = 0;
            for (uint32_t c0 = block_id * (uint32_t)FW_WARPS; c0 < S.nB; c0 += n_warps, round++) {
                for (;;) {                                                    // stay at most FW_AHEAD rounds ahead of the slowest word warp
                    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
                    for (int w = 0; w < FW_WARPS; w++) mn = min(mn, prog[w]);
                    if (mn + FW_AHEAD > round) break;
                    __nanosleep(200);
                }
                for (uint32_t w = 0; w < (uint32_t)FW_WARPS && c0 + w < S.nB; w++) flat_pad_tiles(J, M, c0 + w, S.nB, pad_s, zeros_s, l2_first);
            }
            bulk_wait<0>();                                                   // the constant buffer must outlive the tensor stores that read it
        }
        return;
    }
    uint32_t c = block_id * (uint32_t)FW_WARPS + wib;
    uint32_t round = 0;
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
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&s_prog[wib]) = ++round;      // (the pad warp follows)

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
        // (what k_flat_rows reads right after this kernel: kept in L2 while the text and the planes stream through)
        if (own) { st_keep32(&S.st[q0 >> 5], st, l2_keep); st_keep16(&S.tpref[q0 >> 5], (uint16_t)pref, l2_keep); }
        if (lane == 0) st_keep32(&S.cnt[c], total, l2_keep);
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
        // one word per lane.  Words of up to 16 bytes (99.4 % of the occurrences of the vocabulary's text): the key is the five
        // aligned 32-bit words of the staged text around it, funnel-shifted to its start and cut to its length by a mask
        // from a table; the hash of device_common.cuh::hash_key24 restricted to those four key words; one probe of the word's
        // home slot (16 bytes, 8 more for keys longer than 8 bytes).  Everything else is out of line.
        uint32_t* const wtok = S.wtok + ((size_t)c << FC_SHIFT);
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t e = sm.wl[i];
            const uint32_t p = e & 1023u, len = e >> 10;
            uint32_t val;
            if (len - 1u < 16u) {
                const uint32_t* tw = reinterpret_cast<const uint32_t*>(text + (p & ~3u));
                const uint32_t w0 = tw[0], w1 = tw[1], w2 = tw[2], w3 = tw[3], w4 = tw[4];
                const uint32_t sh = (p & 3u) * 8u;
                const uint4 mk = s_keymask[len];
                const uint32_t k0 = __funnelshift_r(w0, w1, sh) & mk.x, k1 = __funnelshift_r(w1, w2, sh) & mk.y;
                const uint32_t k2 = __funnelshift_r(w2, w3, sh) & mk.z, k3 = __funnelshift_r(w3, w4, sh) & mk.w;
                uint32_t h = len * 0x9E3779B1u + k0 * 0xcc9e2d51u + k1 * 0x1b873593u + k2 * 0x85ebca6bu + k3 * 0xc2b2ae35u;
                h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
                const uint4* slot = reinterpret_cast<const uint4*>(&C.slots[h & C.mask]);
                const uint4 a = ld_keep128(slot, l2_last);
                uint2 b2 = make_uint2(0u, 0u);
                if (len > 8u) b2 = ld_keep64(slot + 1, l2_last);           // (equal lengths <= 8: the rest of both keys is zero)
                val = a.y;
                if (!((a.x == len) & (a.z == k0) & (a.w == k1) & (b2.x == k2) & (b2.y == k3)))
                    val = cache_find_or_insert(C, tb + wq + p, len, ((uint64_t)k1 << 32) | k0, ((uint64_t)k3 << 32) | k2, 0ULL, h, insert_ok != 0);
                else if ((val & VAL_KIND) == VAL_PENDING) val = VAL_PENDING | (h & C.mask);          // its BPE has not run: hand on the slot
            } else if (len - 17u < 8u) {
                // 17..24 bytes (glued or long compound words): the same from seven words of the staged text and both halves of the slot
                const uint32_t* tw = reinterpret_cast<const uint32_t*>(text + (p & ~3u));
                const uint32_t w0 = tw[0], w1 = tw[1], w2 = tw[2], w3 = tw[3], w4 = tw[4], w5 = tw[5], w6 = tw[6];
                const uint32_t sh = (p & 3u) * 8u;
                const uint4 mk = s_keymask[len - 16u];
                const uint32_t k0 = __funnelshift_r(w0, w1, sh), k1 = __funnelshift_r(w1, w2, sh), k2 = __funnelshift_r(w2, w3, sh), k3 = __funnelshift_r(w3, w4, sh);
                const uint32_t k4 = __funnelshift_r(w4, w5, sh) & mk.x, k5 = __funnelshift_r(w5, w6, sh) & mk.y;
                const uint64_t K0 = ((uint64_t)k1 << 32) | k0, K1 = ((uint64_t)k3 << 32) | k2, K2 = ((uint64_t)k5 << 32) | k4;
                const uint32_t h = hash_key24(K0, K1, K2, len);
                const uint4* slot = reinterpret_cast<const uint4*>(&C.slots[h & C.mask]);
                const uint4 a = ld_keep128(slot, l2_last), b4 = ld_keep128(slot + 1, l2_last);
                val = a.y;
                if (!((a.x == len) & (a.z == k0) & (a.w == k1) & (b4.x == k2) & (b4.y == k3) & (b4.z == k4) & (b4.w == k5)))
                    val = cache_find_or_insert(C, tb + wq + p, len, K0, K1, K2, h, insert_ok != 0);
                else if ((val & VAL_KIND) == VAL_PENDING) val = VAL_PENDING | (h & C.mask);
            } else val = flat_lookup_general(C, tb, S.dsb, wq + p, len, hi, insert_ok);
            st_keep32(&wtok[i], val, l2_keep);
        }
    }
    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&s_prog[wib]) = 0xFFFFFFFFu - FW_AHEAD;   // done: the pad warp need not wait for me
    __syncwarp();
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
    int8_t eos_i8, pad_i8;
    uint32_t pad_tile0;            // M.PB != 0: this kernel stores the pad columns of its tiles from this one on (the tiles before it are k_flat_words')
};
struct __align__(16) FlatTile {
    // per row of the tile, in the form the column lanes need (all word indices are into a.wtok; side B's array follows side A's):
    uint4 fa[32];                  // offA1, offA2, thA, nA1: column j in [1, nA1) holds word j - 1 of A, found at j + (j < thA ? offA1 : offA2)
                                   //   (a row's words lie in at most two chunks' segments of wtok; thA = where the second one starts)
    uint4 fb[32];                  // offB1, offB2, thB, e2: column j in [nA1 + 2, e2) holds a word of B, found at j + (j < thB ? offB1 : offB2)
    uint4 fc[32];                  // framed length L, tlo (pairs, the usual row: token types are 1 exactly on [tlo, m); else -1), m, flags
    uint32_t stage[32][32];        // columns 0..31 of the tile's rows in final form (the rebuild buffer of a row on the general path aliases it)
};
static const uint32_t FR_SLOW = 1u;    // the row takes the general path (whole warp, token scan)
static const uint32_t FR_ERR = 2u;     // pairs: the reference raises ValueError for this row
static const uint32_t FF_AGAIN = 8u;   // the row goes to the generic second pass

__device__ __forceinline__ uint32_t flat_rank(const FlatSide& S, uint32_t q, uint32_t* blk) {
    const uint32_t gi = q >> 5;
    *blk = q / (uint32_t)FC_BYTES;
    return (uint32_t)S.tpref[gi] + __popc(S.st[gi] & ((1u << (q & 31)) - 1u));
}
// value of a word as k_flat_words stored it; a word that was pending then has its tokens by now (k_bpe_pending ran in between)
__device__ __forceinline__ uint32_t flat_resolve(const WordCache& C, uint32_t val) {
    return (val & VAL_KIND) == VAL_PENDING ? C.slots[val & VAL_PAYLOAD].val : val;
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
            val = flat_resolve(C, S.wtok[flat_word_index(S, base, avail, i)]);
            nt = (val & VAL_KIND) == VAL_SINGLE ? 1u : ((val & VAL_KIND) == VAL_MULTI ? C.tok_arena[val & VAL_PAYLOAD] : 0u);
        }
        uint32_t sc = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, sc, o); if (lane >= o) sc += t; }
        int32_t q = pos + (int32_t)(sc - nt);
        uint32_t fl = 0;
        if ((val & VAL_KIND) == VAL_SINGLE) {
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

// Rows.  A warp owns a tile of 32 consecutive rows.
//  1. One lane per row turns the row's document offsets into word ranges (ranks in the start bitmap), the framed length and,
//     for pairs, the token-type description (tokenize.py:154-182 in closed form).
//  2. The usual row -- every word one token, at most 32 tokens, usual token types -- is assembled a LANE PER COLUMN: column j is
//     <s>, a word of A, </s>, </s>, a word of B, </s> or <pad> by position alone (tokenize.py:135,237-239,141-146), its value one
//     coalesced load from the word arrays; the 32 x 32 ids go to shared memory.
//  3. The planes are written a lane per 16 BYTES: the staged ids, the pad ids of the other staged columns, and mask / token
//     types as closed forms of the row's lengths (tokenize.py:148-152,254-258) from a table of 16-byte runs of ones.
//  4. A row that is not usual (a multi-token or pending-at-lookup word, more than 32 tokens, three chunks of text, unusual
//     token types) is rebuilt by the whole warp through a token scan; a row that needs more than the KR staged columns, or holds
//     a special id inside the text, goes to the generic second pass.
// The pad columns KR..W-1 are constant boxes written by the TMA unit (by k_flat_words on the side, or here for the tiles from
// pad_tile0 on).
template <int MINB, bool PAIR>
__global__ void __launch_bounds__(256, MINB) k_flat_rows(DevTables T, WordCache C, FlatRowsArgs A, const __grid_constant__ TmaPlanes M) {
    pdl_wait(); pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint4 s_ones[17];                                         // n bytes of 0x01 followed by zeros, n = 0..16
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int32_t W = A.W, D = A.D, KR = M.KR, PB = M.PB;
    constexpr bool pair = PAIR;
    const bool want_tt = pair && A.tt;
    const uint32_t const_b = PB ? (uint32_t)tma_const_bytes(D, PB) : 0u;
    const uint32_t per_warp = (uint32_t)r128(sizeof(FlatTile));
    FlatTile* ts = reinterpret_cast<FlatTile*>(smem_raw + const_b + (size_t)wib * per_warp);
    int32_t* const rebuild = reinterpret_cast<int32_t*>(&ts->stage[0][0]);           // KR <= 256 ids
    const uint32_t const_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t zeros_s = const_s + (uint32_t)r128((size_t)D * PB * 4);
    if (threadIdx.x < 17) {
        auto m32 = [&](int i) -> uint32_t { const int nb = min(max((int)threadIdx.x - 4 * i, 0), 4); return nb ? (0x01010101u >> (8 * (4 - nb))) : 0u; };
        s_ones[threadIdx.x] = make_uint4(m32(0), m32(1), m32(2), m32(3));
    }
    if (PB) {
        const uint4 pad4 = make_uint4((uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad);
        uint4* cp = reinterpret_cast<uint4*>(smem_raw);
        const int n_pad = (int)(r128((size_t)D * PB * 4) >> 4), n_all = (int)(const_b >> 4);
        for (int i = threadIdx.x; i < n_all; i += blockDim.x) cp[i] = i < n_pad ? pad4 : make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
    }
    __syncthreads();
    const int64_t P0a = A.a.off[0] & ~(int64_t)15;
    const int64_t P0b = pair ? (A.b.off[0] & ~(int64_t)15) : 0;
    const uint32_t* __restrict__ const wtok = A.a.wtok;
    const uint32_t segB = pair ? (uint32_t)(A.b.wtok - A.a.wtok) : 0u;   // side B's word array follows side A's in one allocation
    const int32_t limit = W - 1;
    const int32_t cap = min(limit, KR);
    const int32_t KQ = KR >> 2;                                          // 16-byte units of an ids row in the staged columns
    const uint32_t nw_clamp = (uint32_t)W + 8u;                          // more words than this cannot matter
    const uint32_t n_tiles = (uint32_t)((A.n_rows + D - 1) / D);
    const uint32_t n_warps = gridDim.x * (uint32_t)wpb;
    const uint4 pad4 = make_uint4((uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad, (uint32_t)T.pad);
    uint32_t tok_total = 0; uint32_t over32 = 0;
    const uint64_t l2_first = l2_policy_evict_first();
    // Two loads head a tile's chain of dependent loads: the document offsets of my row, then the start bits / prefix of the
    // two granules they point into.  Both are issued ahead: the offsets two tiles ahead, the rank data one tile ahead.
    int64_t pre[2][2] = {{0, 0}, {0, 0}};
    uint32_t nq[2][2] = {{0, 0}, {0, 0}}, ntp[2][2] = {{0, 0}, {0, 0}}, nst[2][2] = {{0, 0}, {0, 0}}, ncn[2] = {0, 0};
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
                // a row that runs over into the next chunk needs the word count of its first one
                if (nq[s][0] / (uint32_t)FC_BYTES != nq[s][1] / (uint32_t)FC_BYTES) ncn[s] = S.cnt[nq[s][0] / (uint32_t)FC_BYTES];
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
        const uint32_t ccn[2] = {ncn[0], ncn[1]};
        prefetch_ranks(tile + n_warps);
        prefetch_offsets(tile + 2 * n_warps);
        // ---- the pad columns of the tile: nothing to compute (rows that turn out longer are redone as a whole later)
        if (PB && tile >= A.pad_tile0 && lane == 0) {
            const int32_t r = (int32_t)r0;
            for (int32_t c0 = KR; c0 < W; c0 += PB) {
                tma_store_2d_hint(&M.ids_pad, const_s, c0, r, l2_first);
                tma_store_2d_hint(&M.mask_pad, zeros_s, c0, r, l2_first);
                if (want_tt) tma_store_2d_hint(&M.tt_pad, zeros_s, c0, r, l2_first);
            }
            bulk_commit();
        }
        // ---- 1. my row (one lane per row): word ranges from the ranks of the document offsets; framed length and token
        //         types as if every word were one token (the usual case)
        uint32_t my_flags = 0;
        if (lane < nd) {
            uint32_t f[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};       // first word's index, words in its chunk, words, index of the next chunk's first word
#pragma unroll
            for (int s = 0; s < (PAIR ? 2 : 1); s++) {
                const FlatSide& S = s ? A.b : A.a;
                const uint32_t q0 = cq[s][0], q1 = cq[s][1];
                const uint32_t b0 = q0 / (uint32_t)FC_BYTES, b1 = q1 / (uint32_t)FC_BYTES;
                const uint32_t rl0 = ctp[s][0] + __popc(cst[s][0] & ((1u << (q0 & 31)) - 1u)), rl1 = ctp[s][1] + __popc(cst[s][1] & ((1u << (q1 & 31)) - 1u));
                uint32_t nw, avail;
                if (b0 == b1) { nw = rl1 - rl0; avail = nw; }
                else {
                    avail = ccn[s] - rl0;
                    nw = avail + rl1;
                    if (b1 > b0 + 1) {                                        // three or more chunks of text: the general path
                        my_flags |= FR_SLOW;
                        for (uint32_t b = b0 + 1; b < b1 && nw < nw_clamp; b++) nw += S.cnt[b];
                    }
                }
                nw = min(nw, nw_clamp);
                f[s][0] = (b0 << FC_SHIFT) + rl0; f[s][1] = avail; f[s][2] = nw; f[s][3] = (b0 + 1u) << FC_SHIFT;
            }
            const int32_t dL = (int32_t)(f[0][2] + 2u + (pair ? f[1][2] + 2u : 0u));
            int32_t tlo = -1, m = 0;
            if (pair) {
                const SeqDesc sd = seq_describe((int32_t)f[0][2], dL, W);
                // no residual None, the two None of the framing become 0 / 1 right behind A, no trailing eos id: 0..0 1..1 0..0
                tlo = (!sd.err && sd.r1 < 0 && sd.r2 < 0 && sd.f1 == sd.p1 && sd.f2 == sd.p1 + 1 && sd.m < W && sd.p1 > 0) ? sd.p1 + 1 : -1;
                m = sd.m;
                if (sd.err) my_flags |= FR_ERR;
                if (tlo < 0) my_flags |= FR_SLOW;
            }
            if (dL > 32 || KR < 32) my_flags |= FR_SLOW;
            const uint32_t nA1 = f[0][2] + 1u, bc0 = nA1 + 2u;
            // the words of my row: asked for now, used in step 2 (they come from DRAM more often than from L2)
            if (!(my_flags & FR_SLOW)) {
                prefetch_l2(wtok + f[0][0]); prefetch_l2(wtok + f[0][0] + (f[0][1] ? f[0][1] - 1u : 0u));
                if (pair) { prefetch_l2(wtok + segB + f[1][0]); prefetch_l2(wtok + segB + f[1][0] + (f[1][1] ? f[1][1] - 1u : 0u)); }
            }
            ts->fa[lane] = make_uint4(f[0][0] - 1u, f[0][3] - f[0][1] - 1u, f[0][1] + 1u, nA1);
            ts->fb[lane] = make_uint4(segB + f[1][0] - bc0, segB + f[1][3] - f[1][1] - bc0, bc0 + f[1][1], bc0 + f[1][2]);
            ts->fc[lane] = make_uint4((uint32_t)dL, (uint32_t)tlo, (uint32_t)m, my_flags);
        }
        __syncwarp();
        uint32_t slow_rows = __ballot_sync(FULL_MASK, (my_flags & FR_SLOW) != 0);
        // ---- 2. the usual rows, a lane per column: 32 ids per row into the staging area.  Eight rows at a time: their word
        //         values are all requested before the first is used (the loads miss L2 more often than not: the planes that
        //         stream through it have pushed the word arrays out).
        {
            const uint32_t j = (uint32_t)lane;
            uint32_t oddbits = 0;
            // (a full tile without rows for the general path -- nearly every tile -- runs without the per-row tests)
            auto assemble = [&](auto guarded) {
                constexpr bool G = decltype(guarded)::value;
                for (int dg = 0; dg < (G ? nd : 32); dg += 8) {
                    uint32_t val[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int d = dg + k;
                        val[k] = VAL_SINGLE;
                        if (!G || (d < nd && !((slow_rows >> d) & 1u))) {
                            const uint4 fa = ts->fa[d];
                            const uint32_t L = ts->fc[d].x;
                            const bool a = (j - 1u) < (fa.w - 1u);
                            uint32_t idx = j + (j < fa.z ? fa.x : fa.y);
                            bool inw = a;
                            if (PAIR) {
                                const uint4 fb = ts->fb[d];
                                const uint32_t bc0 = fa.w + 2u;
                                const bool b = (j - bc0) < (fb.w - bc0);
                                const uint32_t idxb = j + (j < fb.z ? fb.x : fb.y);
                                idx = a ? idx : idxb;
                                inw = a | b;
                            }
                            // <s> A </s> [</s> B </s>] pad...   (tokenize.py:135,237-239,141-146)
                            val[k] = VAL_SINGLE | (uint32_t)(j == 0 ? T.bos : (j < L ? T.eos : T.pad));
                            if (inw) val[k] = wtok[idx];
                        }
                    }
                    // a word whose BPE was pending at look-up (a cold cache: nearly all of them) has its value by now: one more load
#pragma unroll
                    for (int k = 0; k < 8; k++) if ((val[k] >> 30) == 0u) val[k] = C.slots[val[k]].val;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int d = dg + k;
                        if (!G || d < nd) {
                            oddbits |= (uint32_t)((val[k] >> 30) != 1u) << d;   // not "one token": the row takes the general path
                            ts->stage[d][lane] = val[k] & VAL_PAYLOAD;
                        }
                    }
                }
            };
            if (nd == 32 && slow_rows == 0u) assemble(std::false_type{}); else assemble(std::true_type{});
            slow_rows |= __reduce_or_sync(FULL_MASK, oddbits);
        }
        __syncwarp();
        // ---- 3. the planes of the usual rows, a lane per 16 bytes
        auto write_out = [&](auto guarded) {
            constexpr bool G = decltype(guarded)::value;
            const size_t g0 = (size_t)r0 * (size_t)W;
#pragma unroll
            for (int it = 0; it < 8; it++) {                               // ids, columns 0..31: 8 lanes per row
                const int d = it * 4 + (lane >> 3), q = lane & 7;
                if (!G || (d < nd && !((slow_rows >> d) & 1u)))
                    st_cs128(A.ids + g0 + (size_t)(d * W + q * 4), *reinterpret_cast<const uint4*>(&ts->stage[d][q * 4]));
            }
            for (int32_t c0 = 32; c0 < KR; c0 += 32) {                     // ids, the other staged columns: pad ids
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int d = it * 4 + (lane >> 3);
                    const int32_t col = c0 + (lane & 7) * 4;
                    if ((!G || (d < nd && !((slow_rows >> d) & 1u))) && col < KR) st_cs128(A.ids + g0 + (size_t)(d * W + col), pad4);
                }
            }
            for (int32_t c0 = 0; c0 < KR; c0 += 64) {                      // mask and token types: 4 lanes per row and pass
#pragma unroll
                for (int it = 0; it < 4; it++) {
                    const int d = it * 8 + (lane >> 2);
                    const int32_t col = c0 + (lane & 3) * 16;
                    if ((!G || (d < nd && !((slow_rows >> d) & 1u))) && col < KR) {
                        const uint4 fc = ts->fc[d];
                        const size_t g = g0 + (size_t)(d * W + col);
                        st_cs128(A.mask + g, s_ones[min(max((int32_t)fc.x - col, 0), 16)]);
                        if (want_tt) {
                            const uint4 hi1 = s_ones[min(max((int32_t)fc.z - col, 0), 16)], lo1 = s_ones[min(max((int32_t)fc.y - col, 0), 16)];
                            st_cs128(A.tt + g, make_uint4(hi1.x & ~lo1.x, hi1.y & ~lo1.y, hi1.z & ~lo1.z, hi1.w & ~lo1.w));
                        }
                    }
                }
            }
        };
        if (nd == 32 && slow_rows == 0u && KR >= 32) write_out(std::false_type{}); else if (KR >= 32) write_out(std::true_type{});
        // ---- 4. the other rows: the whole warp rebuilds each through a token scan
        slow_rows &= nd >= 32 ? 0xFFFFFFFFu : ((1u << nd) - 1u);
        while (slow_rows) {
            const int d = __ffs(slow_rows) - 1;
            slow_rows &= slow_rows - 1;
            __syncwarp();
            const uint4 fa = ts->fa[d];
            const uint4 fb = PAIR ? ts->fb[d] : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t nwA = fa.w - 1u, bc0 = fa.w + 2u;
            const size_t grow = (size_t)(r0 + d) * (size_t)W;
            __syncwarp();
            for (int32_t j = lane; j < KR; j += 32) rebuild[j] = T.pad;
            __syncwarp();
            if (lane == 0 && cap > 0) rebuild[0] = T.bos;
            uint32_t flags = 0;
            int32_t pos = flat_side_tokens(T, C, A.a, fa.x + 1u, fa.z - 1u, nwA, 1, limit, cap, rebuild, lane, &flags);
            const int32_t nA = pos - 1;
            if (pair) {
                if (lane == 0) { if (pos < cap) rebuild[pos] = T.eos; if (pos + 1 < cap) rebuild[pos + 1] = T.eos; }
                pos = flat_side_tokens(T, C, A.b, fb.x - segB + bc0, fb.z - bc0, fb.w - bc0, pos + 2, limit, cap, rebuild, lane, &flags);
            }
            const int32_t dL = pos + 1;
            if (lane == 0) { if (pos < cap) rebuild[pos] = T.eos; if (dL >= W && W - 1 < KR) rebuild[W - 1] = T.eos; }
            flags = __reduce_or_sync(FULL_MASK, flags);
            SeqDesc sd;
            sd.m = 0; sd.err = 0;
            if (pair) sd = seq_describe(nA, dL, W);
            __syncwarp();
            const int32_t Lr2 = min(dL, W);
            for (int32_t q = lane; q < KQ; q += 32) {
                const int4 t4 = *reinterpret_cast<const int4*>(rebuild + q * 4);
                st_cs128(A.ids + grow + q * 4, make_uint4((uint32_t)t4.x, (uint32_t)t4.y, (uint32_t)t4.z, (uint32_t)t4.w));
                st_cs32(A.mask + grow + q * 4, mask_word(Lr2 - q * 4));
                if (want_tt) { uint32_t ttw, sqw; seq_words4(sd, q * 4, W, A.eos_i8, A.pad_i8, &ttw, &sqw); st_cs32(A.tt + grow + q * 4, ttw); }
            }
            __syncwarp();
            if (lane == 0) ts->fc[d] = make_uint4((uint32_t)dL, 0xFFFFFFFFu, (uint32_t)sd.m, (flags & FF_AGAIN) | (sd.err ? FR_ERR : 0u));
        }
        __syncwarp();
        // ---- 5. row bookkeeping (one lane per row)
        if (lane < nd) {
            const uint4 fc = ts->fc[lane];
            const int32_t dL = (int32_t)fc.x, m = (int32_t)fc.z;
            if (dL > 32 || (pair && m > 32)) over32++;               // (the host stages 32 columns next time when hardly any row needs more)
            const bool again = (fc.w & FF_AGAIN) || (KR < W && (dL > KR || (pair && m > KR)));
            if (again) {
                const unsigned long long k = atomicAdd(&C.ctr[C_REDO], 1ULL);
                A.redo_list[k] = (uint32_t)(r0 + lane);
            } else {
                const int64_t dr = r0 + lane;
                const int32_t Lr = dL < W ? dL : W;
                if (A.row_len) A.row_len[dr] = Lr;
                tok_total += (uint32_t)Lr;
                if (pair) {
                    if (A.seq_len) A.seq_len[dr] = m;
                    if (A.status) A.status[dr] = (uint8_t)((fc.w & FR_ERR) ? 1 : 0);
                }
            }
        }
        __syncwarp();
    }
    if (PB && lane == 0) bulk_wait<0>();          // the constant buffer must outlive the tensor stores that read it
    __syncwarp();
    tok_total = __reduce_add_sync(FULL_MASK, tok_total);
    if (lane == 0 && tok_total) atomicAdd(&C.ctr[C_TOKENS], (unsigned long long)tok_total);
    over32 = __reduce_add_sync(FULL_MASK, over32);
    if (lane == 0 && over32) atomicAdd(&C.ctr[C_OVER32], (unsigned long long)over32);
}

}  // namespace gzt
