// post.cuh -- generic per-row kernels: ragged-layout bookkeeping (row lengths, scan), and the
// value-driven attention_mask / sequence_id / token_type_ids pass used for ragged rows, for rows
// that contain special ids inside the text, and for the public helper methods.
//
//   k_row_lens     tokenize.py:141-146 (__padding) as lengths: what to keep, how long the row gets
//   k_scan_i64     exclusive scan -> row offsets (north_star step 5)
//   k_post_rows    tokenize.py:148-152 (mask), :163-182 (get_sequence_id), :154-161 (get_token_type),
//                  :256-258 (token_type_ids padding), evaluated on the actual id values
//   k_decode_*     tokenize.py:137-139 (decode) as a gather of precomputed per-id forms
#pragma once
#include "device_common.cuh"

namespace gzt {

struct LenArgs {
    const int32_t* L;      // framed length per row
    int64_t n_rows;
    int32_t has_max_len, max_len, padding, truncation;
    int32_t* keep;         // framed tokens copied
    int64_t* out_len;      // final row length
    uint8_t* tail;         // 0 none, 1 pad fill [keep, out_len), 2 eos at keep
    const int32_t* nwA; const int32_t* nwB;   // return_offset: words per side (NULL when not wanted)
    int64_t* span_cnt;     // entries of the row's offset list: (nwA + 2) [+ (nwB + 2)]
};

__global__ void k_row_lens(LenArgs A) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t L = A.L[r];
        int64_t keep = L, out = L; uint8_t tail = 0;
        if (A.has_max_len && A.padding) {                       // tokenize.py:247
            if (L < (int64_t)A.max_len) { out = A.max_len; tail = 1; }          // :142-143
            else if (A.truncation) { keep = py_head(L, (int64_t)A.max_len - 1); out = keep + 1; tail = 2; }   // :144-145
        }
        A.keep[r] = (int32_t)keep; A.out_len[r] = out; A.tail[r] = tail;
        if (A.span_cnt) A.span_cnt[r] = (int64_t)A.nwA[r] + 2 + (A.nwB ? (int64_t)A.nwB[r] + 2 : 0);
    }
}

// out[0..n] = exclusive scan of in[0..n) (out[n] = total). One block.
__global__ void __launch_bounds__(1024) k_scan_i64(const int64_t* in, int64_t* out, int64_t n) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int64_t v = i < n ? in[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += t; }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int64_t s = warp_sums[lane], y = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(FULL_MASK, y, o); if (lane >= o) y += t; }
            warp_sums[lane] = y - s;
        }
        __syncthreads();
        const int64_t c = carry_s;
        const int64_t incl = c + warp_sums[wid] + x;
        if (i < n) out[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// Large inputs: tile-local scans + a scan of the tile totals + a fix-up pass.
static const int SCAN_TILE = 4096;          // elements per block (1024 threads x 4)
__global__ void __launch_bounds__(1024) k_scan_tiles(const int64_t* in, int64_t* out, int64_t* tile_sum, int64_t n) {
    __shared__ int64_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * 4;
    int64_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    int64_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += t; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int64_t w = warp_sums[lane], y = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(FULL_MASK, y, o); if (lane >= o) y += t; }
        warp_sums[lane] = y - w;
        if (lane == 31) tile_sum[blockIdx.x] = y;
    }
    __syncthreads();
    int64_t run = warp_sums[wid] + x - s;       // exclusive prefix of this thread inside the tile
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
}
__global__ void __launch_bounds__(1024) k_scan_fix(int64_t* out, const int64_t* tile_excl, int64_t n, int64_t n_tiles) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * 4;
    const int64_t add = tile_excl[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_excl[n_tiles];
}

struct PostArgs {
    int32_t* ids;              // flat
    const int64_t* row_off;    // NULL -> fixed layout, row r at r*W
    int32_t W;
    int64_t n_rows;
    const uint32_t* row_list;  // optional subset (count at ctr[C_FIX])
    const int32_t* keep;       // with tail: fill [keep, len) first
    const uint8_t* tail;
    uint8_t* mask;             // may be NULL
    int32_t has_pair;
    int8_t* tt; int8_t* seq;   // row-aligned planes, may be NULL
    int32_t* tt_len; int32_t* seq_len; uint8_t* status; int32_t* row_len;
    int32_t has_max_len, max_len, padding, truncation;
    int32_t raw_seq;           // 1: write get_sequence_id's result without get_token_type (helper API)
    int8_t eos_i8;
    unsigned long long* tokens_ctr;   // += sum(mask)
};

__global__ void __launch_bounds__(256) k_post_rows(DevTables T, PostArgs A, const unsigned long long* list_count) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_items = A.row_list ? *list_count : (uint64_t)A.n_rows;
    unsigned long long tok_total = 0;
    for (uint64_t it = warp; it < n_items; it += nwarps) {
        const int64_t r = A.row_list ? (int64_t)A.row_list[it] : (int64_t)it;
        const int64_t start = A.row_off ? A.row_off[r] : r * (int64_t)A.W;
        const int32_t n = A.row_off ? (int32_t)(A.row_off[r + 1] - start) : A.W;
        int32_t* Tk = A.ids + start;
        if (A.tail) {                                               // tokenize.py:143 / :145
            const int32_t kp = A.keep[r]; const uint8_t tl = A.tail[r];
            if (tl == 1) for (int32_t i = kp + lane; i < n; i += 32) Tk[i] = T.pad;
            else if (tl == 2 && lane == 0) Tk[kp] = T.eos;
            __syncwarp();
            if (A.row_len && lane == 0) A.row_len[r] = tl == 1 ? kp : n;
        }
        if (A.mask) {                                               // tokenize.py:148-152
            int32_t cnt = 0;
            for (int32_t i = lane; i < n; i += 32) { uint8_t on = Tk[i] != T.pad; A.mask[start + i] = on; cnt += on; }
            cnt = __reduce_add_sync(FULL_MASK, cnt);
            tok_total += (unsigned long long)cnt;
        }
        if (!A.has_pair) continue;
        // ---- get_sequence_id (tokenize.py:163-182)
        int32_t p1 = n;                                             // first </s>
        for (int32_t base = 0; base < n && p1 == n; base += 32) {
            const int32_t i = base + lane;
            const uint32_t m = __ballot_sync(FULL_MASK, i < n && Tk[i] == T.eos);
            if (m) p1 = base + __ffs(m) - 1;
        }
        int32_t e = -1;                                             // closing </s>: first i >= p1+2 with T[i]==eos and T[i-1]!=eos
        for (int32_t base = p1 + 2; base < n && e < 0; base += 32) {
            const int32_t i = base + lane;
            const uint32_t m = __ballot_sync(FULL_MASK, i < n && Tk[i] == T.eos && Tk[i - 1] != T.eos);
            if (m) e = base + __ffs(m) - 1;
        }
        const int32_t m_len = e >= 0 ? e + 1 : n;
        // ---- get_token_type (tokenize.py:154-161): S[0]=0, S[-1]=1, first two remaining None -> 0, 1
        int32_t f1 = -1, f2 = -1;
        if (!A.raw_seq) {
            for (int32_t base = 0; base < m_len && f2 < 0; base += 32) {
                const int32_t i = base + lane;
                bool none = false;
                if (i < m_len && i != 0 && i != m_len - 1) {
                    const int32_t t = Tk[i];
                    none = (i < p1 && t == T.bos) || i == p1 || (i > p1 && t == T.eos);
                }
                uint32_t m = __ballot_sync(FULL_MASK, none);
                if (m && f1 < 0) { f1 = base + __ffs(m) - 1; m &= m - 1; }
                if (m && f2 < 0) f2 = base + __ffs(m) - 1;
            }
        }
        const bool err = !A.raw_seq && (f2 < 0 || m_len == 0);
        // token_type_ids length (tokenize.py:256-258)
        int32_t tt_keep = m_len, tt_len = m_len; int tt_tail = 0;
        if (A.has_max_len && A.padding) {
            if (m_len < A.max_len) { tt_len = A.max_len; tt_tail = 1; }
            else if (A.truncation) { tt_keep = (int32_t)py_head(m_len, (int64_t)A.max_len - 1); tt_len = tt_keep + 1; tt_tail = 2; }
        }
        const int32_t hi = n;
        for (int32_t i = lane; i < hi; i += 32) {
            int32_t v = -2;
            if (i < m_len) {
                const int32_t t = Tk[i];
                const bool none = (i < p1 && t == T.bos) || i == p1 || (i > p1 && t == T.eos);
                v = none ? -1 : (i < p1 ? 0 : 1);
                if (!A.raw_seq) {
                    if (i == f1) v = 0;
                    if (i == f2) v = 1;
                    if (i == 0) v = 0;
                    if (i == m_len - 1) v = 1;
                }
            }
            if (A.seq) A.seq[start + i] = (int8_t)v;
            if (A.tt) {
                int32_t tv = i < tt_keep ? v : (tt_tail == 2 && i == tt_keep ? (int32_t)A.eos_i8 : 0);
                if (i < tt_len) A.tt[start + i] = (int8_t)tv;
            }
        }
        if (lane == 0) {
            if (A.seq_len) A.seq_len[r] = m_len;
            if (A.tt_len) A.tt_len[r] = tt_len;
            if (A.status) A.status[r] = err ? 1 : 0;
        }
    }
    if (A.tokens_ctr && lane == 0 && tok_total) atomicAdd(A.tokens_ctr, tok_total);
}

// ---- decode (tokenize.py:137-139) ------------------------------------------------------------------
struct DecArgs {
    const int32_t* ids;
    const int64_t* ids_off;   // NULL -> n rows of `width`
    int32_t width;
    int64_t n_rows;
    int64_t* out_len;         // per row bytes (pass 1)
    const int64_t* out_off;   // per row start (pass 2)
    uint8_t* out;
};

__device__ __forceinline__ void dec_form(const DevTables& T, int32_t id, bool last, uint32_t* off, uint32_t* len) {
    const uint32_t k = (id >= 0 && id < T.n_ids) ? (uint32_t)id : (uint32_t)T.n_ids;   // decoder.get(i, unk_token)
    const uint32_t dsc = last ? T.last_desc[k] : T.mid_desc[k];
    *off = (dsc >> 8) * 8u;
    *len = dsc & 255u;
    if (*len == 255u) *len = last ? T.last_len[k] : T.mid_len[k];                       // very long vocab entry
}

static const int DEC_CAP = 2048;       // bytes staged per warp before a flush
static const int DEC_MAXFORM = 32;     // forms longer than this go straight to global memory

// Pass 1 (WRITE == false): bytes per row.  Pass 2: gather the forms into a per-warp shared-memory staging buffer
// and write the text with aligned 16-byte stores (edges bytewise: neighbouring rows belong to other warps).
template <bool WRITE>
__global__ void __launch_bounds__(256, WRITE ? 4 : 6) k_decode(DevTables T, DecArgs A) {
    __shared__ __align__(16) uint8_t stage[WRITE ? 8 : 1][WRITE ? DEC_CAP + 16 : 16];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint8_t* ob = stage[WRITE ? wib : 0];
    // The padding of fixed-width rows decodes to one short form repeated: sixteen-byte units of that periodic text, one per
    // phase, are kept per warp so that a run of pad ids is written as whole units instead of byte by byte.
    __shared__ __align__(16) uint8_t padtab[WRITE ? 8 : 1][WRITE ? 8 : 1][16];
    uint32_t padL = 0; uint64_t padP = 0;
    if (WRITE) {
        uint32_t off, len;
        dec_form(T, T.pad, false, &off, &len);
        if (len >= 1 && len <= 8) {
            padL = len;
            padP = *reinterpret_cast<const uint64_t*>(T.form_blob + off);
            if (lane < (int)len) {
                int ph = lane;
                for (int k = 0; k < 16; k++) { padtab[wib][lane][k] = (uint8_t)(padP >> (8 * ph)); ph = ph + 1 == (int)len ? 0 : ph + 1; }
            }
        }
        __syncwarp();
    }
    for (uint64_t r = warp; r < (uint64_t)A.n_rows; r += nwarps) {
        const int64_t start = A.ids_off ? A.ids_off[r] : (int64_t)r * A.width;
        const int64_t n = A.ids_off ? A.ids_off[r + 1] - start : A.width;
        const int32_t* ids = A.ids + start;
        if (!WRITE) {
            int64_t run = 0;
            for (int64_t base = 0; base < n; base += 256) {              // eight batches of ids in flight
                int32_t idv[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { const int64_t i = base + 32 * k + lane; idv[k] = i < n ? ids[i] : 0; }
                uint32_t sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int64_t i = base + 32 * k + lane;
                    uint32_t off = 0, len = 0;
                    if (i < n) dec_form(T, idv[k], i == n - 1, &off, &len);
                    sum += len;
                }
                run += __reduce_add_sync(FULL_MASK, sum);
            }
            if (lane == 0) A.out_len[r] = run;
            continue;
        }
        const int64_t g0 = A.out_off[r];
        uint8_t* gbase = A.out + (g0 & ~(int64_t)15);      // global address of ob[0]
        int lead = (int)(g0 & 15);                          // ob[0..lead) is not ours (previous row)
        int cur = lead;
        // flush ob[0..upto) (upto multiple of 16, or everything when last): aligned 16-byte stores, foreign/partial units bytewise
        auto flush = [&](int upto, bool last) {
            __syncwarp();
            const int nfull = last ? (cur & ~15) : upto;
            for (int u = lane * 16; u < nfull; u += 512) {
                if (u == 0 && lead) { for (int k = lead; k < 16; k++) gbase[k] = ob[k]; }
                else *reinterpret_cast<uint4*>(gbase + u) = *reinterpret_cast<const uint4*>(ob + u);
            }
            if (last) {
                const int rem = cur - nfull;                // trailing partial unit
                if (lane < rem) { const int k = nfull + lane; if (!(nfull == 0 && k < lead)) gbase[k] = ob[k]; }
                return;
            }
            __syncwarp();
            const int rem = cur - nfull;
            uint8_t keep = 0;
            if (lane < rem) keep = ob[nfull + lane];
            __syncwarp();
            if (lane < rem) ob[lane] = keep;
            gbase += nfull; cur = rem; lead = 0;
            __syncwarp();
        };
        for (int64_t base0 = 0; base0 < n; base0 += 256) {
            int32_t idv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { const int64_t i = base0 + 32 * k + lane; idv[k] = i < n ? ids[i] : 0; }
            uint32_t upad = 0;                                  // batches of this group that are all padding
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int64_t i = base0 + 32 * k + lane;
                if (__all_sync(FULL_MASK, idv[k] == T.pad && i < n - 1)) upad |= 1u << k;
            }
#pragma unroll 1
            for (int kb = 0; kb < 8; kb++) {
                const int64_t base = base0 + 32 * kb;
                if (base >= n) break;
                // a run of batches that hold nothing but pad ids (and not the row's last id): whole units of the periodic text
                if (padL && (upad >> kb) & 1u) {
                    int rl = 1;
                    while (kb + rl < 8 && ((upad >> (kb + rl)) & 1u)) rl++;
                    const int L = (int)padL, total = 32 * L * rl;
                    if (cur + total > DEC_CAP) flush(cur & ~15, false);
                    const int s0 = cur, s1 = cur + total;
                    const int a0 = (s0 + 15) & ~15, a1 = s1 & ~15;                       // whole 16-byte units inside [s0, s1)
                    {
                        int ph = (a0 - s0 + lane * 16) % L;
                        const int step = 512 % L;
                        for (int u = a0 + lane * 16; u < a1; u += 512) {
                            *reinterpret_cast<uint4*>(ob + u) = *reinterpret_cast<const uint4*>(padtab[wib][ph]);
                            ph += step; if (ph >= L) ph -= L;
                        }
                    }
                    if (s0 + lane < a0) ob[s0 + lane] = (uint8_t)(padP >> (8 * (lane % L)));
                    if (a1 + lane < s1) ob[a1 + lane] = (uint8_t)(padP >> (8 * ((a1 - s0 + lane) % L)));
                    cur += total;
                    kb += rl - 1;
                    continue;
                }
                const int64_t i = base + lane;
                int32_t myid = idv[0];
#pragma unroll
                for (int k = 1; k < 8; k++) if (kb == k) myid = idv[k];
                uint32_t off = 0, len = 0;
                if (i < n) dec_form(T, myid, i == n - 1, &off, &len);
                const uint8_t* src = T.form_blob + off;
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
                const uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
                const bool big = __any_sync(FULL_MASK, len > (uint32_t)DEC_MAXFORM);
                if (!big) {
                    if (cur + (int)tot > DEC_CAP) flush(cur & ~15, false);
                    uint8_t* d = ob + cur + (incl - len);
                    for (uint32_t k0 = 0; k0 < len; k0 += 8) {              // forms are 8-byte aligned in the blob: one load per 8 bytes
                        uint64_t v = *reinterpret_cast<const uint64_t*>(src + k0);
                        const uint32_t nb = len - k0 < 8u ? len - k0 : 8u;
#pragma unroll
                        for (uint32_t k = 0; k < 8; k++) if (k < nb) d[k0 + k] = (uint8_t)(v >> (8 * k));
                    }
                    cur += (int)tot;
                } else {                                        // rare: a very long vocab entry -> write this batch directly
                    if (cur > lead || lead) { /* push out what is staged so far, bytewise tail included */
                        __syncwarp();
                        for (int k = lead + lane; k < cur; k += 32) gbase[k] = ob[k];
                    }
                    uint8_t* d = gbase + cur + (incl - len);
                    for (uint32_t k = 0; k < len; k++) d[k] = src[k];
                    // restart the staging buffer at the new position
                    const int64_t gpos = (gbase - A.out) + cur + (int64_t)tot;
                    gbase = A.out + (gpos & ~(int64_t)15);
                    lead = (int)(gpos & 15);
                    cur = lead;
                    __syncwarp();
                }
            }
        }
        flush(0, true);
        __syncwarp();
    }
}

// get_atttention_mask helper on a flat id list
__global__ void k_mask_flat(const int32_t* ids, int64_t n, int32_t pad, uint8_t* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = ids[i] != pad;
}

// ---- word cache housekeeping -------------------------------------------------------------------------
// Decide whether the cache can take the worst case of the next chunk; if not, schedule a reset.
__global__ void k_cache_guard(WordCache C, unsigned long long need_slots, unsigned long long need_keys, unsigned long long need_toks, int force) {
    pdl_wait(); pdl_trigger();
    unsigned long long* c = C.ctr;
    const unsigned long long cap = (unsigned long long)C.mask + 1;
    const bool reset = force || (c[C_SLOTS] + need_slots) * 2 > cap || c[C_KEYS] + need_keys > C.key_cap || c[C_TOKS] + need_toks > C.tok_cap;
    c[C_RESET] = reset;
    if (reset) { c[C_SLOTS] = 0; c[C_KEYS] = 0; c[C_TOKS] = 0; }
    c[C_PENDING] = 0; c[C_REDO] = 0; c[C_FIX] = 0; c[C_FLATFIX_A] = 0; c[C_FLATFIX_B] = 0;
}
// ... and two optional word arrays to zero on the way (the document-start bitmaps of the byte-parallel pipeline)
__global__ void k_cache_clear(WordCache C, uint32_t* z0, uint64_t n0, uint32_t* z1, uint64_t n1) {
    pdl_wait(); pdl_trigger();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n0; i += nth) z0[i] = 0u;
    for (uint64_t i = tid; i < n1; i += nth) z1[i] = 0u;
    if (!C.ctr[C_RESET]) return;
    uint4* p = reinterpret_cast<uint4*>(C.slots);
    const uint64_t n = ((uint64_t)C.mask + 1) * 2;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t i = tid; i < n; i += nth) p[i] = z;
}
__global__ void k_reset_lists(WordCache C) { C.ctr[C_PENDING] = 0; C.ctr[C_REDO] = 0; C.ctr[C_FIX] = 0; C.ctr[C_FLATFIX_A] = 0; C.ctr[C_FLATFIX_B] = 0; }

}  // namespace gzt
