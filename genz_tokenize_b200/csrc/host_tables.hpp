// host_tables.hpp -- load vocab.txt / bpe.codes with the reference's exact dict semantics and
// compile them into the integer tables the CUDA kernels use.
//
// Reference behaviour reproduced (file:line are /root/reference/genz_tokenize/tokenize.py):
//   :31-37  encoder starts as {pad:0,bos:1,eos:2,mask:3,unk:4} (a dict literal: equal strings collapse)
//   :44-51  add_vocab_file: text mode (strict UTF-8, universal newlines), readlines(), strip(),
//           idx = rfind(' '), word = line[:idx] (idx == -1 drops the last code point),
//           encoder[word] = len(encoder)  (duplicate words re-assign without growing the dict)
//   :40     decoder = {v:k for k,v in encoder.items()} (last key in insertion order wins an id)
//   :53-57  add_bpe_file: read().split('\n')[:-1], tuple(line.split()), rank = line index
//           (later duplicates win; only 2-field lines can ever match)
//   :62-101 bpe(): symbols are STRINGS; here every string that can occur as a symbol gets an
//           integer id, and (symL,symR) -> (rank, sym(L+R)) is the pair table.
//   :99-100,:120-121 piece -> id: non-final symbol S is the token S+"@@", the final one is
//           S[:-4]; both are looked up once per symbol here (id_cont / id_fin).
//   :137-139 decode: ' '.join(pieces).replace('@@ ','') is a pure concatenation of per-id forms
//           mid(id) = replace(piece+" ") and last(id) = replace(piece) (proof in DESIGN.md).
#pragma once
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace gzt {

static const uint32_t SYM_NONE = 0xFFFFFFFFu;

inline bool is_space_cp(uint32_t c) {  // str.isspace() / re \s: 29 code points
    if (c <= 0x20) return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20);
    if (c == 0x85 || c == 0xA0 || c == 0x1680) return true;
    if (c >= 0x2000 && c <= 0x200A) return true;
    return c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

// Strict UTF-8 decode of one code point; returns byte length or 0.
inline int utf8_decode(const uint8_t* p, size_t avail, uint32_t* cp) {
    uint8_t b = p[0];
    if (b < 0x80) { *cp = b; return 1; }
    int n; uint32_t c;
    if ((b & 0xE0) == 0xC0) { n = 2; c = b & 0x1F; }
    else if ((b & 0xF0) == 0xE0) { n = 3; c = b & 0x0F; }
    else if ((b & 0xF8) == 0xF0) { n = 4; c = b & 0x07; }
    else return 0;
    if ((size_t)n > avail) return 0;
    for (int i = 1; i < n; i++) {
        if ((p[i] & 0xC0) != 0x80) return 0;
        c = (c << 6) | (p[i] & 0x3F);
    }
    if ((n == 2 && c < 0x80) || (n == 3 && (c < 0x800 || (c >= 0xD800 && c <= 0xDFFF))) ||
        (n == 4 && (c < 0x10000 || c > 0x10FFFF)))
        return 0;
    *cp = c;
    return n;
}

struct PairEntry { uint32_t l, r, rank, merged; };
struct CpEntry { uint32_t cp, sym_mid, sym_fin, pad; };

struct HostTables {
    // ---- dictionaries as the reference sees them
    std::vector<std::string> enc_keys;            // insertion order
    std::vector<int32_t> enc_vals;
    std::unordered_map<std::string, int64_t> enc_index;   // key -> position in enc_keys
    std::vector<int64_t> decoder;                 // id -> position in enc_keys, -1 absent
    std::vector<std::string> merge_lines;         // line i has rank i
    std::unordered_map<std::string, int32_t> ranks2;      // pair_key(L,R) -> rank for 2-field lines
    std::string special[5];                       // pad,bos,eos,mask,unk
    int32_t special_id[5];

    // ---- compiled integer tables
    std::vector<std::string> sym_str;             // symbol id -> string
    std::unordered_map<std::string, uint32_t> sym_index;
    std::vector<int32_t> id_cont, id_fin;         // per symbol
    std::vector<uint32_t> sym_ncp;                // code points per symbol (without "</w>")
    std::vector<PairEntry> pair_slots;            // open addressing, l == SYM_NONE empty
    uint32_t pair_mask = 0;
    std::vector<CpEntry> cp_slots;                // open addressing, cp == SYM_NONE empty
    uint32_t cp_mask = 0;
    // decode forms: blob + (offset,len) per id for "followed by another piece" / "last piece";
    // entry n_ids is the form of any id outside the decoder (the unk STRING, :123-124)
    std::vector<uint8_t> form_blob;
    std::vector<uint32_t> mid_off, mid_len, last_off, last_len;
    std::vector<uint32_t> mid_desc, last_desc;    // (offset/8) << 8 | min(len,255): one load per id on the device
    struct FastForm { uint32_t w[4]; };           // a "followed by another piece" form of up to 15 bytes with its length in the top byte (255: see the blob)
    std::vector<FastForm> mid_fast;               // the form and its length in ONE 16-byte load
    uint32_t max_form = 0;                        // longest decode form in bytes
    int64_t n_ids = 0;

    std::string err;
    int err_code = 0;

    static inline uint64_t mix64(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
        return x;
    }
    static inline uint32_t pair_hash(uint32_t l, uint32_t r) { return (uint32_t)mix64(((uint64_t)l << 32) | r); }
    static inline uint32_t cp_hash(uint32_t c) { return (c * 0x9E3779B1u) ^ (c >> 15); }

    void enc_set(const std::string& k, int32_t v) {
        auto it = enc_index.find(k);
        if (it != enc_index.end()) { enc_vals[it->second] = v; return; }
        enc_index.emplace(k, (int64_t)enc_keys.size());
        enc_keys.push_back(k);
        enc_vals.push_back(v);
    }
    int32_t enc_get(const std::string& k, int32_t dflt) const {
        auto it = enc_index.find(k);
        return it == enc_index.end() ? dflt : enc_vals[it->second];
    }

    // Python text-mode read: strict UTF-8 + universal newlines
    bool read_text(const char* path, std::string* out) {
        FILE* f = fopen(path, "rb");
        if (!f) { err = std::string("FileNotFoundError: [Errno 2] No such file or directory: '") + path + "'"; err_code = -2; return false; }
        std::string raw;
        char buf[1 << 16];
        size_t r;
        while ((r = fread(buf, 1, sizeof buf, f)) > 0) raw.append(buf, r);
        fclose(f);
        out->clear();
        out->reserve(raw.size());
        const uint8_t* p = (const uint8_t*)raw.data();
        for (size_t i = 0; i < raw.size();) {
            uint32_t cp = 0;
            int l = utf8_decode(p + i, raw.size() - i, &cp);
            if (l == 0) {
                char m[256];
                snprintf(m, sizeof m, "UnicodeDecodeError: 'utf-8' codec can't decode byte 0x%02x in position %zu (%s)", p[i], i, path);
                err = m; err_code = -3;
                return false;
            }
            if (cp == '\r') {
                out->push_back('\n');
                i++;
                if (i < raw.size() && p[i] == '\n') i++;
            } else {
                out->append((const char*)p + i, (size_t)l);
                i += (size_t)l;
            }
        }
        return true;
    }

    static void strip_ws(const uint8_t* p, size_t n, size_t* b, size_t* e) {
        size_t s = 0, t = n;
        while (s < t) {
            uint32_t cp = 0; int l = utf8_decode(p + s, t - s, &cp);
            if (l == 0 || !is_space_cp(cp)) break;
            s += (size_t)l;
        }
        while (t > s) {
            size_t k = t - 1;
            while (k > s && (p[k] & 0xC0) == 0x80) k--;
            uint32_t cp = 0; int l = utf8_decode(p + k, t - k, &cp);
            if (l == 0 || !is_space_cp(cp)) break;
            t = k;
        }
        *b = s; *e = t;
    }

    bool load_vocab(const char* path) {
        std::string txt;
        if (!read_text(path, &txt)) return false;
        const uint8_t* p = (const uint8_t*)txt.data();
        size_t i = 0, N = txt.size();
        while (i < N) {
            size_t j = i;
            while (j < N && p[j] != '\n') j++;
            size_t b, e;
            strip_ws(p + i, j - i, &b, &e);
            const uint8_t* line = p + i + b;
            size_t n = e - b, wl;
            size_t k = n;
            while (k > 0 && line[k - 1] != ' ') k--;
            if (k > 0) wl = k - 1;
            else { wl = n; if (wl > 0) { wl--; while (wl > 0 && (line[wl] & 0xC0) == 0x80) wl--; } }
            enc_set(std::string((const char*)line, wl), (int32_t)enc_keys.size());
            i = j < N ? j + 1 : j;
        }
        return true;
    }

    static void split_fields(const std::string& line, std::vector<std::string>* out) {
        out->clear();
        const uint8_t* p = (const uint8_t*)line.data();
        size_t k = 0, n = line.size();
        while (k < n) {
            uint32_t cp = 0; int l = utf8_decode(p + k, n - k, &cp);
            if (l == 0) l = 1;
            if (is_space_cp(cp)) { k += (size_t)l; continue; }
            size_t s = k;
            while (k < n) {
                l = utf8_decode(p + k, n - k, &cp);
                if (l == 0) l = 1;
                if (is_space_cp(cp)) break;
                k += (size_t)l;
            }
            out->emplace_back(line, s, k - s);
        }
    }

    bool load_merges(const char* path) {
        std::string txt;
        if (!read_text(path, &txt)) return false;
        size_t i = 0, N = txt.size();
        while (i < N) {
            size_t j = txt.find('\n', i);
            if (j == std::string::npos) break;     // [:-1] drops the unterminated tail
            merge_lines.emplace_back(txt, i, j - i);
            i = j + 1;
        }
        return true;
    }

    static std::string pair_key(const std::string& l, const std::string& r) {   // unambiguous even with '\0' inside
        return std::to_string(l.size()) + ":" + l + r;
    }

    uint32_t sym_of(const std::string& s) {
        auto it = sym_index.find(s);
        if (it != sym_index.end()) return it->second;
        uint32_t id = (uint32_t)sym_str.size();
        sym_index.emplace(s, id);
        sym_str.push_back(s);
        return id;
    }

    static void append_cp_set(const std::string& s, std::vector<uint32_t>* cps) {
        const uint8_t* p = (const uint8_t*)s.data();
        for (size_t i = 0; i < s.size();) {
            uint32_t cp = 0; int l = utf8_decode(p + i, s.size() - i, &cp);
            if (l == 0) { i++; continue; }
            cps->push_back(cp);
            i += (size_t)l;
        }
    }
    static std::string cp_to_utf8(uint32_t c) {
        std::string s;
        if (c < 0x80) s.push_back((char)c);
        else if (c < 0x800) { s.push_back((char)(0xC0 | (c >> 6))); s.push_back((char)(0x80 | (c & 0x3F))); }
        else if (c < 0x10000) { s.push_back((char)(0xE0 | (c >> 12))); s.push_back((char)(0x80 | ((c >> 6) & 0x3F))); s.push_back((char)(0x80 | (c & 0x3F))); }
        else { s.push_back((char)(0xF0 | (c >> 18))); s.push_back((char)(0x80 | ((c >> 12) & 0x3F))); s.push_back((char)(0x80 | ((c >> 6) & 0x3F))); s.push_back((char)(0x80 | (c & 0x3F))); }
        return s;
    }
    static std::string replace_cont(const std::string& s) {   // str.replace('@@ ', '')
        std::string o;
        for (size_t i = 0; i < s.size();) {
            if (i + 3 <= s.size() && s[i] == '@' && s[i + 1] == '@' && s[i + 2] == ' ') i += 3;
            else o.push_back(s[i++]);
        }
        return o;
    }

    bool build(const char* vocab_path, const char* bpe_path, const char* const specials[5]) {
        static const char* defaults[5] = {"<pad>", "<s>", "</s>", "<mask>", "<unk>"};
        for (int i = 0; i < 5; i++) {
            special[i] = (specials && specials[i]) ? specials[i] : defaults[i];
            enc_set(special[i], i);
        }
        if (!load_vocab(vocab_path)) return false;
        if (!load_merges(bpe_path)) return false;
        for (int i = 0; i < 5; i++) special_id[i] = enc_get(special[i], -1);
        int32_t unk = special_id[4];

        int32_t maxid = -1;
        for (int32_t v : enc_vals) if (v > maxid) maxid = v;
        n_ids = (int64_t)maxid + 1;
        decoder.assign((size_t)n_ids, -1);
        for (size_t e = 0; e < enc_keys.size(); e++) decoder[(size_t)enc_vals[e]] = (int64_t)e;

        // ---- symbols and the pair table
        std::vector<std::string> f;
        std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> pairs;   // (l,r) -> (rank, merged)
        std::vector<uint32_t> cps;
        for (size_t r = 0; r < merge_lines.size(); r++) {
            split_fields(merge_lines[r], &f);
            if (f.size() != 2) continue;
            ranks2[pair_key(f[0], f[1])] = (int32_t)r;
            uint32_t l = sym_of(f[0]), rr = sym_of(f[1]), m = sym_of(f[0] + f[1]);
            pairs[((uint64_t)l << 32) | rr] = {(uint32_t)r, m};
            append_cp_set(f[0], &cps);
            append_cp_set(f[1], &cps);
        }
        for (const std::string& k : enc_keys) append_cp_set(k, &cps);
        // every code point seen anywhere gets both of its initial symbols ("c" and "c</w>")
        std::unordered_map<uint32_t, std::pair<uint32_t, uint32_t>> cpmap;
        for (uint32_t c : cps) {
            if (cpmap.count(c)) continue;
            std::string u = cp_to_utf8(c);
            cpmap[c] = {sym_of(u), sym_of(u + "</w>")};
        }
        size_t ns = sym_str.size();
        id_cont.resize(ns); id_fin.resize(ns); sym_ncp.resize(ns);
        for (size_t s = 0; s < ns; s++) {
            const std::string& S = sym_str[s];
            id_cont[s] = enc_get(S + "@@", unk);
            bool fin = S.size() >= 4 && S.compare(S.size() - 4, 4, "</w>") == 0;
            id_fin[s] = fin ? enc_get(S.substr(0, S.size() - 4), unk) : unk;
            std::vector<uint32_t> t;
            append_cp_set(S, &t);
            sym_ncp[s] = (uint32_t)t.size();
        }
        size_t cap = 64;
        while (cap < pairs.size() * 2 + 2) cap <<= 1;
        pair_slots.assign(cap, PairEntry{SYM_NONE, SYM_NONE, 0, 0});
        pair_mask = (uint32_t)cap - 1;
        for (auto& kv : pairs) {
            uint32_t l = (uint32_t)(kv.first >> 32), r = (uint32_t)kv.first;
            uint32_t i = pair_hash(l, r) & pair_mask;
            while (pair_slots[i].l != SYM_NONE) i = (i + 1) & pair_mask;
            pair_slots[i] = PairEntry{l, r, kv.second.first, kv.second.second};
        }
        cap = 64;
        while (cap < cpmap.size() * 2 + 2) cap <<= 1;
        cp_slots.assign(cap, CpEntry{SYM_NONE, SYM_NONE, SYM_NONE, 0});
        cp_mask = (uint32_t)cap - 1;
        for (auto& kv : cpmap) {
            uint32_t i = cp_hash(kv.first) & cp_mask;
            while (cp_slots[i].cp != SYM_NONE) i = (i + 1) & cp_mask;
            cp_slots[i] = CpEntry{kv.first, kv.second.first, kv.second.second, 0};
        }

        // ---- decode forms (each form starts on an 8-byte boundary so that the kernel can fetch it with 8-byte loads)
        mid_off.resize((size_t)n_ids + 1); mid_len.resize((size_t)n_ids + 1);
        last_off.resize((size_t)n_ids + 1); last_len.resize((size_t)n_ids + 1);
        auto put = [&](const std::string& f, uint32_t* off, uint32_t* len) {
            while (form_blob.size() % 8) form_blob.push_back(0);
            *off = (uint32_t)form_blob.size(); *len = (uint32_t)f.size();
            form_blob.insert(form_blob.end(), f.begin(), f.end());
        };
        for (int64_t id = 0; id <= n_ids; id++) {
            const std::string& piece = (id < n_ids && decoder[(size_t)id] >= 0) ? enc_keys[(size_t)decoder[(size_t)id]] : special[4];
            put(replace_cont(piece + " "), &mid_off[(size_t)id], &mid_len[(size_t)id]);
            put(replace_cont(piece), &last_off[(size_t)id], &last_len[(size_t)id]);
        }
        // one 32-bit descriptor per form: (offset / 8) << 8 | min(len, 255)
        mid_desc.resize((size_t)n_ids + 1); last_desc.resize((size_t)n_ids + 1);
        max_form = 0;
        for (size_t i = 0; i < mid_len.size(); i++) max_form = std::max(max_form, std::max(mid_len[i], last_len[i]));
        for (int64_t id = 0; id <= n_ids; id++) {
            mid_desc[(size_t)id] = ((mid_off[(size_t)id] / 8) << 8) | std::min<uint32_t>(mid_len[(size_t)id], 255u);
            last_desc[(size_t)id] = ((last_off[(size_t)id] / 8) << 8) | std::min<uint32_t>(last_len[(size_t)id], 255u);
        }
        mid_fast.assign((size_t)n_ids + 1, FastForm{{0, 0, 0, 0}});
        for (int64_t id = 0; id <= n_ids; id++) {
            FastForm& f = mid_fast[(size_t)id];
            const uint32_t len = mid_len[(size_t)id];
            if (len > 15) { f.w[3] = 255u << 24; continue; }
            uint8_t b[16] = {0};
            std::memcpy(b, form_blob.data() + mid_off[(size_t)id], len);
            b[15] = (uint8_t)len;
            std::memcpy(f.w, b, 16);
        }
        while (form_blob.size() % 16) form_blob.push_back(0);
        return true;
    }

    int32_t rank_get(const std::string& l, const std::string& r) const {
        auto it = ranks2.find(pair_key(l, r));
        return it == ranks2.end() ? -1 : it->second;
    }
};

}  // namespace gzt
