// Native reader for the haplotagged per-contig BAMs: BGZF (zlib raw inflate) + BAM record walk,
// extracting only QNAME and the fields the reference looks at.
//
// The reference never sees BAM records: it reads `samtools view` TEXT and applies
//     s = line.split();  if 'PC:i:' in s[-2]:  d[s[0]] = {hap: int(s[-3][5:]), ps: int(s[-1][5:]), pc: int(s[-2][5:])}
// (/root/reference/src/duet/sv_phasing_fn.py:25-29) -- the last three WHITESPACE-separated tokens of
// the line, whatever they are.  To give the same answer on the same file this reader renders, for
// every record, just enough of the END of that text line (aux fields from the last one backwards, the
// way samtools prints them: TAG:TYPE:VALUE, every integer type as 'i') to know its last three tokens,
// and then applies the very same rule.  Aux strings containing blanks therefore shift the tokens
// exactly as they do for the reference.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/duet_b200.h"

extern "C" void duet_hash_names(const char *buf, const int64_t *off, int64_t n, uint64_t *lo, uint64_t *hi);

namespace {

inline uint32_t rd32(const unsigned char *p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const unsigned char *p) { uint16_t v; std::memcpy(&v, p, 2); return v; }

// ---- BGZF ---------------------------------------------------------------------------------------
std::atomic<int> g_inflate_threads{1};                        // duet_set_decode_threads

// one BGZF block (a raw-deflate member with its own ISIZE) -> its slice of the output
bool inflate_block(const unsigned char *blk, int64_t bsize, unsigned char *dst) {
    const int xlen = rd16(blk + 10);
    const uint32_t isize = rd32(blk + bsize - 4);
    if (isize == 0) return true;
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<unsigned char *>(blk + 12 + xlen);
    zs.avail_in = (uInt)(bsize - xlen - 20);
    zs.next_out = dst;
    zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.avail_out == 0;
}

int inflate_bgzf(const unsigned char *data, int64_t len, std::vector<unsigned char> &out) {
    int64_t pos = 0;
    size_t total = 0;
    // first pass: sizes (every block states its own compressed and uncompressed size, so the blocks can
    // be inflated independently, each into its own slice of the output)
    std::vector<std::pair<int64_t, int64_t>> blocks;          // (offset, block size)
    std::vector<size_t> dst_off;
    while (pos < len) {
        if (len - pos < 18 || data[pos] != 31 || data[pos + 1] != 139 || data[pos + 2] != 8 || !(data[pos + 3] & 4))
            return DUET_DECODE_ERR_FORMAT;
        const int xlen = rd16(data + pos + 10);
        int64_t x = pos + 12, xend = x + xlen;
        int bsize = -1;
        while (x + 4 <= xend) {
            const int slen = rd16(data + x + 2);
            if (data[x] == 'B' && data[x + 1] == 'C' && slen == 2) bsize = rd16(data + x + 4) + 1;
            x += 4 + slen;
        }
        if (bsize < 0 || pos + bsize > len || bsize < xlen + 20) return DUET_DECODE_ERR_FORMAT;
        dst_off.push_back(total);
        total += rd32(data + pos + bsize - 4);
        blocks.emplace_back(pos, bsize);
        pos += bsize;
    }
    out.resize(total);
    const int n_thr = (int)std::min<size_t>((size_t)std::max(1, g_inflate_threads.load()), blocks.size() / 16 + 1);
    if (n_thr > 1) {                                          // contiguous runs of blocks per worker
        std::atomic<bool> ok{true};
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; ++t)
            pool.emplace_back([&, t] {
                const size_t b0 = blocks.size() * (size_t)t / n_thr, b1 = blocks.size() * (size_t)(t + 1) / n_thr;
                for (size_t i = b0; i < b1 && ok.load(std::memory_order_relaxed); ++i)
                    if (!inflate_block(data + blocks[i].first, blocks[i].second, out.data() + dst_off[i])) ok = false;
            });
        for (auto &th : pool) th.join();
        return ok ? (int)DUET_OK : (int)DUET_DECODE_ERR_FORMAT;
    }
    for (size_t i = 0; i < blocks.size(); ++i)
        if (!inflate_block(data + blocks[i].first, blocks[i].second, out.data() + dst_off[i])) return DUET_DECODE_ERR_FORMAT;
    return DUET_OK;
}

// ---- text rendering of one aux field, as `samtools view` prints it ---------------------------------
// returns the number of bytes the field occupies in the record, 0 on malformed data
size_t aux_size(const unsigned char *p, const unsigned char *end) {
    if (end - p < 3) return 0;
    const char t = (char)p[2];
    const unsigned char *v = p + 3;
    switch (t) {
        case 'A': case 'c': case 'C': return end - v >= 1 ? 4 : 0;
        case 's': case 'S': return end - v >= 2 ? 5 : 0;
        case 'i': case 'I': case 'f': return end - v >= 4 ? 7 : 0;
        case 'Z': case 'H': {
            const void *z = std::memchr(v, 0, (size_t)(end - v));
            return z ? (size_t)((const unsigned char *)z - p) + 1 : 0;
        }
        case 'B': {
            if (end - v < 5) return 0;
            const char st = (char)v[0];
            const uint32_t n = rd32(v + 1);
            const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
            if (!es || (size_t)(end - v - 5) < es * n) return 0;
            return 3 + 5 + es * n;
        }
        default: return 0;
    }
}

void render_num(std::string &s, char t, const unsigned char *v) {
    char buf[48];
    switch (t) {
        case 'c': std::snprintf(buf, sizeof buf, "%d", (int)(int8_t)v[0]); break;
        case 'C': std::snprintf(buf, sizeof buf, "%u", (unsigned)v[0]); break;
        case 's': { int16_t x; std::memcpy(&x, v, 2); std::snprintf(buf, sizeof buf, "%d", (int)x); break; }
        case 'S': std::snprintf(buf, sizeof buf, "%u", (unsigned)rd16(v)); break;
        case 'i': { int32_t x; std::memcpy(&x, v, 4); std::snprintf(buf, sizeof buf, "%d", x); break; }
        case 'I': std::snprintf(buf, sizeof buf, "%u", rd32(v)); break;
        case 'f': { float x; std::memcpy(&x, v, 4); std::snprintf(buf, sizeof buf, "%g", x); break; }
        default: buf[0] = 0;
    }
    s += buf;
}

void render_aux(std::string &s, const unsigned char *p) {
    const char t = (char)p[2];
    const unsigned char *v = p + 3;
    s.push_back((char)p[0]); s.push_back((char)p[1]); s.push_back(':');
    switch (t) {
        case 'A': s += "A:"; s.push_back((char)v[0]); break;
        case 'c': case 'C': case 's': case 'S': case 'i': case 'I': s += "i:"; render_num(s, t, v); break;
        case 'f': s += "f:"; render_num(s, t, v); break;
        case 'Z': s += "Z:"; s += reinterpret_cast<const char *>(v); break;
        case 'H': s += "H:"; s += reinterpret_cast<const char *>(v); break;
        case 'B': {
            const char st = (char)v[0];
            const uint32_t n = rd32(v + 1);
            const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            s += "B:"; s.push_back(st);
            for (uint32_t i = 0; i < n; ++i) { s.push_back(','); render_num(s, st, v + 5 + es * i); }
            break;
        }
    }
}

inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

struct Tok { size_t b, e; };
// whitespace tokens of `s`, appended in order
void tokenize(const std::string &s, std::vector<Tok> &out) {
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && is_space((unsigned char)s[i])) ++i;
        if (i >= s.size()) break;
        const size_t b = i;
        while (i < s.size() && !is_space((unsigned char)s[i])) ++i;
        out.push_back(Tok{b, i});
    }
}

// int(text) of CPython (ASCII, no surrounding blanks): [+-]digits, single '_' between digits
bool parse_int(const char *p, const char *e, long long *out, bool *overflow) {
    *overflow = false;
    if (p == e) return false;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    if (p == e || *p < '0' || *p > '9') return false;
    unsigned long long v = 0;
    bool prev_us = false;
    for (; p < e; ++p) {
        if (*p == '_') { if (prev_us) return false; prev_us = true; continue; }
        if (*p < '0' || *p > '9') return false;
        prev_us = false;
        if (v > 100000000000000ull) *overflow = true; else v = v * 10 + (unsigned)(*p - '0');
    }
    if (prev_us) return false;
    *out = neg ? -(long long)v : (long long)v;
    return true;
}

}  // namespace

extern "C" {

void duet_free(void *p) { std::free(p); }

int duet_set_decode_threads(int n) {
    return g_inflate_threads.exchange(n < 1 ? 1 : n);
}

int duet_decode_bam(const unsigned char *data, int64_t len, uint64_t **key_out, duet_read_tag **tag_out,
                    int64_t *n_rows, int64_t *n_records, int64_t *err_record) {
    *key_out = nullptr; *tag_out = nullptr; *n_rows = 0; *n_records = 0; *err_record = -1;
    std::vector<unsigned char> raw;
    int rc = inflate_bgzf(data, len, raw);
    if (rc != DUET_OK) return rc;
    const unsigned char *p = raw.data(), *end = p + raw.size();
    if (end - p < 12 || std::memcmp(p, "BAM\1", 4) != 0) return DUET_DECODE_ERR_FORMAT;
    const uint32_t l_text = rd32(p + 4);
    p += 8;
    if ((size_t)(end - p) < (size_t)l_text + 4) return DUET_DECODE_ERR_FORMAT;
    p += l_text;
    const uint32_t n_ref = rd32(p);
    p += 4;
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (end - p < 4) return DUET_DECODE_ERR_FORMAT;
        const uint32_t l_name = rd32(p);
        if ((size_t)(end - p) < (size_t)l_name + 8) return DUET_DECODE_ERR_FORMAT;
        p += 4 + l_name + 4;
    }
    std::vector<uint64_t> keys;
    std::vector<duet_read_tag> tags;
    std::vector<const unsigned char *> aux;
    std::vector<Tok> toks;
    std::string tail, piece;
    int64_t rec = 0;
    auto bail = [&](int code) { *err_record = rec; return code; };
    while (p < end) {
        if (end - p < 4) return bail(DUET_DECODE_ERR_FORMAT);
        const uint32_t bs = rd32(p);
        const unsigned char *r = p + 4, *rend = r + bs;
        if (bs < 32 || rend > end) return bail(DUET_DECODE_ERR_FORMAT);
        const unsigned l_read_name = r[8];
        const unsigned n_cigar = rd16(r + 12);
        const uint32_t l_seq = rd32(r + 16);
        const int32_t tlen = (int32_t)rd32(r + 28);
        const unsigned char *name = r + 32;
        const unsigned char *cigar = name + l_read_name;
        const unsigned char *seq = cigar + 4ull * n_cigar;
        const unsigned char *qual = seq + (l_seq + 1) / 2;
        const unsigned char *ax = qual + l_seq;
        if (ax > rend || l_read_name == 0) return bail(DUET_DECODE_ERR_FORMAT);
        aux.clear();
        for (const unsigned char *q = ax; q < rend;) {
            const size_t sz = aux_size(q, rend);
            if (!sz) return bail(DUET_DECODE_ERR_FORMAT);
            aux.push_back(q);
            q += sz;
        }
        // fast path: the last three aux fields are integers -> they ARE the last three tokens
        const size_t na = aux.size();
        auto is_int = [](const unsigned char *q) { const char t = (char)q[2]; return t == 'c' || t == 'C' || t == 's' || t == 'S' || t == 'i' || t == 'I'; };
        auto int_of = [](const unsigned char *q) -> long long {
            const unsigned char *v = q + 3;
            switch ((char)q[2]) {
                case 'c': return (int8_t)v[0];
                case 'C': return v[0];
                case 's': { int16_t x; std::memcpy(&x, v, 2); return x; }
                case 'S': return rd16(v);
                case 'i': { int32_t x; std::memcpy(&x, v, 4); return x; }
                default: return rd32(v);
            }
        };
        if (na >= 3 && is_int(aux[na - 1]) && is_int(aux[na - 2]) && is_int(aux[na - 3])) {
            if (aux[na - 2][0] == 'P' && aux[na - 2][1] == 'C') {      // "PC:i:" can only sit at the start of "XX:i:<digits>"
                const long long v0 = int_of(aux[na - 3]), v1 = int_of(aux[na - 2]), v2 = int_of(aux[na - 1]);
                if (v0 < 0 || v0 > 255 || v1 < INT32_MIN || v1 > INT32_MAX || v2 < INT32_MIN || v2 > INT32_MAX)
                    return bail(DUET_DECODE_ERR_RANGE);
                const size_t nl = l_read_name - 1;
                for (size_t i = 0; i < nl; ++i) if (name[i] >= 0x80) return bail(DUET_DECODE_ERR_ASCII);
                const int64_t off[2] = {0, (int64_t)nl};
                uint64_t lo, hi;
                duet_hash_names(reinterpret_cast<const char *>(name), off, 1, &lo, &hi);
                duet_read_tag t;
                std::memset(&t, 0, sizeof(t));
                t.hp = (uint8_t)v0; t.pc = (int32_t)v1; t.ps = (int32_t)v2; t.chk = (uint32_t)hi;
                keys.push_back(lo);
                tags.push_back(t);
            }
            ++rec;
            p = rend;
            continue;
        }
        // general path: last three whitespace tokens of the text line, built from the end backwards
        tail.clear();
        toks.clear();
        int need = 3;
        int ai = (int)aux.size() - 1, mand = 0;                  // mandatory fields consumed from the end: QUAL, SEQ, TLEN
        std::vector<std::string> pieces;                         // reversed order
        while (need > 0) {
            piece.clear();
            if (ai >= 0) render_aux(piece, aux[ai--]);
            else if (mand == 0) {
                ++mand;
                if (l_seq == 0 || qual[0] == 0xFF) piece = "*";
                else for (uint32_t i = 0; i < l_seq; ++i) piece.push_back((char)(qual[i] + 33));
            } else if (mand == 1) {
                ++mand;
                if (l_seq == 0) piece = "*";
                else for (uint32_t i = 0; i < l_seq; ++i) piece.push_back("=ACMGRSVTWYHKDBN"[(seq[i >> 1] >> ((~i & 1) << 2)) & 15]);
            } else if (mand == 2) {
                ++mand;
                piece = std::to_string(tlen);
            } else break;                                        // earlier fields cannot be reached: >= 3 tokens by now
            std::vector<Tok> t;
            tokenize(piece, t);
            need -= (int)t.size();
            pieces.push_back(piece);
        }
        for (int i = (int)pieces.size() - 1; i >= 0; --i) { tail += pieces[i]; tail.push_back('\t'); }
        for (char c : tail) if ((unsigned char)c >= 0x80) return bail(DUET_DECODE_ERR_ASCII);
        tokenize(tail, toks);
        const size_t nt = toks.size();
        if (nt < 3) return bail(DUET_DECODE_ERR_FORMAT);
        const Tok t3 = toks[nt - 3], t2 = toks[nt - 2], t1 = toks[nt - 1];
        if (tail.substr(t2.b, t2.e - t2.b).find("PC:i:") != std::string::npos) {
            const Tok tk[3] = {t3, t2, t1};
            long long v[3];
            for (int k = 0; k < 3; ++k) {
                bool ovf;
                const size_t b = tk[k].b + 5 <= tk[k].e ? tk[k].b + 5 : tk[k].e;
                if (!parse_int(tail.data() + b, tail.data() + tk[k].e, &v[k], &ovf)) return bail(DUET_DECODE_ERR_VALUE);
                if (ovf) return bail(DUET_DECODE_ERR_RANGE);
            }
            if (v[0] < 0 || v[0] > 255 || v[1] < INT32_MIN || v[1] > INT32_MAX || v[2] < INT32_MIN || v[2] > INT32_MAX)
                return bail(DUET_DECODE_ERR_RANGE);
            size_t nl = l_read_name - 1;                          // QNAME is NUL terminated
            for (size_t i = 0; i < nl; ++i) if (name[i] >= 0x80) return bail(DUET_DECODE_ERR_ASCII);
            const int64_t off[2] = {0, (int64_t)nl};
            uint64_t lo, hi;
            duet_hash_names(reinterpret_cast<const char *>(name), off, 1, &lo, &hi);
            duet_read_tag t;
            std::memset(&t, 0, sizeof(t));
            t.hp = (uint8_t)v[0]; t.pc = (int32_t)v[1]; t.ps = (int32_t)v[2]; t.chk = (uint32_t)hi;
            keys.push_back(lo);
            tags.push_back(t);
        }
        ++rec;
        p = rend;
    }
    const size_t n = keys.size();
    uint64_t *k = static_cast<uint64_t *>(std::malloc(n ? n * 8 : 8));
    duet_read_tag *t = static_cast<duet_read_tag *>(std::malloc(n ? n * sizeof(duet_read_tag) : 16));
    if (!k || !t) { std::free(k); std::free(t); return DUET_DECODE_ERR_CAPACITY; }
    if (n) { std::memcpy(k, keys.data(), n * 8); std::memcpy(t, tags.data(), n * sizeof(duet_read_tag)); }
    *key_out = k; *tag_out = t; *n_rows = (int64_t)n; *n_records = rec;
    return DUET_OK;
}

}  // extern "C"
