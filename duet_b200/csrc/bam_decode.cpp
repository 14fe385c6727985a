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
#include <algorithm>
#include <new>
#include <vector>

#include "../../include/duet_b200.h"

extern "C" void duet_hash_names(const char *buf, const int64_t *off, int64_t n, uint64_t *lo, uint64_t *hi);

namespace {

inline uint32_t rd32(const unsigned char *p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const unsigned char *p) { uint16_t v; std::memcpy(&v, p, 2); return v; }

// ---- BGZF ---------------------------------------------------------------------------------------
std::atomic<int> g_inflate_threads{1};                        // duet_set_decode_threads

constexpr uint32_t kBgzfMaxIsize = 1u << 16;                  // a BGZF block inflates to at most 64 KiB (SAMv1 4.1)
constexpr size_t kBatchBytes = 16u << 20;                     // inflated bytes walked per batch: bounds the memory

struct Block { int64_t off; int32_t bsize; uint32_t isize; };

// The header of the BGZF block at `pos`: total size and inflated size.  Every length is checked against the
// end of the input before it is used.
int read_block_header(const unsigned char *data, int64_t len, int64_t pos, Block *out) {
    if (len - pos < 18 || data[pos] != 31 || data[pos + 1] != 139 || data[pos + 2] != 8 || !(data[pos + 3] & 4))
        return DUET_DECODE_ERR_FORMAT;
    const int xlen = rd16(data + pos + 10);
    if (pos + 12 + xlen > len) return DUET_DECODE_ERR_FORMAT;
    int64_t x = pos + 12;
    const int64_t xend = x + xlen;
    int bsize = -1;
    while (x + 4 <= xend) {
        const int slen = rd16(data + x + 2);
        if (x + 4 + slen > xend) return DUET_DECODE_ERR_FORMAT;
        if (data[x] == 'B' && data[x + 1] == 'C' && slen == 2) bsize = rd16(data + x + 4) + 1;
        x += 4 + slen;
    }
    if (bsize < 0 || pos + bsize > len || bsize < xlen + 20) return DUET_DECODE_ERR_FORMAT;
    const uint32_t isize = rd32(data + pos + bsize - 4);
    if (isize > kBgzfMaxIsize) return DUET_DECODE_ERR_FORMAT;   // a forged ISIZE must not size a buffer
    out->off = pos; out->bsize = bsize; out->isize = isize;
    return DUET_OK;
}

// one BGZF block (a raw-deflate member with its own ISIZE) -> its slice of the batch buffer
bool inflate_block(const unsigned char *data, const Block &b, unsigned char *dst) {
    if (b.isize == 0) return true;
    const unsigned char *blk = data + b.off;
    const int xlen = rd16(blk + 10);
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<unsigned char *>(blk + 12 + xlen);
    zs.avail_in = (uInt)(b.bsize - xlen - 20);
    zs.next_out = dst;
    zs.avail_out = b.isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.avail_out == 0;
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

// ---- the record walk, fed batch by batch ------------------------------------------------------------
// State of one file's walk: where in the BAM layout the stream is (magic, header text, reference list,
// alignments), what has been kept so far.  feed() consumes whole items and reports how far it got; the
// caller carries the rest over to the next batch, so memory is bounded by a batch plus one record.
struct BamWalker {
    enum { MAGIC, TEXT, NREF, REFS, RECORDS } state = MAGIC;
    uint64_t skip = 0;                 // bytes of header text still to pass over
    uint32_t refs_left = 0;
    int64_t rec = 0;                   // alignments seen
    int err = DUET_OK;
    std::vector<uint64_t> keys;
    std::vector<duet_read_tag> tags;
    // scratch of the general (text) path
    std::vector<const unsigned char *> aux;
    std::vector<Tok> toks;
    std::string tail, piece;

    bool keep(const unsigned char *name, size_t nl, long long hp, long long pc, long long ps) {
        if (hp < 0 || hp > 255 || pc < INT32_MIN || pc > INT32_MAX || ps < INT32_MIN || ps > INT32_MAX) { err = DUET_DECODE_ERR_RANGE; return false; }
        for (size_t i = 0; i < nl; ++i) if (name[i] >= 0x80) { err = DUET_DECODE_ERR_ASCII; return false; }
        const int64_t off[2] = {0, (int64_t)nl};
        uint64_t lo, hi;
        duet_hash_names(reinterpret_cast<const char *>(name), off, 1, &lo, &hi);
        duet_read_tag t;
        std::memset(&t, 0, sizeof(t));
        t.hp = (uint8_t)hp; t.pc = (int32_t)pc; t.ps = (int32_t)ps; t.chk = (uint32_t)hi;
        keys.push_back(lo);
        tags.push_back(t);
        return true;
    }

    // one complete alignment record [r, rend): the reference's rule on its last three text tokens
    bool record(const unsigned char *r, const unsigned char *rend) {
        const unsigned l_read_name = r[8];
        const unsigned n_cigar = rd16(r + 12);
        const uint32_t l_seq = rd32(r + 16);
        const int32_t tlen = (int32_t)rd32(r + 28);
        const unsigned char *name = r + 32;
        const unsigned char *cigar = name + l_read_name;
        const unsigned char *seq = cigar + 4ull * n_cigar;
        const unsigned char *qual = seq + ((uint64_t)l_seq + 1) / 2;
        const unsigned char *ax = qual + l_seq;
        if (ax > rend || ax < r || l_read_name == 0) { err = DUET_DECODE_ERR_FORMAT; return false; }
        aux.clear();
        for (const unsigned char *q = ax; q < rend;) {
            const size_t sz = aux_size(q, rend);
            if (!sz) { err = DUET_DECODE_ERR_FORMAT; return false; }
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
            if (aux[na - 2][0] == 'P' && aux[na - 2][1] == 'C')      // "PC:i:" can only sit at the start of "XX:i:<digits>"
                return keep(name, l_read_name - 1, int_of(aux[na - 3]), int_of(aux[na - 2]), int_of(aux[na - 1]));
            return true;
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
        for (char c : tail) if ((unsigned char)c >= 0x80) { err = DUET_DECODE_ERR_ASCII; return false; }
        tokenize(tail, toks);
        const size_t nt = toks.size();
        if (nt < 3) { err = DUET_DECODE_ERR_FORMAT; return false; }
        const Tok t3 = toks[nt - 3], t2 = toks[nt - 2], t1 = toks[nt - 1];
        if (tail.substr(t2.b, t2.e - t2.b).find("PC:i:") == std::string::npos) return true;
        const Tok tk[3] = {t3, t2, t1};
        long long v[3];
        for (int k = 0; k < 3; ++k) {
            bool ovf;
            const size_t b = tk[k].b + 5 <= tk[k].e ? tk[k].b + 5 : tk[k].e;
            if (!parse_int(tail.data() + b, tail.data() + tk[k].e, &v[k], &ovf)) { err = DUET_DECODE_ERR_VALUE; return false; }
            if (ovf) { err = DUET_DECODE_ERR_RANGE; return false; }
        }
        return keep(name, l_read_name - 1, v[0], v[1], v[2]);
    }

    // consume what is complete in [p, end); returns the first byte not consumed (nullptr on error: see err)
    const unsigned char *feed(const unsigned char *p, const unsigned char *end) {
        for (;;) {
            switch (state) {
                case MAGIC:
                    if (end - p < 8) return p;
                    if (std::memcmp(p, "BAM\1", 4) != 0) { err = DUET_DECODE_ERR_FORMAT; return nullptr; }
                    skip = rd32(p + 4);
                    p += 8;
                    state = TEXT;
                    break;
                case TEXT: {
                    const uint64_t n = std::min<uint64_t>(skip, (uint64_t)(end - p));
                    p += n; skip -= n;
                    if (skip) return p;
                    state = NREF;
                    break;
                }
                case NREF:
                    if (end - p < 4) return p;
                    refs_left = rd32(p);
                    p += 4;
                    state = REFS;
                    break;
                case REFS:
                    while (refs_left) {
                        if (end - p < 4) return p;
                        const uint64_t l_name = rd32(p);
                        if ((uint64_t)(end - p) < 4 + l_name + 4) {
                            if (l_name > (1u << 20)) { err = DUET_DECODE_ERR_FORMAT; return nullptr; }   // no reference name is a megabyte long
                            return p;
                        }
                        p += 4 + l_name + 4;
                        --refs_left;
                    }
                    state = RECORDS;
                    break;
                case RECORDS:
                    for (;;) {
                        if (end - p < 4) return p;
                        const uint32_t bs = rd32(p);
                        if (bs < 32 || bs > (1u << 30)) { err = DUET_DECODE_ERR_FORMAT; return nullptr; }
                        if ((uint64_t)(end - p) < 4ull + bs) return p;
                        if (!record(p + 4, p + 4 + bs)) return nullptr;
                        ++rec;
                        p += 4 + bs;
                    }
            }
        }
    }
};

// the whole file, batch by batch: at most kBatchBytes of inflated data (plus one record carried over) at a time
int walk_bam(const unsigned char *data, int64_t len, BamWalker &w) {
    std::vector<unsigned char> buf;                            // [carry | this batch's inflated blocks]
    std::vector<Block> batch;
    std::vector<size_t> dst;
    size_t carry = 0;
    int64_t pos = 0;
    while (pos < len) {
        batch.clear();
        dst.clear();
        size_t bytes = 0;
        while (pos < len && bytes < kBatchBytes) {
            Block b;
            const int rc = read_block_header(data, len, pos, &b);
            if (rc != DUET_OK) return rc;
            dst.push_back(bytes);
            bytes += b.isize;
            batch.push_back(b);
            pos += b.bsize;
        }
        buf.resize(carry + bytes);
        unsigned char *base = buf.data() + carry;
        const int n_thr = (int)std::min<size_t>((size_t)std::max(1, g_inflate_threads.load()), batch.size() / 16 + 1);
        if (n_thr > 1) {                                      // contiguous runs of blocks per worker
            std::atomic<bool> ok{true};
            std::vector<std::thread> pool;
            for (int t = 0; t < n_thr; ++t)
                pool.emplace_back([&, t] {
                    const size_t b0 = batch.size() * (size_t)t / n_thr, b1 = batch.size() * (size_t)(t + 1) / n_thr;
                    for (size_t i = b0; i < b1 && ok.load(std::memory_order_relaxed); ++i)
                        if (!inflate_block(data, batch[i], base + dst[i])) ok = false;
                });
            for (auto &th : pool) th.join();
            if (!ok) return DUET_DECODE_ERR_FORMAT;
        } else {
            for (size_t i = 0; i < batch.size(); ++i)
                if (!inflate_block(data, batch[i], base + dst[i])) return DUET_DECODE_ERR_FORMAT;
        }
        const unsigned char *end = buf.data() + buf.size();
        const unsigned char *stop = w.feed(buf.data(), end);
        if (!stop) return w.err;
        carry = (size_t)(end - stop);
        if (carry) std::memmove(buf.data(), stop, carry);
    }
    if (carry || w.state != BamWalker::RECORDS) return DUET_DECODE_ERR_FORMAT;     // truncated record / header
    return DUET_OK;
}

}  // namespace

// The kept rows of one haplotagged file, held by the library until the caller has made room for them
// (duet_rows_take copies them straight into the caller's -- page-locked -- column slices).
struct duet_rows {
    std::vector<uint64_t> keys;
    std::vector<duet_read_tag> tags;
};

extern "C" {

void duet_free(void *p) { std::free(p); }

int duet_set_decode_threads(int n) {
    return g_inflate_threads.exchange(n < 1 ? 1 : n);
}

int duet_decode_reads(const unsigned char *data, int64_t len, int kind, duet_rows **out, int64_t *n_rows,
                      int64_t *n_records, int64_t *err_at) {
    *out = nullptr; *n_rows = 0; *n_records = 0; *err_at = -1;
    duet_rows *rows = new (std::nothrow) duet_rows();
    if (!rows) return DUET_DECODE_ERR_CAPACITY;
    int rc;
    if (kind == DUET_READS_BAM) {
        BamWalker w;
        rc = walk_bam(data, len, w);
        *n_records = w.rec;
        if (rc != DUET_OK) *err_at = w.rec;
        rows->keys.swap(w.keys);
        rows->tags.swap(w.tags);
    } else {
        const int64_t cap = duet_count_lines(reinterpret_cast<const char *>(data), len);
        rows->keys.resize((size_t)cap);
        rows->tags.resize((size_t)cap);
        int64_t kept = 0;
        rc = duet_decode_sam_text(reinterpret_cast<const char *>(data), len, cap, rows->keys.data(), rows->tags.data(), &kept,
                                  n_records, err_at);
        rows->keys.resize((size_t)kept);
        rows->tags.resize((size_t)kept);
    }
    if (rc != DUET_OK) { delete rows; return rc; }
    *n_rows = (int64_t)rows->keys.size();
    *out = rows;
    return DUET_OK;
}

int duet_rows_take(duet_rows *rows, uint64_t *key_dst, duet_read_tag *tag_dst) {
    if (!rows) return DUET_ERR_INVALID;
    const size_t n = rows->keys.size();
    if (n) {
        std::memcpy(key_dst, rows->keys.data(), n * sizeof(uint64_t));
        std::memcpy(tag_dst, rows->tags.data(), n * sizeof(duet_read_tag));
    }
    delete rows;
    return DUET_OK;
}

void duet_rows_free(duet_rows *rows) { delete rows; }

int duet_decode_bam(const unsigned char *data, int64_t len, uint64_t **key_out, duet_read_tag **tag_out,
                    int64_t *n_rows, int64_t *n_records, int64_t *err_record) {
    *key_out = nullptr; *tag_out = nullptr; *n_rows = 0; *n_records = 0; *err_record = -1;
    BamWalker w;
    const int rc = walk_bam(data, len, w);
    *n_records = w.rec;
    if (rc != DUET_OK) { *err_record = w.rec; return rc; }
    const size_t n = w.keys.size();
    uint64_t *k = static_cast<uint64_t *>(std::malloc(n ? n * 8 : 8));
    duet_read_tag *t = static_cast<duet_read_tag *>(std::malloc(n ? n * sizeof(duet_read_tag) : 16));
    if (!k || !t) { std::free(k); std::free(t); return DUET_DECODE_ERR_CAPACITY; }
    if (n) { std::memcpy(k, w.keys.data(), n * 8); std::memcpy(t, w.tags.data(), n * sizeof(duet_read_tag)); }
    *key_out = k; *tag_out = t; *n_rows = (int64_t)n;
    return DUET_OK;
}

}  // extern "C"
