// Host-side decoders feeding the device path: read-name hashing and the haplotag text scan.
//
// duet_decode_sam_text restates what the reference does to every line of `samtools view`
// output (/root/reference/src/duet/sv_phasing_fn.py:25-29):
//     s = line.split();  if 'PC:i:' in s[-2]:  d[s[0]] = {hap: int(s[-3][5:]), ps: int(s[-1][5:]),
//                                                        pc: int(s[-2][5:])}
// i.e. the three tags are taken POSITIONALLY from the last three whitespace-separated fields
// and a row is kept when its second-to-last field contains "PC:i:".  Kept rows are emitted in
// file order (the device resolves duplicate QNAMEs to the last row).  Everything the reference
// would raise on (a line with fewer fields than it indexes, a non-integer tag value, a
// non-ASCII byte) is reported as an error code + line number so the Python layer can raise the
// same exception type.  Header lines ('@...') of a SAM file read directly are passed over: the
// reference only ever sees `samtools view` output, which does not contain them.
#include <cstdint>
#include <cstring>

#include "../../include/duet_b200.h"

namespace {

constexpr uint64_t C1 = 0x87C37B91114253D5ull, C2 = 0x4CF5AD432745937Full;
constexpr uint64_t SEED1 = 0x9E3779B97F4A7C15ull, SEED2 = 0xD1B54A32D192ED03ull;

inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t fmix(uint64_t k) {
    k ^= k >> 33; k *= 0xFF51AFD7ED558CCDull; k ^= k >> 33; k *= 0xC4CEB9FE1A85EC53ull; k ^= k >> 33;
    return k;
}
inline uint64_t load_le(const unsigned char *p, size_t n) {   // n <= 8 bytes, zero padded
    uint64_t v = 0;
    std::memcpy(&v, p, n);            // little-endian hosts only (x86-64 / aarch64)
    return v;
}

// duet_b200/namehash.py::hash128 -- keep bit-identical
void hash128(const unsigned char *s, size_t n, uint64_t *lo, uint64_t *hi) {
    uint64_t h1 = SEED1, h2 = SEED2;
    const size_t nblocks = n / 16;
    for (size_t b = 0; b < nblocks; ++b) {
        uint64_t k1 = load_le(s + 16 * b, 8), k2 = load_le(s + 16 * b + 8, 8);
        k1 *= C1; k1 = rotl(k1, 31); k1 *= C2; h1 ^= k1;
        h1 = rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52DCE729ull;
        k2 *= C2; k2 = rotl(k2, 33); k2 *= C1; h2 ^= k2;
        h2 = rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495AB5ull;
    }
    const unsigned char *t = s + 16 * nblocks;
    const size_t rem = n - 16 * nblocks;
    uint64_t k1 = load_le(t, rem < 8 ? rem : 8);
    uint64_t k2 = rem > 8 ? load_le(t + 8, rem - 8) : 0;
    k2 *= C2; k2 = rotl(k2, 33); k2 *= C1; h2 ^= k2;
    k1 *= C1; k1 = rotl(k1, 31); k1 *= C2; h1 ^= k1;
    h1 ^= (uint64_t)n; h2 ^= (uint64_t)n;
    h1 += h2; h2 += h1;
    h1 = fmix(h1); h2 = fmix(h2);
    h1 += h2; h2 += h1;
    if (h1 == 0xFFFFFFFFFFFFFFFFull) h1 = 0xFFFFFFFFFFFFFFFEull;   // the table's EMPTY sentinel
    *lo = h1; *hi = h2;
}

// str.split() separators for ASCII text
inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

// int(text) of CPython for ASCII input without surrounding blanks: [+-]digits with single '_' between digits
bool parse_int(const unsigned char *p, const unsigned char *e, long long *out, bool *overflow) {
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
        if (v > 100000000000000ull) *overflow = true; else v = v * 10 + (*p - '0');
    }
    if (prev_us) return false;
    *out = neg ? -(long long)v : (long long)v;
    return true;
}

}  // namespace

extern "C" {

void duet_hash_names(const char *buf, const int64_t *off, int64_t n, uint64_t *lo, uint64_t *hi) {
    for (int64_t i = 0; i < n; ++i)
        hash128(reinterpret_cast<const unsigned char *>(buf) + off[i], (size_t)(off[i + 1] - off[i]), lo + i, hi + i);
}

int64_t duet_hash_name_lists(const char *blob, int64_t len, int64_t n_lists, int64_t cap, int64_t *lens,
                             uint64_t *lo, uint64_t *hi) {
    // `blob` = the comma-separated name lists of n_lists records joined with '\n' (no trailing newline).
    // Splits like Python's str.split(','): "" is one empty name, "a,,b" has an empty name in the middle.
    const unsigned char *p = reinterpret_cast<const unsigned char *>(blob), *end = p + len;
    int64_t k = 0, rec = 0;
    if (n_lists <= 0) return 0;
    const unsigned char *start = p;
    int64_t in_rec = 0;
    for (;; ++p) {
        const bool at_end = p == end;
        if (at_end || *p == ',' || *p == '\n') {
            if (k >= cap) return -1;
            hash128(start, (size_t)(p - start), lo + k, hi + k);
            ++k; ++in_rec;
            start = p + 1;
            if (at_end || *p == '\n') {
                if (rec >= n_lists) return -1;
                lens[rec++] = in_rec;
                in_rec = 0;
                if (at_end) break;
            }
        }
    }
    return rec == n_lists ? k : -1;
}

void duet_pack_tags(int64_t n, const uint8_t *hp, const int32_t *ps, const int32_t *pc, const uint64_t *hi,
                    duet_read_tag *out) {
    for (int64_t i = 0; i < n; ++i) {
        duet_read_tag t;
        std::memset(&t, 0, sizeof(t));
        t.ps = ps[i]; t.pc = pc[i]; t.hp = hp[i];
        t.chk = hi ? (uint32_t)hi[i] : 0u;
        out[i] = t;
    }
}

int duet_decode_sam_text(const char *text, int64_t len, int64_t cap, uint64_t *key, duet_read_tag *tag,
                         int64_t *n_rows, int64_t *n_lines, int64_t *err_line) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(text);
    const unsigned char *end = p + len;
    int64_t rows = 0, line_no = 0;
    *n_rows = 0; *n_lines = 0; *err_line = -1;
    while (p < end) {
        const unsigned char *nl = static_cast<const unsigned char *>(std::memchr(p, '\n', (size_t)(end - p)));
        if (!nl) break;                                    // text after the last newline is dropped (:25 [:-1])
        if (*p == '@') { p = nl + 1; continue; }           // a SAM header line: `samtools view` (:25) does not print those
        // field scan: remember the first field and the last three
        const unsigned char *f0b = nullptr, *f0e = nullptr;
        const unsigned char *fb[3] = {nullptr, nullptr, nullptr}, *fe[3] = {nullptr, nullptr, nullptr};
        int nf = 0;
        const unsigned char *q = p;
        while (q < nl) {
            if (*q >= 0x80) { *err_line = line_no; return DUET_DECODE_ERR_ASCII; }
            if (is_space(*q)) { ++q; continue; }
            const unsigned char *b = q;
            while (q < nl && !is_space(*q)) {
                if (*q >= 0x80) { *err_line = line_no; return DUET_DECODE_ERR_ASCII; }
                ++q;
            }
            if (nf == 0) { f0b = b; f0e = q; }
            fb[0] = fb[1]; fe[0] = fe[1]; fb[1] = fb[2]; fe[1] = fe[2]; fb[2] = b; fe[2] = q;
            ++nf;
        }
        if (nf < 2) { *err_line = line_no; return DUET_DECODE_ERR_INDEX; }      // s[-2] raises IndexError
        // 'PC:i:' in s[-2]
        bool has = false;
        for (const unsigned char *c = fb[1]; c + 5 <= fe[1]; ++c)
            if (std::memcmp(c, "PC:i:", 5) == 0) { has = true; break; }
        if (has) {
            if (nf < 3) { *err_line = line_no; return DUET_DECODE_ERR_INDEX; }  // s[-3] raises IndexError
            long long v[3];
            for (int k = 0; k < 3; ++k) {
                bool ovf;
                const unsigned char *b = fb[k] + 5 <= fe[k] ? fb[k] + 5 : fe[k];
                if (!parse_int(b, fe[k], &v[k], &ovf)) { *err_line = line_no; return DUET_DECODE_ERR_VALUE; }
                if (ovf) { *err_line = line_no; return DUET_DECODE_ERR_RANGE; }
            }
            if (v[0] < 0 || v[0] > 255 || v[1] < INT32_MIN || v[1] > INT32_MAX || v[2] < INT32_MIN || v[2] > INT32_MAX) {
                *err_line = line_no; return DUET_DECODE_ERR_RANGE;
            }
            if (rows >= cap) { *err_line = line_no; return DUET_DECODE_ERR_CAPACITY; }
            uint64_t hi;
            hash128(f0b, (size_t)(f0e - f0b), key + rows, &hi);
            duet_read_tag t;
            std::memset(&t, 0, sizeof(t));
            t.hp = (uint8_t)v[0];
            t.pc = (int32_t)v[1];
            t.ps = (int32_t)v[2];
            t.chk = (uint32_t)hi;
            tag[rows] = t;
            ++rows;
        }
        ++line_no;
        p = nl + 1;
    }
    for (; p < end; ++p)                                   // .decode('ascii') covers the dropped tail too
        if (*p >= 0x80) { *err_line = line_no; return DUET_DECODE_ERR_ASCII; }
    *n_rows = rows;
    *n_lines = line_no;
    return DUET_OK;
}

int64_t duet_count_lines(const char *text, int64_t len) {
    int64_t n = 0;
    const char *p = text, *end = text + len;
    while (p < end) {
        const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        if (!nl) break;
        ++n;
        p = nl + 1;
    }
    return n;
}

}  // extern "C"
