// Native single-pass reader of the SV VCF (cuteSV / Sniffles2 / SVIM dialects): the fast path of
// duet_b200/read_file.py::parse_vcf, which restates /root/reference/src/duet/read_file.py:25-76.
//
// What the reference does per contig c of its contig list: keep the lines whose first whitespace token is
// 'c' or 'chr'+c (:30), and from each kept line take
//     [10] SVLEN   first INFO item containing 'SVLEN='  (missing or 'SVLEN=.' -> 0; 'SVLEN=>N' handled)   :34-36
//     [11] SVTYPE  first INFO item containing 'SVTYPE='                                                    :38
//     [12] support first INFO item containing SUPPORT= / SR= / RE=; prefix length decided by the contig's
//                  FIRST record (8 or 3)                                                                   :40-47
//     [13] names   first item containing RNAMES= / READS= (7 or 6), split on ','                            :48-55
//     [14..16]     GT and two counts from the sample column, layout decided by the first record            :56-76
// This reader makes ONE pass over the text with several threads (line ranges), hashes the support-read
// names in the same pass and hands back columns.  It only claims inputs it reproduces exactly: every kept
// record must be regular (the needles at the start of their INFO items, plain decimal integers, the same
// dialect as the contig's first record, every CHROM string claimed by one contig, ...).  Anything else
// returns DUET_DECODE_FALLBACK and the caller uses the general Python reader, which keeps the reference's
// behaviour -- exceptions included -- on irregular input.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/duet_b200.h"

extern "C" void duet_hash_names(const char *buf, const int64_t *off, int64_t n, uint64_t *lo, uint64_t *hi);

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

struct Span { const char *b, *e; size_t size() const { return (size_t)(e - b); } };

// plain decimal integer: [+-]digits, nothing else (Python's int() accepts more: that is fallback territory)
bool plain_int(Span s, long long *out) {
    const char *p = s.b;
    if (p == s.e) return false;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    if (p == s.e || s.e - p > 18) return false;
    long long v = 0;
    for (; p < s.e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    *out = neg ? -v : v;
    return true;
}
bool fits_i32(long long v) { return v >= INT32_MIN && v <= INT32_MAX; }

// earliest occurrence of any needle in [b, e); which needle it was
const char *find_first(Span info, const char *const *needles, const size_t *lens, int n, int *which) {
    const char *best = nullptr;
    for (int k = 0; k < n; ++k) {
        const char *p = info.b;
        const size_t L = lens[k];
        while ((size_t)(info.e - p) >= L) {
            p = static_cast<const char *>(std::memchr(p, needles[k][0], (size_t)(info.e - p) - L + 1));
            if (!p) break;
            if (std::memcmp(p, needles[k], L) == 0) { if (!best || p < best) { best = p; *which = k; } break; }
            ++p;
        }
    }
    return best;
}
// the ';'-separated item around position `at`
Span item_at(Span info, const char *at) {
    const char *a = at;
    while (a > info.b && a[-1] != ';') --a;
    const char *z = static_cast<const char *>(std::memchr(at, ';', (size_t)(info.e - at)));
    return Span{a, z ? z : info.e};
}

struct Rec {
    int32_t contig;
    int32_t pos, svlen, svread, refread;
    uint8_t gt_missing;
    uint8_t sup_cut, nm_cut, ad_mode;        // the dialect this record would imply if it were a contig's first
    uint8_t n_sample_fields;
    Span chrom, ref, alt, svtype, names;
};

struct Part {                                  // one thread's line range
    std::vector<Rec> recs;
    bool fallback = false;
};

struct Ctx {
    std::unordered_map<std::string, int> owner;     // CHROM string -> contig index
};

void parse_range(const Ctx &cx, const char *b, const char *e, Part &out) {
    static const char *const kSup[3] = {"SUPPORT=", "SR=", "RE="};
    static const size_t kSupLen[3] = {8, 3, 3};
    static const char *const kNm[2] = {"RNAMES=", "READS="};
    static const size_t kNmLen[2] = {7, 6};
    static const char *const kLen[1] = {"SVLEN="};
    static const size_t kLenLen[1] = {6};
    static const char *const kTyp[1] = {"SVTYPE="};
    static const size_t kTypLen[1] = {7};
    std::string key;
    const char *p = b;
    while (p < e) {
        const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(e - p)));
        const char *le = nl ? nl : e;
        // tokens 0..9
        Span tok[10];
        int nt = 0;
        const char *q = p;
        while (q < le && nt < 10) {
            while (q < le && is_space((unsigned char)*q)) ++q;
            if (q >= le) break;
            const char *s = q;
            while (q < le && !is_space((unsigned char)*q)) ++q;
            tok[nt++] = Span{s, q};
        }
        for (const char *c = p; c < le; ++c)                                       // text decoding and universal newlines
            if ((unsigned char)*c >= 0x80 || (*c == '\r' && c + 1 != le)) { out.fallback = true; return; }   // are Python's business
        p = nl ? nl + 1 : e;
        if (nt == 0) { out.fallback = true; return; }                             // blank line: the reference raises IndexError
        key.assign(tok[0].b, tok[0].size());
        auto it = cx.owner.find(key);
        if (it == cx.owner.end()) continue;                                       // header line or a contig not listed
        if (it->second < 0 || nt < 10) { out.fallback = true; return; }           // claimed by two contigs / short record
        Rec r;
        std::memset(&r, 0, sizeof(r));
        r.contig = it->second;
        r.chrom = tok[0]; r.ref = tok[3]; r.alt = tok[4];
        long long v;
        if (!plain_int(tok[1], &v) || !fits_i32(v)) { out.fallback = true; return; }
        r.pos = (int32_t)v;
        const Span info = tok[7];
        int which = 0;
        // SVLEN (missing -> 0)
        const char *at = find_first(info, kLen, kLenLen, 1, &which);
        if (at) {
            const Span it2 = item_at(info, at);
            if (it2.b != at) { out.fallback = true; return; }
            Span val{at + 6, it2.e};
            if (val.size() == 1 && *val.b == '.') v = 0;
            else {
                if (std::memchr(it2.b, '>', it2.size())) {
                    if (val.size() < 2 || *val.b != '>') { out.fallback = true; return; }
                    ++val.b;
                }
                if (!plain_int(val, &v)) { out.fallback = true; return; }
            }
            if (!fits_i32(v) || v == INT32_MIN) { out.fallback = true; return; }
            r.svlen = (int32_t)v;
        }
        // SVTYPE
        at = find_first(info, kTyp, kTypLen, 1, &which);
        if (!at) { out.fallback = true; return; }
        {
            const Span it2 = item_at(info, at);
            if (it2.b != at) { out.fallback = true; return; }
            r.svtype = Span{at + 7, it2.e};
        }
        // support count
        at = find_first(info, kSup, kSupLen, 3, &which);
        if (!at) { out.fallback = true; return; }
        {
            const Span it2 = item_at(info, at);
            if (it2.b != at) { out.fallback = true; return; }
            r.sup_cut = (uint8_t)kSupLen[which];
            if (!plain_int(Span{at + kSupLen[which], it2.e}, &v) || !fits_i32(v)) { out.fallback = true; return; }
            r.svread = (int32_t)v;
        }
        // read names
        at = find_first(info, kNm, kNmLen, 2, &which);
        if (!at) { out.fallback = true; return; }
        {
            const Span it2 = item_at(info, at);
            if (it2.b != at) { out.fallback = true; return; }
            r.nm_cut = (uint8_t)kNmLen[which];
            r.names = Span{at + kNmLen[which], it2.e};
        }
        // sample column
        Span f[8];
        int nf = 0;
        {
            const char *s = tok[9].b;
            for (const char *c = tok[9].b;; ++c) {
                if (c == tok[9].e || *c == ':') {
                    if (nf < 8) f[nf] = Span{s, c};
                    ++nf;
                    s = c + 1;
                    if (c == tok[9].e) break;
                }
            }
        }
        if (nf < 3 || nf > 8) { out.fallback = true; return; }
        r.n_sample_fields = (uint8_t)nf;
        r.gt_missing = f[0].size() == 3 && std::memcmp(f[0].b, "./.", 3) == 0;
        const Span last = f[nf - 1];
        const char *comma = static_cast<const char *>(std::memchr(last.b, ',', last.size()));
        r.ad_mode = nf <= 4 && comma != nullptr;
        auto count = [&](Span s, long long *o) {                  // '.' -> 0
            if (s.size() == 1 && *s.b == '.') { *o = 0; return true; }
            return plain_int(s, o);
        };
        long long ref = 0, alt = 0;
        if (r.ad_mode) {
            if (!count(Span{last.b, comma}, &ref) || !count(Span{comma + 1, last.e}, &alt)) { out.fallback = true; return; }
        } else {
            if (nf <= 4 && comma) { out.fallback = true; return; }
            if (!count(f[1], &ref) || !count(f[2], &alt)) { out.fallback = true; return; }
        }
        if (!fits_i32(ref)) { out.fallback = true; return; }
        r.refread = (int32_t)ref;
        out.recs.push_back(r);
    }
}

}  // namespace

struct duet_svs {
    int n_contigs = 0;
    const char *text = nullptr;
    std::vector<int64_t> sv_off;               // [n_contigs + 1]
    std::vector<Rec> recs;                     // contig-major, VCF order inside a contig
    std::vector<int32_t> group;                // rank of the CHROM string inside its contig
    bool any_group = false;
    std::vector<int64_t> csr_off;
    std::vector<uint64_t> lo, hi;
};

extern "C" {

int duet_decode_sv_vcf(const char *text, int64_t len, const char *contigs, int64_t contigs_len, int threads,
                       duet_svs **out, int64_t *n_svs, int64_t *n_joins) {
    *out = nullptr; *n_svs = 0; *n_joins = 0;
    Ctx cx;
    int n_contigs = 0;
    {   // contig list: '\n'-separated; contig c claims the CHROM strings 'chr'+c and c (read_file.py:30)
        const char *p = contigs, *e = contigs + contigs_len;
        while (p <= e) {
            const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(e - p)));
            const char *z = nl ? nl : e;
            const std::string c(p, (size_t)(z - p));
            for (const std::string &nm : {std::string("chr") + c, c}) {
                auto it = cx.owner.find(nm);
                if (it == cx.owner.end()) cx.owner.emplace(nm, n_contigs);
                else if (it->second != n_contigs) it->second = -1;           // claimed twice: not this reader's case
            }
            ++n_contigs;
            if (!nl) break;
            p = nl + 1;
        }
    }
    const bool dbg = std::getenv("DUET_DEBUG_TIMES") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    const int T = std::max(1, std::min(threads, 64));
    std::vector<Part> parts((size_t)T);
    {
        std::vector<const char *> cut((size_t)T + 1);
        cut[0] = text; cut[(size_t)T] = text + len;
        for (int t = 1; t < T; ++t) {
            const char *p = text + len * t / T;
            const char *nl = p < text + len ? static_cast<const char *>(std::memchr(p, '\n', (size_t)(text + len - p))) : nullptr;
            cut[(size_t)t] = nl ? nl + 1 : text + len;
        }
        for (int t = 1; t <= T; ++t) cut[(size_t)t] = std::max(cut[(size_t)t], cut[(size_t)t - 1]);
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back([&, t] { parse_range(cx, cut[(size_t)t], cut[(size_t)t + 1], parts[(size_t)t]); });
        parse_range(cx, cut[0], cut[1], parts[0]);
        for (auto &th : pool) th.join();
    }
    const double t1 = now();
    size_t total = 0;
    for (const Part &p : parts) { if (p.fallback) return DUET_DECODE_FALLBACK; total += p.recs.size(); }
    duet_svs *h = new (std::nothrow) duet_svs();
    if (!h) return DUET_DECODE_ERR_CAPACITY;
    h->n_contigs = n_contigs;
    h->text = text;
    // contig-major order, file order inside a contig (the parts are in file order): counting sort
    h->sv_off.assign((size_t)n_contigs + 1, 0);
    for (const Part &p : parts) for (const Rec &r : p.recs) ++h->sv_off[(size_t)r.contig + 1];
    for (int c = 0; c < n_contigs; ++c) h->sv_off[(size_t)c + 1] += h->sv_off[(size_t)c];
    h->recs.resize(total);
    {
        std::vector<int64_t> fill(h->sv_off.begin(), h->sv_off.end() - 1);
        for (const Part &p : parts) for (const Rec &r : p.recs) h->recs[(size_t)fill[(size_t)r.contig]++] = r;
    }
    // the contig's first record decides the dialect (:40-76); every other record has to agree with it
    h->group.assign(total, 0);
    for (int c = 0; c < n_contigs; ++c) {
        const size_t b = (size_t)h->sv_off[(size_t)c], e = (size_t)h->sv_off[(size_t)c + 1];
        if (b == e) continue;
        const Rec &first = h->recs[b];
        const bool ad = first.n_sample_fields > 4 ? false : first.ad_mode;
        Span alt_name{nullptr, nullptr};
        for (size_t i = b; i < e; ++i) {
            const Rec &r = h->recs[i];
            const bool r_ad = r.n_sample_fields > 4 ? false : r.ad_mode;
            if (r.sup_cut != first.sup_cut || r.nm_cut != first.nm_cut || r_ad != ad ||
                (first.n_sample_fields > 4) != (r.n_sample_fields > 4)) { delete h; return DUET_DECODE_FALLBACK; }
            if (r.chrom.size() != first.chrom.size() || std::memcmp(r.chrom.b, first.chrom.b, r.chrom.size()) != 0) alt_name = r.chrom;
        }
        if (alt_name.b) {      // 'c' and 'chr'+c rows in one contig: rank = order of the CHROM strings (the final sort key, :229)
            const std::string a(first.chrom.b, first.chrom.size()), z(alt_name.b, alt_name.size());
            const bool first_is_low = a < z;
            for (size_t i = b; i < e; ++i) {
                const Rec &r = h->recs[i];
                const bool same = r.chrom.size() == first.chrom.size() && std::memcmp(r.chrom.b, first.chrom.b, r.chrom.size()) == 0;
                h->group[i] = same == first_is_low ? 0 : 1;
            }
            h->any_group = true;
        }
    }
    const double t2 = now();
    // support-read names: split on ',' like str.split (an empty list string is one empty name) and hashed here
    h->csr_off.assign(total + 1, 0);
    {
        std::vector<int64_t> cnt(total);
        auto count_range = [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                const Span s = h->recs[i].names;
                int64_t k = 1;
                for (const char *c = s.b; c < s.e; ++c) k += *c == ',';
                cnt[i] = k;
            }
        };
        count_range(0, total);
        for (size_t i = 0; i < total; ++i) h->csr_off[i + 1] = h->csr_off[i] + cnt[i];
    }
    const size_t J = (size_t)h->csr_off[total];
    if (J >= (1ull << 30)) { delete h; return DUET_DECODE_FALLBACK; }
    h->lo.resize(J);
    h->hi.resize(J);
    {
        auto hash_range = [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                const Span s = h->recs[i].names;
                size_t k = (size_t)h->csr_off[i];
                const char *st = s.b;
                for (const char *c = s.b;; ++c) {
                    if (c == s.e || *c == ',') {
                        const int64_t off[2] = {0, (int64_t)(c - st)};
                        duet_hash_names(st, off, 1, &h->lo[k], &h->hi[k]);
                        ++k;
                        st = c + 1;
                        if (c == s.e) break;
                    }
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back([&, t] { hash_range(total * (size_t)t / (size_t)T, total * (size_t)(t + 1) / (size_t)T); });
        hash_range(0, total / (size_t)T);
        for (auto &th : pool) th.join();
    }
    if (dbg) std::fprintf(stderr, "vcf decode: parse %.4f merge %.4f names %.4f s\n", t1 - t0, t2 - t1, now() - t2);
    *n_svs = (int64_t)total;
    *n_joins = (int64_t)J;
    *out = h;
    return DUET_OK;
}

int duet_svs_take(duet_svs *h, int64_t *sv_off, int32_t *pos, int32_t *svlen, int32_t *svread, int32_t *refread,
                  uint8_t *flags, int32_t *group, int32_t *has_groups, int64_t *csr_off, uint64_t *csr_key,
                  uint32_t *csr_chk, int64_t *str_span) {
    if (!h) return DUET_ERR_INVALID;
    const size_t S = h->recs.size();
    std::memcpy(sv_off, h->sv_off.data(), h->sv_off.size() * sizeof(int64_t));
    for (size_t i = 0; i < S; ++i) {
        const Rec &r = h->recs[i];
        pos[i] = r.pos; svlen[i] = r.svlen; svread[i] = r.svread; refread[i] = r.refread;
        flags[i] = r.gt_missing ? DUET_SV_GT_MISSING : 0;
        if (group) group[i] = h->group[i];
        const Span sp[4] = {r.chrom, r.ref, r.alt, r.svtype};
        for (int k = 0; k < 4; ++k) {
            str_span[(i * 4 + (size_t)k) * 2] = (int64_t)(sp[k].b - h->text);
            str_span[(i * 4 + (size_t)k) * 2 + 1] = (int64_t)sp[k].size();
        }
    }
    *has_groups = h->any_group ? 1 : 0;
    std::memcpy(csr_off, h->csr_off.data(), h->csr_off.size() * sizeof(int64_t));
    const size_t J = h->lo.size();
    if (J) std::memcpy(csr_key, h->lo.data(), J * sizeof(uint64_t));
    for (size_t j = 0; j < J; ++j) csr_chk[j] = (uint32_t)h->hi[j];
    delete h;
    return DUET_OK;
}

void duet_svs_free(duet_svs *h) { delete h; }

}  // extern "C"
