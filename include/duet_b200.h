/*
 * duet_b200 -- C ABI of the B200-native sv_phasing hot path.
 *
 * The reference (yekaizhou/duet) has no FFI of its own; its seam for this path is the
 * Python call
 *     duet.sv_phasing_fn.generate_phased_callset(vcf_path, sam_home, svlen_thres,
 *                                                suppread_thres, thread, include_all_ctgs)
 * (src/duet/sv_phasing_fn.py:185), reached from src/duet/sv_phasing.py:17.  A maintainer
 * binds this library with ctypes (INTEGRATION.md shows the stub); the entry points below
 * are what that binding calls after the host has decoded the haplotagged alignments and
 * the SV VCF into columns.  Plain pointers and sizes only: no Python, torch or C++ types.
 *
 * Units follow the reference's domain:
 *   shard   one (sample, contig) pair -- every dict / set / loop of the reference is per
 *           contig (sv_phasing_fn.py:15-18,195-212), so shards never exchange data
 *   read    one haplotagged alignment row kept by read_hap_bam (sv_phasing_fn.py:28-29)
 *   SV      one VCF record kept by parse_vcf for a contig (read_file.py:30)
 *   join    one support-read name of one SV (an RNAMES / READS entry, read_file.py:48-55)
 *
 * All functions return DUET_OK (0) or a DUET_ERR_* code; duet_last_error() gives the text.
 * There is no CPU fallback: without a usable CUDA device duet_create() fails.
 * One host thread per handle; handles on different GPUs are independent.
 */
#ifndef DUET_B200_H
#define DUET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUET_ABI_VERSION 1

enum {
    DUET_OK = 0,
    DUET_ERR_INVALID = 1,        /* bad argument / inconsistent offsets                       */
    DUET_ERR_CUDA = 2,           /* a CUDA call failed (text in duet_last_error)              */
    DUET_ERR_NO_DEVICE = 3,      /* no CUDA device: the product path refuses to run           */
    DUET_ERR_HASH_COLLISION = 4, /* two different names share a 64-bit key (hi words differ)  */
    DUET_ERR_BAD_HP = 5,         /* HP not in {1,2} inside a multi-phase-set SV: the reference
                                    raises KeyError there (sv_phasing_fn.py:96)               */
    DUET_ERR_ZERO_DIVISION = 6,  /* svread + refread == 0 (sv_phasing_fn.py:123)              */
    DUET_ERR_STATE = 7           /* call order violated (execute before upload, ...)          */
};

enum {
    DUET_MEM_HOST = 0,           /* host columns (pageable or page-locked): duet_phase_upload copies them    */
    DUET_MEM_DEVICE = 1,         /* device columns, used in place                                            */
    DUET_MEM_HOST_MAPPED = 2     /* page-locked host columns (duet_host_alloc / cudaHostAlloc /
                                    cudaHostRegister): everything is copied EXCEPT read_tag, which the
                                    kernels read in place over the bus -- only the rows that joined are
                                    ever touched (~8 % of them at WGS 30x), so the largest column never
                                    crosses PCIe in full                                                     */
};

/* sv_flags bits */
enum { DUET_SV_GT_MISSING = 1 };   /* GT == "./." (sv_phasing_fn.py:190) */

/* value of `cls` for SVs removed by the svlen / support / GT filter (sv_phasing_fn.py:189-190) */
#define DUET_CLS_FILTERED 255

/* The literals of sv_phasing_fn.py:76,146-177, kept together so they live in __constant__
 * memory.  duet_default_thresholds() fills in the reference's values. */
typedef struct duet_thresholds {
    int32_t svlen_thres;        /* -s, default 50  (utils.py)                          */
    int32_t suppread_thres;     /* -r, default 2                                        */
    int32_t pc_max;             /* 8100   :76,88,201                                    */
    int32_t c0_sv_num_min;      /* 4      :146                                          */
    int32_t c2_sv_num_min;      /* 3      :151                                          */
    int32_t c2_hap0_min;        /* 6      :154                                          */
    int32_t c1_ref_num_max;     /* 10     :172                                          */
    int32_t _pad;
    double c2_sv_ratio_min;     /* 0.72   :149                                          */
    double c2_avgsc_diff_max;   /* 1369.5 :150                                          */
    double c1_one_ratio_lo;     /* 0.24   :160                                          */
    double c1_one_ratio_hi;     /* 0.9    :162                                          */
    double c1_hapread_ratio;    /* 0.75   :163,166                                      */
    double c1_avgsc_diff_max;   /* 2400   :163,166                                      */
    double c1_two_ratio_a;      /* 0.3    :169                                          */
    double c1_two_ratio_b;      /* 0.45   :171                                          */
    double c1_two_ratio_c;      /* 0.75   :176                                          */
    double c1_totsc_ratio_max;  /* 9.72   :177                                          */
} duet_thresholds;

/* Tags of one haplotagged read, one 16-byte record = half a DRAM sector: a joined read costs the
 * device exactly one random access.  `chk` is the low 32 bits of the HIGH word of the read-name hash
 * (the join key is the low word): a joined pair whose checks differ is two different names sharing a
 * 64-bit key and aborts the call (DUET_ERR_HASH_COLLISION). */
typedef struct duet_read_tag {
    int32_t ps;       /* PS:i  (sv_phasing_fn.py:29 int(s[-1][5:]))  */
    int32_t pc;       /* PC:i  (int(s[-2][5:]))                      */
    uint32_t chk;     /* name-hash check word                        */
    uint8_t hp;       /* HP:i  (int(s[-3][5:]))                      */
    uint8_t _pad[3];
} duet_read_tag;

/* Columnar input of one call: any number of shards, laid out back to back.
 *
 * Reads of shard s are rows [read_off[s], read_off[s+1]) IN FILE ORDER (a later row with the
 * same name overrides an earlier one, sv_phasing_fn.py:29).  SVs of shard s are
 * [sv_off[s], sv_off[s+1]) in VCF order; the support reads of SV i are CSR entries
 * [csr_off[i], csr_off[i+1]) in RNAMES order.
 *
 * read_off / sv_off are ALWAYS host pointers (tiny descriptors).  Every other array lives
 * where `mem` says: DUET_MEM_HOST (pageable or pinned; duet_phase_upload copies it),
 * DUET_MEM_DEVICE (used in place, must stay valid and unchanged until the next upload) or
 * DUET_MEM_HOST_MAPPED (as DUET_MEM_HOST, but `read_tag` must be page-locked and is read in place
 * by duet_phase_execute: keep it valid and unchanged until the results have been downloaded).
 * `csr_chk` may be NULL: then 64-bit key equality is trusted (no collision check).
 * `sv_group` may be NULL (= all zero).
 */
typedef struct duet_phase_input {
    int32_t mem;                 /* DUET_MEM_HOST | DUET_MEM_DEVICE | DUET_MEM_HOST_MAPPED       */
    int32_t n_shards;
    int64_t n_reads;             /* R < 2^31                                                     */
    int64_t n_svs;               /* S < 2^31                                                     */
    int64_t n_joins;             /* J < 2^30                                                     */
    const int64_t *read_off;     /* [n_shards+1] host                                            */
    const int64_t *sv_off;       /* [n_shards+1] host                                            */
    const uint64_t *read_key;    /* [R] low 64 bits of the name hash, never 0xFFFF...F; 16-byte aligned */
    const duet_read_tag *read_tag; /* [R] HP / PS / PC + check word                              */
    const int32_t *sv_pos;       /* [S]                                                          */
    const int32_t *sv_svlen;     /* [S] |SVLEN| (sv_phasing_fn.py:62)                            */
    const int32_t *sv_svread;    /* [S] SUPPORT / RE / SR value (read_file.py:40-47)             */
    const int32_t *sv_refread;   /* [S] column [15] of parse_vcf (read_file.py:56-76)            */
    const uint8_t *sv_flags;     /* [S] DUET_SV_* bits                                           */
    const int32_t *sv_group;     /* [S] or NULL: rank of the SV's CHROM string inside its shard  */
    const int64_t *csr_off;      /* [S+1]                                                        */
    const uint64_t *csr_key;     /* [J] 16-byte aligned (streamed with bulk copies, like read_key) */
    const uint32_t *csr_chk;     /* [J] check word of each support-read name, or NULL            */
} duet_phase_input;

#define DUET_N_FEATURES 6        /* hapread_ratio, sv_ratio, hap1_avgsc, hap2_avgsc, totsc_ratio,
                                    hap_avgsc_diff (sv_phasing_fn.py:112-132)                    */
#define DUET_N_COUNTERS 8        /* per shard: n_sv, n_kept, n_emitted, n_1|0, n_0|1, n_1|1,
                                    n_joins, n_hits                                              */

/* Caller-allocated HOST arrays filled by duet_phase_download.  Any pointer may be NULL
 * (that output is skipped).  Values of SVs with cls == DUET_CLS_FILTERED are unspecified
 * except gt == 0. */
typedef struct duet_phase_output {
    uint8_t *gt;           /* [S] 0 dropped, 1 "1|0", 2 "0|1", 3 "1|1" (sv_phasing_fn.py:213-222)    */
    int32_t *ps;           /* [S] phase set written for the SV (f['ps'])                             */
    uint8_t *cls;          /* [S] number of distinct PS among joined reads, capped at 2 (:192-194)   */
    int32_t *hap1, *hap2, *hap0, *allhap;   /* [S] each                                              */
    int64_t *totsc1, *totsc2;               /* [S] each                                              */
    double *features;      /* [DUET_N_FEATURES][S] planes                                            */
    int32_t *join_row;     /* [J] row (into the read columns) each support read joined to, -1 = miss */
    int32_t *order;        /* [S] first n_emitted: SV indices, shard by shard, each shard sorted by
                              (group, pos, class, VCF order) = the reference's stable order (:206-229) */
    int64_t *shard_counts; /* [n_shards][DUET_N_COUNTERS]                                            */
    int64_t n_emitted;     /* out                                                                    */
} duet_phase_output;

typedef struct duet_timings {
    float h2d_ms;          /* upload: host -> device copies                       */
    float device_ms;       /* all kernels of one execute                          */
    float d2h_ms;          /* download                                            */
    float kernel_ms[8];    /* init (k_init), build (k_table), probe, reduce, tail (k_tail: the three per-contig
                              steps in one cluster launch), oneps, predict, order (the same steps as separate
                              kernels, only when a contig has more SVs than a cluster holds) */
} duet_timings;

typedef struct duet_handle duet_handle;

int duet_abi_version(void);
void duet_default_thresholds(duet_thresholds *t);

int duet_create(int device_id, duet_handle **out);
void duet_destroy(duet_handle *h);
const char *duet_last_error(const duet_handle *h);   /* h may be NULL: error of the last failed duet_create */

int duet_set_thresholds(duet_handle *h, const duet_thresholds *t);

/* Use this stream (a cudaStream_t / CUstream, e.g. torch's current stream) for all work of the
 * handle; NULL = the handle's own non-blocking stream (default). */
int duet_set_stream(duet_handle *h, void *cuda_stream);

/* Arena layouts.  The library keeps the small input columns (sv_pos, sv_svlen, sv_svread, sv_refread, sv_flags,
 * sv_group, csr_off, csr_key, csr_chk -- in this order) in one device buffer and the results (gt, ps, cls, hap1,
 * hap2, hap0, allhap, totsc1, totsc2, features, shard_counts, join_row -- in this order) in another.  A caller
 * whose host arrays sit at the same offsets inside ONE page-locked allocation gets one copy each way instead of
 * one per column: `offsets` receives the 9 / 12 byte offsets, the return value is the arena size.  (Trailing
 * results may be left NULL -- e.g. no join_row; `order` is separate in any case.)  Arrays placed any other way
 * work too: they are copied one by one. */
int64_t duet_phase_input_layout(int64_t n_svs, int64_t n_joins, int64_t *offsets);
int64_t duet_phase_output_layout(int64_t n_svs, int64_t n_joins, int32_t n_shards, int64_t *offsets);

/* Stage the columns on the device (copy for HOST, alias for DEVICE) and size the join table. */
int duet_phase_upload(duet_handle *h, const duet_phase_input *in);
/* Launch the whole path on the staged columns (asynchronous; may be called repeatedly).
 * `per_kernel` != 0 records an event after every kernel (for duet_get_timings). */
int duet_phase_execute(duet_handle *h, int per_kernel);
/* Wait for the device, turn device-side error flags into a status, copy results out. */
int duet_phase_download(duet_handle *h, duet_phase_output *out);
/* upload + execute + download. */
int duet_phase_run(duet_handle *h, const duet_phase_input *in, duet_phase_output *out);

/* Page-locked host memory for the columns (so uploads run at full PCIe speed and decoders can
 * write straight into it).  Usable without a handle; fails with DUET_ERR_NO_DEVICE on a box
 * without a GPU. */
int duet_host_alloc(void **ptr, int64_t bytes);
int duet_host_free(void *ptr);
int duet_host_is_pinned(const void *ptr);   /* 1 if ptr lies in page-locked host memory known to CUDA, else 0 */

int duet_sync(duet_handle *h);

/* Developer instrumentation: with `enable` != 0 the kernels launched after the next upload stamp
 * (globaltimer ns, SM clock) per block at up to 8 marks; `out` (may be NULL) receives the stamps of
 * the last execute as int64[4 kernels: k_table, k_probe, k_reduce, k_tail][2048 blocks][8 marks][2].
 * Off by default: costs nothing. */
int duet_debug_timers(duet_handle *h, int enable, int64_t *out);

/* ---- kernel set B: span-position-distance clustering of SV signatures --------------------------
 * Stands where the reference shells out to `svim alignment ... --cluster_max_distance c`
 * (src/duet/sv_calling.py:14-15; default 0.9, src/duet/utils.py:27-28).  The clustering arithmetic
 * is NOT in the reference (svim 1.4.2 is an external tool): the spec implemented is written down in
 * csrc/cluster_kernels.cuh and DESIGN.md; parity with svim itself is unpinned. */
typedef struct duet_cluster_params {
    double max_distance;          /* 0.9                                                       */
    double position_normalizer;   /* 900                                                       */
    int32_t partition_window;     /* 1000: centres further apart are never compared            */
    int32_t _pad;
} duet_cluster_params;

typedef struct duet_cluster_input {
    int32_t mem;                  /* DUET_MEM_HOST | DUET_MEM_DEVICE (cluster_id lives there too) */
    int32_t _pad;
    int64_t n;                    /* signatures, < 2^31                                         */
    const int32_t *contig;        /* [n] contig id, < 65536                                     */
    const int32_t *type;          /* [n] signature type id, < 256 (DEL / INS / INV / DUP_TAN ...) */
    const int32_t *start;         /* [n] 0 <= start                                             */
    const int32_t *end;           /* [n] start <= end (insertions: start + length); start+end < 2^32 */
} duet_cluster_input;

void duet_default_cluster_params(duet_cluster_params *p);
/* cluster_id[i] = smallest signature index in i's cluster; *n_clusters = number of clusters;
 * *device_ms (optional) = device time of the call. */
int duet_cluster_run(duet_handle *h, const duet_cluster_input *in, const duet_cluster_params *params,
                     int32_t *cluster_id, int64_t *n_clusters, float *device_ms);
/* Event-timed stages of the last duet_cluster_run -- bucketed path: split (maxima, histogram, scatter); buckets;
 * boundary fix-up.  General path: keys; sort passes; runs + windowed distances; labels.  Fills names[k] / ms[k]
 * (cap >= 6), returns how many (0 if nothing was run). */
int duet_cluster_timings(duet_handle *h, const char **names, float *ms, int cap);

/* ---- host-side decoders (no GPU needed) -------------------------------------------------- */

enum {
    DUET_DECODE_ERR_INDEX = 20,    /* a line has fewer fields than the reference indexes -> IndexError  */
    DUET_DECODE_ERR_VALUE = 21,    /* int() of a tag value fails -> ValueError                          */
    DUET_DECODE_ERR_ASCII = 22,    /* byte >= 0x80 -> UnicodeDecodeError (.decode('ascii'), :25)        */
    DUET_DECODE_ERR_RANGE = 23,    /* HP outside 0..255 or PS/PC outside int32 (BAM aux ints are 32 bit) */
    DUET_DECODE_ERR_CAPACITY = 24, /* output arrays too small / out of memory                           */
    DUET_DECODE_ERR_FORMAT = 25,   /* not a BGZF-compressed BAM / truncated record                       */
    DUET_DECODE_FALLBACK = 26      /* duet_decode_sv_vcf: input outside what the native reader reproduces
                                      exactly -- use the general reader (duet_b200/read_file.py)          */
};

/* 128-bit name hash (duet_b200/namehash.py): name i is buf[off[i] .. off[i+1]).  key = lo,
 * check word = (uint32_t)hi. */
void duet_hash_names(const char *buf, const int64_t *off, int64_t n, uint64_t *lo, uint64_t *hi);

/* The same hash for whole RNAMES / READS lists without materialising the names: `blob` holds the
 * comma-separated lists of `n_lists` records joined by '\n' (read_file.py:48-55 splits them on ',').
 * Writes lens[n_lists] (names per record, an empty list string counts as one empty name, as
 * str.split does) and lo/hi for every name in order; `cap` = room in lo/hi.  Returns the number
 * of names, or -1 if the blob does not hold exactly n_lists records or cap is too small. */
int64_t duet_hash_name_lists(const char *blob, int64_t len, int64_t n_lists, int64_t cap, int64_t *lens,
                             uint64_t *lo, uint64_t *hi);

/* Separate HP / PS / PC (+ hash high words, may be NULL -> chk 0) columns -> tag records. */
void duet_pack_tags(int64_t n, const uint8_t *hp, const int32_t *ps, const int32_t *pc, const uint64_t *hi,
                    duet_read_tag *out);

/* `samtools view` text of one haplotagged BAM -> columns of the kept rows, in file order
 * (sv_phasing_fn.py:25-29).  Output arrays hold `cap` rows (duet_count_lines(text) is enough).
 * On error returns a DUET_DECODE_ERR_* code and the 0-based line number in *err_line. */
int duet_decode_sam_text(const char *text, int64_t len, int64_t cap, uint64_t *key, duet_read_tag *tag,
                         int64_t *n_rows, int64_t *n_lines, int64_t *err_line);
int64_t duet_count_lines(const char *text, int64_t len);

/* Threads ONE duet_decode_bam call may use to inflate its BGZF blocks (each block carries its own
 * sizes, so they inflate independently); default 1.  Process-wide; returns the previous value.  The
 * reference hands its `thread` count to `samtools view -@` (sv_phasing_fn.py:25); the host layer
 * here decodes contigs in parallel first and gives what is left to the blocks of each file. */
int duet_set_decode_threads(int n);

/* A whole haplotagged BAM file (BGZF bytes) -> the same columns `samtools view` + duet_decode_sam_text
 * would give: every record's last three text tokens are reconstructed and the reference's rule
 * (sv_phasing_fn.py:28-29) is applied to them.  *key_out / *tag_out are malloc'ed; release them with
 * duet_free.  On error *err_record is the 0-based record number. */
int duet_decode_bam(const unsigned char *data, int64_t len, uint64_t **key_out, duet_read_tag **tag_out,
                    int64_t *n_rows, int64_t *n_records, int64_t *err_record);
void duet_free(void *p);

/* The same decoders as a two-step job, so that the rows can land directly in columns the caller sizes from
 * the counts (page-locked, one slice per contig): duet_decode_reads scans one file's bytes -- kind
 * DUET_READS_SAM_TEXT (`samtools view` text) or DUET_READS_BAM (BGZF; inflated and walked in bounded
 * batches, so memory does not grow with the file) -- and holds the kept rows; duet_rows_take copies them to
 * `key_dst` / `tag_dst` (room for *n_rows each) and releases the job; duet_rows_free drops it unread. */
enum { DUET_READS_SAM_TEXT = 0, DUET_READS_BAM = 1 };
typedef struct duet_rows duet_rows;
int duet_decode_reads(const unsigned char *data, int64_t len, int kind, duet_rows **out, int64_t *n_rows,
                      int64_t *n_records, int64_t *err_at);
int duet_rows_take(duet_rows *rows, uint64_t *key_dst, duet_read_tag *tag_dst);
void duet_rows_free(duet_rows *rows);

/* The SV VCF (cuteSV / Sniffles2 / SVIM dialects) in one multi-threaded pass: what read_file.py::parse_vcf
 * (reference: src/duet/read_file.py:25-76) returns, as columns, with the support-read names hashed in the
 * same pass.  `contigs` = the contig list joined by '\n' (read_file.py:6-16); record i of contig c is row
 * sv_off[c] + i.  Regular input only: anything the native reader does not reproduce exactly (irregular INFO
 * items, integers that are not plain decimals, mixed dialects inside a contig, a CHROM string claimed by two
 * contigs, blank lines, non-ASCII bytes ...) returns DUET_DECODE_FALLBACK -- the general reader then decides,
 * raising what the reference raises.  duet_svs_take copies the columns out and releases the job:
 *   sv_off[n_contigs+1]; pos, svlen (signed, as parsed), svread, refread [S]; flags [S] (DUET_SV_*);
 *   group [S] (rank of the CHROM string inside its contig; *has_groups = 0 when all zero);
 *   csr_off[S+1], csr_key[J], csr_chk[J]; str_span[S][4][2] = (offset, length) into `text` of CHROM, REF, ALT
 *   and the SVTYPE value.  `text` must stay valid until duet_svs_take / duet_svs_free. */
typedef struct duet_svs duet_svs;
int duet_decode_sv_vcf(const char *text, int64_t len, const char *contigs, int64_t contigs_len, int threads,
                       duet_svs **out, int64_t *n_svs, int64_t *n_joins);
int duet_svs_take(duet_svs *h, int64_t *sv_off, int32_t *pos, int32_t *svlen, int32_t *svread, int32_t *refread,
                  uint8_t *flags, int32_t *group, int32_t *has_groups, int64_t *csr_off, uint64_t *csr_key,
                  uint32_t *csr_chk, int64_t *str_span);
void duet_svs_free(duet_svs *h);
int duet_get_timings(duet_handle *h, duet_timings *t);
/* Number of kernels this library launched on the handle since creation (bench "gpu_launches"). */
int64_t duet_launch_count(const duet_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* DUET_B200_H */
