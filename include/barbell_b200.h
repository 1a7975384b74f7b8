/*
 * barbell_b200.h -- C ABI of the B200-native `annotate` hot path (libbarbell_b200.so).
 *
 * This is the drop-in boundary a Rust host (rickbeeloo/barbell @ 9a2b814) binds with `extern "C"`; see
 * INTEGRATION.md for the binding.  No strings, no torch types, no C++ types cross it: plain pointers and sizes.
 * Every entry point cites the reference interface it replaces (paths relative to the reference root).
 *
 * Threading: a bb_ctx belongs to one host thread and one GPU (the reference keeps one Demuxer per worker thread,
 * src/annotate/annotator.rs:88-101).  Errors: every call returns 0 on success or a negative bb_status; the message is
 * available through bb_last_error().  There is NO CPU fallback: without a CUDA device bb_create fails.
 */
#ifndef BARBELL_B200_H
#define BARBELL_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 2

typedef enum {
    BB_OK = 0,
    BB_ERR_INVALID = -1,      /* bad argument / unsupported geometry (message says which) */
    BB_ERR_CUDA = -2,         /* CUDA runtime error (sticky on the ctx) */
    BB_ERR_OVERFLOW = -3,     /* caller's row buffer too small (n_rows tells how many were produced) */
    BB_ERR_NO_DEVICE = -4,    /* no CUDA device: the product has no CPU path */
    BB_ERR_KIT = -5,          /* unknown kit / malformed query set (the reference panics: kits.rs:704, barcodes.rs:113-143) */
    BB_ERR_IO = -6
} bb_status;

/* BarcodeType, src/annotate/barcodes.rs:8-14 */
enum { BB_FTAG = 0, BB_RTAG = 1, BB_FFLANK = 2, BB_RFLANK = 3 };
/* sassy::Strand as serialised in annotation.tsv, src/annotate/searcher.rs:100-142 */
enum { BB_FWD = 0, BB_RC = 1 };

/* One BarcodeGroup (src/annotate/barcodes.rs:57-71) in flat form. */
typedef struct {
    const char *flank;      /* prefix + 'N'*mask + suffix                        barcodes.rs:145-154 */
    int32_t flank_len;
    int32_t k_flank;        /* k_cutoff after set_flank_threshold                annotator.rs:216-229 */
    int32_t bar0, bar1;     /* bar_region, INCLUSIVE end                         barcodes.rs:192 */
    int32_t pad0, pad1;     /* pad_region, pad1 not clamped to the sequence      barcodes.rs:160-163 */
    int32_t match_type;     /* BB_FTAG / BB_RTAG */
    int32_t n_barcodes;
    int32_t bar_len;        /* length of every padded barcode pattern            barcodes.rs:165-173 */
    const char *barcodes;   /* n_barcodes * bar_len bytes, forward orientation */
} bb_group;

/* One annotation.tsv row (BarbellMatch, src/annotate/searcher.rs:31-64) without its strings:
   read_id is the caller's (read_idx), label is bb_groupset_label(group_idx, label_idx) or "flank" when -1. */
typedef struct {
    uint32_t read_idx;      /* index of the read inside the submitted batch */
    uint32_t read_len;
    int64_t  rel_dist_to_end;
    int64_t  read_start_bar, read_end_bar;
    int64_t  read_start_flank, read_end_flank;
    int64_t  bar_start, bar_end;
    int32_t  flank_cost, barcode_cost;
    int32_t  label_idx;
    int32_t  group_idx;
    uint8_t  match_type;
    uint8_t  strand;
    uint8_t  pad_[6];
} bb_row;

/* AnnotateConfig (src/config.rs:3-12) fields the search needs + device selection. */
typedef struct {
    int32_t device;             /* CUDA device ordinal */
    float   alpha;              /* --alpha, bin/main.rs:110-111 (default 0.4) */
    double  min_score;          /* --min-score, bin/main.rs:98-101 (default 0.2) */
    double  min_score_diff;     /* --min-score-diff, bin/main.rs:102-105 (default 0.1) */
    uint64_t max_batch_bytes;   /* capacity of the staging buffers for bb_annotate (0 = 256 MiB) */
    uint32_t max_batch_reads;   /* (0 = 4 Mi reads) */
    uint32_t flags;             /* bit 0: disable the lossless pre-filter (exact full-length scan everywhere);
                                   bit 1: nibble-pack the head of every batch on the host cores before the PCIe copy while the tail is copied
                                   as it is; the split adapts to the measured pack and link rates (bb_annotate / bb_submit);
                                   bit 2: like bit 1 with the denser wire format: 2 bits per base for A/C/G/T plus an exception list for every
                                   other byte (a batch with more than ~1.5 % such bytes switches the context to the nibble format) */
    uint32_t policy;            /* BB_POL_* bits: choices of sassy 0.2.1 (Cargo.lock:1060-1063; called at src/annotate/searcher.rs:282-288, 438)
                                   that the reference's own tests (src/annotate/cigar_parse.rs:104-176) do not pin.  0 = the documented defaults;
                                   a maintainer who can run upstream flips them here -- no kernel changes */
} bb_opts;

enum {
    BB_POL_S1_LEFT = 1,        /* `search` reports the LEFT end of a bottom-row cost plateau (default: the right end) */
    BB_POL_S2_PAT_FIRST = 2,   /* traceback prefers a pattern-only step [Del] over a text-only step [Ins] (default: text-only first) */
    BB_POL_S5_LAST = 4,        /* best match per barcode pattern = the LAST of equal lowest-cost minima (default: the first, searcher.rs:294-300) */
    BB_POL_S6_RC_FIRST = 8,    /* `search` lists reverse-complement matches before forward ones (default: forward first) */
    BB_POL_S3_ROUND = 16,      /* overhang cost of t rows = round-to-nearest(t*alpha) (default: floor) */
    BB_POL_S3_CEIL = 32        /* ... = ceil(t*alpha) */
};
/* the setting hosts should pass unless they know better; the one constant to change (with g_policy in oracle/barbell_oracle.c) once
   tools/ref_parity.sh --policy auto has been run against upstream */
#define BB_POL_DEFAULT 0u

typedef struct bb_ctx bb_ctx;
typedef struct bb_groupset bb_groupset;

/* ---- pattern-set construction (host side; replaces BarcodeGroup::new_from_kit / new_from_fasta / new,
 *      src/annotate/barcodes.rs:106-197, 251-315, and get_kit_info, src/kits/kits.rs:635-708) ---- */
int  bb_groups_from_kit(const char *kit, int use_extended, bb_groupset **out, char *err, size_t errlen);
/* one FASTA per group; types[i] = BB_FTAG / BB_RTAG (bin/main.rs:78-96) */
int  bb_groups_from_fasta(const char *const *paths, const int32_t *types, int32_t n, bb_groupset **out, char *err,
                          size_t errlen);
/* one group from in-memory sequences (BarcodeGroup::new); appends to *out if it is non-NULL */
int  bb_groups_add(bb_groupset **out, const char *const *seqs, const char *const *labels, int32_t n, int32_t type,
                   char *err, size_t errlen);
/* annotate_with_groups, annotator.rs:216-229: max_flank_errors < 0 selects get_edit_cut_off(prefix+suffix) */
int  bb_groups_set_flank_threshold(bb_groupset *gs, int32_t max_flank_errors);
int32_t bb_groups_count(const bb_groupset *gs);
const bb_group *bb_groups_data(const bb_groupset *gs);
const char *bb_groups_label(const bb_groupset *gs, int32_t group_idx, int32_t label_idx);
void bb_groups_free(bb_groupset *gs);
/* get_edit_cut_off, src/annotate/edit_model.rs:2-11 */
int32_t bb_edit_cut_off(int32_t effective_len);
/* get_barcodes, src/kits/kits.rs:741-816: comma-joined labels into `out`; returns their count or a negative status */
int  bb_label_range(const char *from_label, const char *to_label, int use_12a, char *out, size_t outlen);
/* lookup_barcode_seq, src/kits/kits.rs:1074-1103: NULL when unknown */
const char *bb_lookup_barcode_seq(const char *label);

/* get_kit_info, src/kits/kits.rs:635-708: preset name, its templates' label ranges ("BC01 - BC96; ...") and whether the kit
 * uses the double-label filter pattern sets (safe_patterns / maximize_patterns of its KitConfig, kits.rs:466-633) */
int  bb_kit_info(const char *kit, char *name, size_t namelen, char *ranges, size_t rangeslen, int *double_label, char *err,
                 size_t errlen);

/* ---- the operator (replaces Demuxer::{new,add_query_group,demux}, src/annotate/searcher.rs:202-227, 430-490) ---- */
int  bb_create(const bb_opts *opts, bb_ctx **out, char *err, size_t errlen);
void bb_destroy(bb_ctx *ctx);
const char *bb_last_error(const bb_ctx *ctx);
int  bb_set_groups(bb_ctx *ctx, const bb_group *groups, int32_t n_groups);

/* Batch form of DemuxProcessor::process_record (annotator.rs:122-135) with HOST buffers: `bases` holds the reads'
   raw sequence bytes back to back, offsets[n_reads+1] their byte offsets.  Rows come back grouped by read in input
   order, within a read in collapse_overlapping_matches order (interval.rs:4-28).  Reads without a hit emit no row. */
int  bb_annotate(bb_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, bb_row *rows,
                 uint64_t rows_cap, uint64_t *n_rows);
/* Same with DEVICE-resident inputs (d_bases 16-byte aligned, d_offsets = uint64[n_reads+1] on the device), enqueued on
   `stream` (a cudaStream_t, may be NULL); rows stay on the device until bb_fetch_rows. */
int  bb_annotate_device(bb_ctx *ctx, const void *d_bases, const void *d_offsets, uint32_t n_reads, uint64_t total_bytes,
                        void *stream, uint64_t *n_rows);
int  bb_fetch_rows(bb_ctx *ctx, bb_row *rows, uint64_t rows_cap, uint64_t *n_rows);

/* Pipelined form: up to BB_MAX_INFLIGHT batches may be submitted before the first collect; batches rotate over that many
   CUDA streams (each with its own worker thread and device buffers) so that the host-side packing, the host->device copy
   and the kernels of different batches overlap.  The caller owns `bases` and
   `offsets` until bb_collect has returned that batch_tag (they are read by the DMA engine in place: use pinned memory
   for full PCIe speed).  `rows` returned by bb_collect live in the result buffer of the engine
   that ran the batch (batch number i of a ctx runs on engine i % BB_MAX_INFLIGHT): they stay valid until BB_MAX_INFLIGHT further
   batches have been submitted on the ctx, so a host may hand them to a writer thread and go on submitting. */
#define BB_MAX_INFLIGHT 4
/* optional, before the first bb_submit: sizes the device buffers of every engine for batches of up to max_reads reads / max_bases
   bases and loads the kernels, by running an all-'A' batch of that shape through each engine (all engines at once).  Without it
   the first BB_MAX_INFLIGHT batches pay for the allocations (tens of ms each); results are the same either way. */
int  bb_reserve(bb_ctx *ctx, uint32_t max_reads, uint64_t max_bases);
int  bb_submit(bb_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, uint64_t batch_tag);
int  bb_collect(bb_ctx *ctx, uint64_t *batch_tag, const bb_row **rows, uint64_t *n_rows);
/* bb_submit for a host that packs while it parses (the FASTQ reader of `barbell annotate` does: src/io/io.rs:27-32 -> pinned ring):
   `crumbs` is the 2-bit wire format of the batch's n_bases bases as ONE gapless stream (built read by read with
   bb_pack_crumbs_append), exc[0, n_exc) its exception entries, offsets[n_reads + 1] the reads' base offsets (offsets[n_reads] ==
   n_bases).  A quarter of the bytes cross PCIe and no host core touches the bases again.  Ownership as for bb_submit. */
int  bb_submit_packed(bb_ctx *ctx, const uint8_t *crumbs, uint64_t n_bases, const uint64_t *exc, uint64_t n_exc,
                      const uint64_t *offsets, uint32_t n_reads, uint64_t batch_tag);

/* ProgressTracker counters (annotator.rs:109-113): out = {total reads, reads with >=1 row, reads with none} */
int  bb_counters(const bb_ctx *ctx, uint64_t out[3]);
/* Stage timings of the last bb_annotate_device call, milliseconds on the device: {scan, sort+resolve, trace, barcode, collapse} */
int  bb_last_stage_ms(bb_ctx *ctx, float out[5]);
/* number of kernels this library launched since bb_create */
uint64_t bb_kernel_launches(const bb_ctx *ctx);
/* bytes copied host -> device by bb_annotate / bb_submit so far (with flags bit 1 the head of every batch travels nibble-packed) */
uint64_t bb_h2d_bytes(const bb_ctx *ctx);
/* debugging / parity of the flank stage alone: the hit list after the flank search of the last bb_annotate_device call,
   6 int32 per hit {read_idx, group, strand, text_start, text_end, cost} in (read, group, Fwd-before-Rc, end) order */
int  bb_fetch_flank_hits(bb_ctx *ctx, int32_t *out6, uint64_t cap, uint64_t *n_hits);

/* page-locked host memory for the batch buffers handed to bb_submit / bb_annotate (NULL on failure) */
void *bb_host_alloc(size_t bytes);
void bb_host_free(void *p);

/* the wire format of flags bit 1, exposed for hosts that want to pack while parsing: dst[i] = set(src[2i]) | set(src[2i+1]) << 4,
   set() = 4-bit IUPAC base set (A=1, C=2, G=4, T=8; non-IUPAC bytes 0); dst holds (n+1)/2 bytes */
int  bb_pack_nibbles(const uint8_t *src, uint64_t n, uint8_t *dst);
/* the wire format of flags bit 2: dst[i] = crumb(src[4i]) | crumb(src[4i+1]) << 2 | crumb(src[4i+2]) << 4 | crumb(src[4i+3]) << 6 with
   A C G T (any case, U as T) = 0 1 2 3, dst holds (n+3)/4 bytes; every other byte has crumb 0 and an entry (position << 4 | set())
   in exc[0, *n_exc), appended in blocks whose unused entries are ~0.  BB_ERR_OVERFLOW when exc_cap entries do not suffice */
int  bb_pack_crumbs(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t *exc, uint64_t exc_cap, uint64_t *n_exc);
/* the same format built incrementally: appends n bases at base position *pos of the stream in dst (base i lives in dst[i >> 2], bits
   2 * (i & 3); dst bytes past the stream must be zero or unwritten), advances *pos, appends exception entries (stream position << 4 |
   set()) at exc[*n_exc ...) and advances *n_exc.  One writer per stream.  BB_ERR_OVERFLOW when exc_cap entries do not suffice. */
int  bb_pack_crumbs_append(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t *pos, uint64_t *exc, uint64_t exc_cap, uint64_t *n_exc);
/* the same for a text line whose length is not known yet (the sequence line of a FASTQ record, read in ONE pass): appends src[0, e),
   e = index of the first '\n' in src[0, n) (*found = 1) or n (*found = 0); a '\r' in front of the '\n' is not part of the line;
   *line_len = e without that '\r'.  dst needs 32 bytes of slack past the end of the stream (they are written as zero). */
int  bb_pack_crumbs_append_line(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t *pos, uint64_t *exc, uint64_t exc_cap, uint64_t *n_exc,
                                uint64_t *line_len, int *found);

/* ---- the stages that consume annotation.tsv (host side, no GPU): filter, inspect, trim -- so that `barbell kit` runs the
 *      reference's whole pipeline, src/kits/use_kit.rs:11-109 ---- */
/* the kit pattern sets of src/kits/kits.rs:175-236 (single/double label x safe/maximize) as pattern strings */
int  bb_kit_filter_patterns(int double_label, int maximize, const char *const **patterns, int32_t *n);
/* pattern_from_str!, src/filter/pattern.rs:242-383: parses one pattern string; `canonical` receives a normalised dump
 * of the parsed elements.  A string the reference's macro panics on returns BB_ERR_INVALID with the message in err. */
int  bb_pattern_parse(const char *pattern, char *canonical, size_t canonical_len, char *err, size_t errlen);
/* filter, src/filter/filter.rs:10-119 (+ check_filter_pass :183-214, match_pattern pattern.rs:205-240): reads
 * annotation.tsv, writes the rows of every read whose annotations are covered by its longest matching pattern (cuts
 * column filled in) to `output` and, when `dropped` is non-NULL, the other reads' rows there.
 * counts = {total, kept, dropped} reads (progress counters, filter.rs:55-97). */
int  bb_filter(const char *annotated, const char *output, const char *dropped, const char *const *patterns,
               int32_t n_patterns, uint64_t counts[3], char *err, size_t errlen);
/* inspect, src/inspect/inspect.rs:133-208: pattern string per read (get_group_structure :15-117) into
 * read_pattern_out (nullable), histogram of the top_n patterns on stdout */
int  bb_inspect(const char *annotated, int32_t top_n, const char *read_pattern_out, int32_t bucket_size, char *err,
                size_t errlen);
/* TrimConfig, src/config.rs:19-32 */
typedef struct {
    int32_t add_labels, add_orientation, add_flank, sort_labels;
    int32_t only_side;          /* 0 = none, 1 = LabelSide::Left, 2 = LabelSide::Right (trim.rs:25-29) */
    int32_t write_full_header, skip_trim, flip, gzip;
    const char *failed_out;     /* ids of reads with annotations but no slice; NULL = not written */
    int32_t threads;            /* plain FASTQ inputs are cut into chunks and trimmed by this many threads (0 = up to 16), output order unchanged */
} bb_trim_opts;
/* trim_matches, src/trim/trim.rs:317-480: cuts every read of the FASTQ files that has rows in filtered.tsv and writes
 * <out_dir>/<label>.trimmed.fastq[.gz].  counts = {total, trimmed, trimmed_split, failed} reads (trim.rs:19-22). */
int  bb_trim(const char *filtered, const char *const *fastq, int32_t n_fastq, const char *out_dir, const bb_trim_opts *opts,
             uint64_t counts[4], char *err, size_t errlen);

int  bb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
