/*
 * barbell_oracle.h -- CPU ORACLE for the `annotate` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (barbell_b200/, include/barbell_b200.h) never links, imports or calls it.
 *
 * It restates, in plain C, the per-read algorithm of rickbeeloo/barbell @ 9a2b814:
 *   Demuxer::demux                      src/annotate/searcher.rs:430-490
 *   collect_candidates_for_region       src/annotate/searcher.rs:267-337
 *   score_and_push_result               src/annotate/searcher.rs:339-426
 *   push_flank_only_result              src/annotate/searcher.rs:241-265
 *   rel_dist_to_end                     src/annotate/searcher.rs:183-199
 *   get_matching_region / map_pat_to_text_with_cost / compute_subpath_cost
 *                                       src/annotate/cigar_parse.rs:6-82
 *   collapse_overlapping_matches        src/annotate/interval.rs:4-79
 *   get_edit_cut_off                    src/annotate/edit_model.rs:2-11
 * and the published algorithms of the un-vendored crates the reference calls (absent from /root/reference):
 *   sassy 0.2.1          (Cargo.lock:1060-1063)  approximate search, local-minimum reporting, traceback, to_path
 *   cigar-lodhi-rs 0.1.0 (Cargo.lock:242-245)    subsequence score S_3(C, 1/2), paper Appendix B
 *   pa-types 1.2.0       (Cargo.lock:764-767)    Cigar / Pos conventions
 *
 * PARITY STATUS: the Sassy boundary is pinned only by the reference's five known-answer tests
 * (src/annotate/cigar_parse.rs:104-176) and the Lodhi score by the paper's two worked examples plus the three
 * perfect-score constants; everything those do not discriminate (reporting rule on plateaus, traceback
 * tie-break, overhang rounding, Rc path orientation) is a NAMED POLICY in barbell_oracle.c.
 * "parity unpinned" against upstream for those -- neither cargo nor the crates exist in this environment.
 */
#ifndef BARBELL_ORACLE_H
#define BARBELL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* alignment ops (sassy/pa-types names in brackets) */
enum { ORC_OP_MATCH = 0, ORC_OP_SUB = 1, ORC_OP_TEXT = 2 /* text char only [Ins] */, ORC_OP_PAT = 3 /* pattern char only [Del] */ };
enum { ORC_FWD = 0, ORC_RC = 1 };
enum { ORC_FTAG = 0, ORC_RTAG = 1, ORC_FFLANK = 2, ORC_RFLANK = 3 };

typedef struct {
    int32_t text_start, text_end;       /* half-open, forward text coordinates */
    int32_t pattern_start, pattern_end; /* half-open, pattern coordinates (differ from 0..m only with overhang) */
    int32_t cost;
    int32_t strand;
    int32_t n_ops;
    uint8_t *ops;                       /* n_ops ops in pattern order (Fwd: ascending text; Rc: descending text) */
} orc_match;

typedef struct {
    const char *flank;      /* prefix + 'N'*mask + suffix                        barcodes.rs:145-154 */
    int32_t flank_len;
    int32_t k_flank;        /* k_cutoff                                          annotator.rs:216-229 */
    int32_t bar0, bar1;     /* bar_region, INCLUSIVE end                         barcodes.rs:192 */
    int32_t pad0, pad1;     /* pad_region, pad1 not clamped                      barcodes.rs:160-163 */
    int32_t match_type;     /* ORC_FTAG / ORC_RTAG */
    int32_t n_barcodes;
    int32_t bar_len;        /* length of every padded barcode pattern            barcodes.rs:165-173 */
    const char *barcodes;   /* n_barcodes * bar_len bytes, forward orientation */
} orc_group;

typedef struct {
    float alpha;            /* overhang cost per pattern char, <0 = none         main.rs:110-111 */
    double min_score;       /* main.rs:98-101 */
    double min_score_diff;  /* main.rs:102-105 */
} orc_params;

/* One annotation.tsv row without the strings (searcher.rs:31-64); same field meaning as bb_row. */
typedef struct {
    uint32_t read_idx;
    uint32_t read_len;
    int64_t  rel_dist_to_end;
    int64_t  read_start_bar, read_end_bar;
    int64_t  read_start_flank, read_end_flank;
    int64_t  bar_start, bar_end;
    int32_t  flank_cost, barcode_cost;
    int32_t  label_idx;     /* index into the group's barcodes, -1 = "flank" */
    int32_t  group_idx;
    uint8_t  match_type;    /* ORC_FTAG.. */
    uint8_t  strand;        /* ORC_FWD / ORC_RC */
    uint8_t  pad_[6];
} orc_row;

/* --- sassy restatement ------------------------------------------------------------------ */
/* Searcher::<Iupac>::search(pattern, text, k): rc!=0 also searches the reverse complement; alpha<0 = no overhang.
   Returns number of matches, *out is malloc'ed (free with orc_free_matches). */
int  orc_search(const uint8_t *pattern, int m, const uint8_t *text, int n, int k, float alpha, int rc,
                orc_match **out);
void orc_free_matches(orc_match *ms, int n);
/* Match::to_path(): writes n_ops (i,j) pairs into ij (2*n_ops int32). */
void orc_to_path(const orc_match *mt, int32_t *ij);
/* cigar_parse.rs:71-82; returns 0 if None */
int  orc_get_matching_region(const orc_match *mt, int start, int end, int64_t *rs, int64_t *re);
/* cigar_parse.rs:6-45; returns 0 if None; out = {pi, ei+1, pj, ej+1, cost} */
int  orc_map_pat_to_text_with_cost(const orc_match *mt, int p_start, int p_end, int64_t out[5]);
/* Lodhi::new(3, 0.5).compute(cigar) over an op array */
double orc_lodhi(const uint8_t *ops, int n_ops);
/* bottom-row costs (debug/test): c has n+1 (+m if alpha>=0) entries; forward strand of the given text */
int  orc_bottom_row(const uint8_t *pattern, int m, const uint8_t *text, int n, float alpha, int32_t *c);

/* --- barbell restatement ---------------------------------------------------------------- */
/* edit_model.rs:2-11 */
int  orc_edit_cut_off(int l);
/* interval.rs:4-79 on rows of one read; returns new count, rows rewritten in place */
int  orc_collapse(orc_row *rows, int n, float threshold);
/* searcher.rs:430-490 for one read; rows appended to out (cap rows); returns count or -1 on overflow */
int  orc_demux(const orc_group *groups, int n_groups, const orc_params *prm, uint32_t read_idx,
               const uint8_t *read, int n, orc_row *out, int cap);
/* batch driver (annotator.rs:122-135 per record), n_threads OpenMP threads, rows in input order.
   offsets has n_reads+1 entries. Returns row count or -1 on overflow. */
int64_t orc_demux_batch(const orc_group *groups, int n_groups, const orc_params *prm, const uint8_t *bases,
                        const uint64_t *offsets, uint32_t n_reads, int n_threads, orc_row *out, int64_t cap);
/* flank stage only (used to check the GPU flank stage): out6 = {read_idx, group, strand, text_start, text_end, cost} */
int64_t orc_flank_hits_batch(const orc_group *groups, int n_groups, const orc_params *prm, const uint8_t *bases,
                             const uint64_t *offsets, uint32_t n_reads, int n_threads, int32_t *out6, int64_t cap);
int  orc_max_threads(void);

/* policies (S1..S7 of SURVEY.md A.3). Defaults documented in barbell_oracle.c */
enum {
    ORC_POL_S1_LEFT = 1,       /* S1: report the LEFT end of a bottom-row cost plateau (default: the right end) */
    ORC_POL_S2_PAT_FIRST = 2,  /* S2: traceback prefers pattern-only [Del] over text-only [Ins] (default: text-only first) */
    ORC_POL_S5_LAST = 4,       /* S5: best-per-pattern keeps the LAST of equal lowest-cost minima (default: the first) */
    ORC_POL_S6_RC_FIRST = 8,   /* S6: `search` lists the Rc matches before the forward ones (default: forward first) */
    ORC_POL_S3_ROUND = 16,     /* S3: overhang cost of t rows = round-to-nearest(t*alpha) (default: floor) */
    ORC_POL_S3_CEIL = 32       /* S3: ... = ceil(t*alpha) */
};
typedef struct {
    int use_myers;         /* 1 = bit-vector scan + windowed traceback (fast); 0 = naive full DP matrix (cross-check) */
    int flags;             /* ORC_POL_* bits; 0 = the documented defaults.  Same bits as bb_opts.policy of the product. */
} orc_policy;
void orc_set_policy(const orc_policy *p);

#ifdef __cplusplus
}
#endif
#endif
