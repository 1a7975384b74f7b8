/*
 * barbell_oracle.c -- CPU ORACLE (test infrastructure, see barbell_oracle.h). Plain C, scalar, one Demuxer state
 * per worker thread (pthreads) exactly like the reference's one-Demuxer-per-worker (src/annotate/annotator.rs:88-101).
 *
 * NAMED POLICIES ("parity unpinned" against sassy 0.2.1 for all of them):
 *  S1 reporting   : `search` reports LOCAL MINIMA of the bottom-row cost: walking end positions left to right,
 *                   position p-1 is reported when cost[p] > cost[p-1], the last strict change before p-1 was a
 *                   decrease (initially true) and cost[p-1] <= k; the final position is reported if the walk is
 *                   still "decreasing" and its cost <= k.  (= rightmost position of a plateau.)
 *  S2 traceback   : from the end cell, prefer  match > substitution > text-only step [Ins] > pattern-only step [Del]
 *                   (the 5 reference KATs require diagonal before pattern-only; the rest is free).
 *  S3 overhang    : with alpha >= 0 the first DP column is floor((float)i * alpha) (f32 arithmetic) and, past the
 *                   text end, `m` virtual end positions n+t (t=1..m) carry cost D[m-t][n] + floor((float)t*alpha);
 *                   they take part in the S1 walk; a match reported there has pattern_end = m-t, text_end = n.
 *                   A traceback that reaches text column 0 at pattern row i>0 stops there (pattern_start = i).
 *  S4 Rc matches  : the Rc strand is the forward search of the SAME pattern in reverse_complement(text)
 *                   (sassy searches complement(pattern) in reversed text -- same DP).  text_start/end are mapped
 *                   back to forward coordinates; ops stay in pattern order; to_path() walks the text downwards
 *                   from text_end-1 for strand==Rc.  The reference overwrites `strand` of forward-computed barcode
 *                   matches with the flank's strand BEFORE calling to_path (searcher.rs:333 vs :385): reproduced.
 *  S5 encoded     : search_encoded_patterns = per pattern, forward only, no overhang, S1 minima in ascending order.
 *  S6 order       : `search` returns all forward matches (ascending end) and then all Rc matches (ascending end in the
 *                   reversed frame).
 *  S7 alphabet    : IUPAC letters, case-insensitive, U=T, X and every non-letter byte match nothing.
 *  Lodhi          : S_3(C, 1/2) by the forward recurrence in orc_lodhi() in IEEE f64, positions advance on every op.
 * Every one of S1, S2, S3, S5, S6 can be FLIPPED at run time (orc_policy.flags, ORC_POL_* in barbell_oracle.h); the product
 * has the same switches (bb_opts.policy), and the GPU == oracle suite runs under every setting, so a maintainer with the
 * upstream crates can move both to whatever sassy 0.2.1 really does without touching a kernel (tools/ref_parity.sh --policy).
 */
#include "barbell_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NW_MAX 4          /* patterns up to 256 characters */
#define PADDING 10        /* src/lib.rs:10 */

static orc_policy g_policy = {1, 0};
void orc_set_policy(const orc_policy *p) { g_policy = *p; }
int orc_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

/* ---------------- alphabet (S7) ---------------- */
static uint8_t CODE[256];
static uint8_t RCCHAR[256];   /* barcodes.rs:398-441 */
static int g_tables_ready = 0;
static void init_tables(void) {
    if (g_tables_ready) return;
    memset(CODE, 0, sizeof CODE);
    const char *L = "ACGTURYSWKMBDHVN";
    const uint8_t V[] = {1, 2, 4, 8, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15};
    for (int i = 0; L[i]; i++) { CODE[(uint8_t)L[i]] = V[i]; CODE[(uint8_t)(L[i] | 0x20)] = V[i]; }
    for (int i = 0; i < 256; i++) RCCHAR[i] = (uint8_t)i;
    const char *a = "ACTGRYSWKMBDHVNX", *b = "TGACYRSWMKVHDBNX";
    for (int i = 0; a[i]; i++) { RCCHAR[(uint8_t)a[i]] = (uint8_t)b[i]; RCCHAR[(uint8_t)(a[i] | 0x20)] = (uint8_t)(b[i] | 0x20); }
    __sync_synchronize();
    g_tables_ready = 1;
}
static inline uint8_t comp_code(uint8_t c) { /* A<->T, C<->G : reverse the 4 bits */
    return (uint8_t)(((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3));
}

/* ---------------- bit-vector column machine ---------------- */
typedef struct {
    int m, nw, last_bit;
    uint64_t eq[16][NW_MAX];
    uint64_t pv_plain[NW_MAX];   /* column D[i] = i */
    uint64_t pv_over[NW_MAX];    /* column D[i] = floor(i*alpha)  (S3) */
    int ov[64 * NW_MAX + 1];     /* floor(t*alpha) */
    int has_over;
} pat_t;

static void pat_build(pat_t *P, const uint8_t *pc, int m, float alpha) {
    memset(P, 0, sizeof *P);
    P->m = m; P->nw = (m + 63) / 64; if (P->nw < 1) P->nw = 1; P->last_bit = (m - 1) & 63;
    for (int i = 0; i < m; i++)
        for (int c = 0; c < 16; c++)
            if (pc[i] & c) P->eq[c][i >> 6] |= 1ull << (i & 63);
    for (int i = 0; i < m; i++) P->pv_plain[i >> 6] |= 1ull << (i & 63);
    P->has_over = alpha >= 0.0f;
    for (int t = 0; t <= m; t++) {            /* S3: floor (default) / round-to-nearest / ceil of the f32 product */
        float v = (float)t * alpha;
        P->ov[t] = !P->has_over ? t : (g_policy.flags & ORC_POL_S3_ROUND) ? (int)floorf(v + 0.5f) : (g_policy.flags & ORC_POL_S3_CEIL) ? (int)ceilf(v) : (int)floorf(v);
    }
    for (int i = 0; i < m; i++)
        if (P->ov[i + 1] - P->ov[i]) P->pv_over[i >> 6] |= 1ull << (i & 63);
}

static inline int col_step(const pat_t *P, uint64_t *pv, uint64_t *mv, int code) {
    int hin = 0;
    for (int b = 0; b < P->nw; b++) {
        uint64_t eq = P->eq[code][b], Pv = pv[b], Mv = mv[b];
        uint64_t xv = eq | Mv;
        if (hin < 0) eq |= 1;
        uint64_t xh = (((eq & Pv) + Pv) ^ Pv) | eq;
        uint64_t ph = Mv | ~(xh | Pv);
        uint64_t mh = Pv & xh;
        int bit = (b == P->nw - 1) ? P->last_bit : 63;
        int hout = (int)((ph >> bit) & 1) - (int)((mh >> bit) & 1);
        ph <<= 1; mh <<= 1;
        if (hin < 0) mh |= 1; else if (hin > 0) ph |= 1;
        pv[b] = mh | ~(xv | ph);
        mv[b] = ph & xv;
        hin = hout;
    }
    return hin;
}
/* D[i] of a column given its vertical deltas */
static inline int col_val(const uint64_t *pv, const uint64_t *mv, int i) {
    int v = 0, b = 0;
    while (i >= 64) { v += __builtin_popcountll(pv[b]) - __builtin_popcountll(mv[b]); b++; i -= 64; }
    if (i > 0) { uint64_t msk = (~0ull) >> (64 - i); v += __builtin_popcountll(pv[b] & msk) - __builtin_popcountll(mv[b] & msk); }
    return v;
}

/* ---------------- frame-level match ---------------- */
typedef struct { int ts, te, ps, pe, cost, n_ops; uint8_t *ops; } fmatch;
typedef struct { fmatch *v; int n, cap; } fvec;
static void fvec_push(fvec *f, fmatch m) {
    if (f->n == f->cap) { f->cap = f->cap ? f->cap * 2 : 4; f->v = (fmatch *)realloc(f->v, sizeof(fmatch) * f->cap); }
    f->v[f->n++] = m;
}

/* S1: walk the extended cost row and collect reported end positions */
static int local_minima(const int32_t *c, int P, int k, int *pos) {
    int n = 0, dec = 1, prev = c[0], pstart = 0;
    const int left = g_policy.flags & ORC_POL_S1_LEFT;      /* report the left instead of the right end of the plateau */
    for (int p = 1; p < P; p++) {
        int cur = c[p];
        if (cur > prev && dec && prev <= k) pos[n++] = left ? pstart : p - 1;
        if (cur < prev) dec = 1; else if (cur > prev) dec = 0;
        if (cur != prev) pstart = p;
        prev = cur;
    }
    if (dec && prev <= k) pos[n++] = left ? pstart : P - 1;
    return n;
}

/* cell accessor over either a recorded bit-vector window or a naive matrix */
typedef struct {
    int m, s;                 /* window starts at text column s */
    int edge;                 /* artificial window edge column (s if s>0), -1 if the window starts at the text start */
    const uint64_t *hpv, *hmv; int nw;  /* history: column jj -> hpv[jj*nw ..] */
    const int32_t *mat;       /* naive: mat[(j-s)*(m+1)+i] */
} cells_t;
static inline int cell(const cells_t *C, int i, int j) {
    int jj = j - C->s;
    if (C->mat) return C->mat[(size_t)jj * (C->m + 1) + i];
    return col_val(C->hpv + (size_t)jj * C->nw, C->hmv + (size_t)jj * C->nw, i);
}

/* S2/S3 traceback from cell (iend, jend); fills fm (ts, ps, ops) */
static void traceback(const cells_t *C, const uint8_t *pc, const uint8_t *tc, int iend, int jend, int has_over, fmatch *fm) {
    int i = iend, j = jend, n = 0, cap = iend + 64;
    uint8_t *rev = (uint8_t *)malloc((size_t)cap + 8);
    while (i > 0) {
        if (n >= cap) { cap *= 2; rev = (uint8_t *)realloc(rev, (size_t)cap + 8); }
        if (j == 0) {
            if (has_over) break;          /* S3: the rest of the pattern hangs over the text start */
            rev[n++] = ORC_OP_PAT; i--; continue;
        }
        if (j == C->edge) { rev[n++] = ORC_OP_PAT; i--; continue; }   /* window edge (unreachable for cost<=k) */
        int g = cell(C, i, j);
        int d = cell(C, i - 1, j - 1);
        if ((pc[i - 1] & tc[j - 1]) && d == g) { rev[n++] = ORC_OP_MATCH; i--; j--; }
        else if (d + 1 == g) { rev[n++] = ORC_OP_SUB; i--; j--; }
        else if (!(g_policy.flags & ORC_POL_S2_PAT_FIRST)) {
            if (cell(C, i, j - 1) + 1 == g) { rev[n++] = ORC_OP_TEXT; j--; }
            else { rev[n++] = ORC_OP_PAT; i--; }
        } else {
            if (cell(C, i - 1, j) + 1 == g) { rev[n++] = ORC_OP_PAT; i--; }
            else { rev[n++] = ORC_OP_TEXT; j--; }
        }
    }
    fm->ps = i; fm->ts = j; fm->n_ops = n;
    fm->ops = (uint8_t *)malloc((size_t)n + 1);
    for (int q = 0; q < n; q++) fm->ops[q] = rev[n - 1 - q];
    free(rev);
}

/* pattern machine + codes */
typedef struct { pat_t P; const uint8_t *pc; } pattern_t;

/* Extended bottom row of pattern vs text codes tc[0..n): c[0..n] (+ c[n+1..n+m] with overhang). Returns its length.
   In naive mode *mat_out receives the full matrix (caller frees). */
static int frame_costs(const pattern_t *pt, const uint8_t *tc, int n, int32_t *c, int32_t **mat_out) {
    const pat_t *P = &pt->P; const uint8_t *pc = pt->pc; int m = P->m;
    *mat_out = NULL;
    if (g_policy.use_myers) {
        uint64_t pv[NW_MAX], mv[NW_MAX] = {0};
        memcpy(pv, P->has_over ? P->pv_over : P->pv_plain, sizeof pv);
        int score = P->ov[m];
        c[0] = score;
        if (P->nw == 1) {            /* single-word fast path (same arithmetic as col_step) */
            uint64_t Pv = pv[0], Mv = 0;
            const int lb = P->last_bit;
            for (int j = 0; j < n; j++) {
                const uint64_t eq = P->eq[tc[j]][0];
                const uint64_t xv = eq | Mv;
                const uint64_t xh = (((eq & Pv) + Pv) ^ Pv) | eq;
                uint64_t ph = Mv | ~(xh | Pv), mh = Pv & xh;
                score += (int)((ph >> lb) & 1) - (int)((mh >> lb) & 1);
                ph <<= 1; mh <<= 1;
                Pv = mh | ~(xv | ph); Mv = ph & xv;
                c[j + 1] = score;
            }
            pv[0] = Pv; mv[0] = Mv;
        } else {
            for (int j = 0; j < n; j++) { score += col_step(P, pv, mv, tc[j]); c[j + 1] = score; }
        }
        if (P->has_over) for (int t = 1; t <= m; t++) c[n + t] = col_val(pv, mv, m - t) + P->ov[t];
    } else {
        int32_t *mat = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1) * (m + 1));
        for (int i = 0; i <= m; i++) mat[i] = P->ov[i];
        for (int j = 1; j <= n; j++) {
            int32_t *cur = mat + (size_t)j * (m + 1), *prv = cur - (m + 1);
            cur[0] = 0;
            for (int i = 1; i <= m; i++) {
                int d = prv[i - 1] + ((pc[i - 1] & tc[j - 1]) ? 0 : 1);
                int a = prv[i] + 1, b = cur[i - 1] + 1;
                cur[i] = d < a ? (d < b ? d : b) : (a < b ? a : b);
            }
        }
        for (int j = 0; j <= n; j++) c[j] = mat[(size_t)j * (m + 1) + m];
        if (P->has_over) for (int t = 1; t <= m; t++) c[n + t] = mat[(size_t)n * (m + 1) + (m - t)] + P->ov[t];
        *mat_out = mat;
    }
    return n + 1 + (P->has_over ? m : 0);
}

/* Build the match reported at extended position p (cost c_p) with threshold k. */
static fmatch frame_trace(const pattern_t *pt, const uint8_t *tc, int n, int k, int p, int c_p, const int32_t *mat) {
    const pat_t *P = &pt->P; int m = P->m;
    int jend = p <= n ? p : n, t = p <= n ? 0 : p - n, iend = m - t;
    fmatch fm; memset(&fm, 0, sizeof fm);
    fm.te = jend; fm.pe = iend; fm.cost = c_p;
    cells_t C; memset(&C, 0, sizeof C); C.m = m; C.edge = -1;
    uint64_t *hpv = NULL, *hmv = NULL;
    if (mat) { C.s = 0; C.mat = mat; }
    else {
        int W = m + 2 * (k < m ? k : m) + 8, s = jend - W; if (s < 0) s = 0;
        int nc = jend - s + 1;
        hpv = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nc * P->nw);
        hmv = (uint64_t *)calloc((size_t)nc * P->nw, sizeof(uint64_t));
        uint64_t pv[NW_MAX], mv[NW_MAX] = {0};
        memcpy(pv, (s == 0 && P->has_over) ? P->pv_over : P->pv_plain, sizeof pv);
        memcpy(hpv, pv, sizeof(uint64_t) * P->nw);
        for (int j = s; j < jend; j++) {
            col_step(P, pv, mv, tc[j]);
            memcpy(hpv + (size_t)(j - s + 1) * P->nw, pv, sizeof(uint64_t) * P->nw);
            memcpy(hmv + (size_t)(j - s + 1) * P->nw, mv, sizeof(uint64_t) * P->nw);
        }
        C.s = s; C.hpv = hpv; C.hmv = hmv; C.nw = P->nw; C.edge = s > 0 ? s : -1;
    }
    traceback(&C, pt->pc, tc, iend, jend, P->has_over, &fm);
    free(hpv); free(hmv);
    return fm;
}

/* Search one pattern in text codes tc[0..n) (already in the frame's orientation): all S1 minima, traced. */
static void search_frame(const pattern_t *pt, const uint8_t *tc, int n, int k, fvec *out) {
    int m = pt->P.m;
    int32_t *c = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + m + 2));
    int *pos = (int *)malloc(sizeof(int) * (size_t)(n + m + 2));
    int32_t *mat;
    int Ptot = frame_costs(pt, tc, n, c, &mat);
    int np = local_minima(c, Ptot, k, pos);
    for (int q = 0; q < np; q++) fvec_push(out, frame_trace(pt, tc, n, k, pos[q], c[pos[q]], mat));
    free(mat); free(c); free(pos);
}

/* ---------------- public sassy-level API ---------------- */
static void encode(const uint8_t *s, int n, uint8_t *codes) { for (int i = 0; i < n; i++) codes[i] = CODE[s[i]]; }

static void push_frame_matches(fvec *fv, int n, int strand, orc_match **out, int *cnt, int *cap) {
    for (int q = 0; q < fv->n; q++) {
        if (*cnt == *cap) { *cap = *cap ? *cap * 2 : 4; *out = (orc_match *)realloc(*out, sizeof(orc_match) * (size_t)*cap); }
        orc_match *o = &(*out)[(*cnt)++];
        fmatch *f = &fv->v[q];
        if (strand == ORC_FWD) { o->text_start = f->ts; o->text_end = f->te; }
        else { o->text_start = n - f->te; o->text_end = n - f->ts; }
        o->pattern_start = f->ps; o->pattern_end = f->pe; o->cost = f->cost; o->strand = strand;
        o->n_ops = f->n_ops; o->ops = f->ops;
    }
    free(fv->v); fv->v = NULL; fv->n = fv->cap = 0;
}

/* search with pre-encoded text (tc forward codes, trc = reverse-complement frame codes or NULL) */
static int search_codes(const pattern_t *pt, const uint8_t *tc, const uint8_t *trc, int n, int k, orc_match **out) {
    int cnt = 0, cap = 0; *out = NULL;
    fvec fv = {0};
    const int rc_first = (g_policy.flags & ORC_POL_S6_RC_FIRST) && trc;    /* S6 */
    for (int pass = 0; pass < 2; pass++) {
        const int strand = pass == (rc_first ? 1 : 0) ? ORC_FWD : ORC_RC;
        if (strand == ORC_RC && !trc) continue;
        search_frame(pt, strand == ORC_FWD ? tc : trc, n, k, &fv);
        push_frame_matches(&fv, n, strand, out, &cnt, &cap);
    }
    return cnt;
}

int orc_search(const uint8_t *pattern, int m, const uint8_t *text, int n, int k, float alpha, int rc, orc_match **out) {
    init_tables();
    uint8_t *pc = (uint8_t *)malloc((size_t)m + 1), *tc = (uint8_t *)malloc((size_t)n + 1), *trc = NULL;
    encode(pattern, m, pc); encode(text, n, tc);
    if (rc) { trc = (uint8_t *)malloc((size_t)n + 1); for (int j = 0; j < n; j++) trc[j] = comp_code(tc[n - 1 - j]); }
    pattern_t pt; pat_build(&pt.P, pc, m, alpha); pt.pc = pc;
    int r = search_codes(&pt, tc, trc, n, k, out);
    free(pc); free(tc); free(trc);
    return r;
}
void orc_free_matches(orc_match *ms, int n) { for (int i = 0; i < n; i++) free(ms[i].ops); free(ms); }

int orc_bottom_row(const uint8_t *pattern, int m, const uint8_t *text, int n, float alpha, int32_t *c) {
    init_tables();
    uint8_t *pc = (uint8_t *)malloc((size_t)m + 1), *tc = (uint8_t *)malloc((size_t)n + 1);
    encode(pattern, m, pc); encode(text, n, tc);
    pat_t P; pat_build(&P, pc, m, alpha);
    uint64_t pv[NW_MAX], mv[NW_MAX] = {0};
    memcpy(pv, P.has_over ? P.pv_over : P.pv_plain, sizeof pv);
    int score = P.ov[m]; c[0] = score;
    for (int j = 0; j < n; j++) { score += col_step(&P, pv, mv, tc[j]); c[j + 1] = score; }
    int tot = n + 1;
    if (P.has_over) { for (int t = 1; t <= m; t++) c[n + t] = col_val(pv, mv, m - t) + P.ov[t]; tot += m; }
    free(pc); free(tc);
    return tot;
}

/* Match::to_path (S4): one Pos per op, taken BEFORE the op is applied */
void orc_to_path(const orc_match *mt, int32_t *ij) {
    int i = mt->pattern_start, j = mt->strand == ORC_FWD ? mt->text_start : mt->text_end - 1;
    int dj = mt->strand == ORC_FWD ? 1 : -1;
    for (int q = 0; q < mt->n_ops; q++) {
        ij[2 * q] = i; ij[2 * q + 1] = j;
        switch (mt->ops[q]) {
            case ORC_OP_MATCH: case ORC_OP_SUB: i++; j += dj; break;
            case ORC_OP_TEXT: j += dj; break;
            default: i++; break;
        }
    }
}

/* cigar_parse.rs:71-82 */
int orc_get_matching_region(const orc_match *mt, int start, int end, int64_t *rs, int64_t *re) {
    int32_t *ij = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(mt->n_ops + 1));
    orc_to_path(mt, ij);
    int first = -1, last = -1;
    for (int q = 0; q < mt->n_ops; q++) if (ij[2 * q] >= start && ij[2 * q] <= end) { if (first < 0) first = q; last = q; }
    int ok = first >= 0 && last != first;     /* next() then next_back() need two distinct entries */
    if (ok) {
        int64_t a = ij[2 * first + 1], b = ij[2 * last + 1];
        if (a < 0) a = 0;
        if (b < 0) b = 0;    /* `as usize` of -1 is unreachable in practice; clamp (policy) */
        *rs = a < b ? a : b; *re = a < b ? b : a;
    }
    free(ij);
    return ok;
}

/* cigar_parse.rs:6-68 */
int orc_map_pat_to_text_with_cost(const orc_match *mt, int p_start, int p_end, int64_t out[5]) {
    int32_t *ij = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(mt->n_ops + 1));
    orc_to_path(mt, ij);
    int si = -1, ei = -1;
    for (int q = 0; q < mt->n_ops; q++) if (ij[2 * q] >= p_start && ij[2 * q] < p_end) { if (si < 0) si = q; ei = q; }
    if (si >= 0) {
        int cost = 0;
        for (int q = si; q <= ei; q++) cost += mt->ops[q] != ORC_OP_MATCH;
        out[0] = ij[2 * si]; out[1] = (int64_t)ij[2 * ei] + 1;
        out[2] = ij[2 * si + 1]; out[3] = (int64_t)ij[2 * ei + 1] + 1;
        out[4] = cost;
    }
    free(ij);
    return si >= 0;
}

/* cigar-lodhi-rs: S_3(C, 1/2) = sum over triples i1<i2<i3 of Match positions of (1/2)^(i3-i1+1)  (paper App. B) */
double orc_lodhi(const uint8_t *ops, int n_ops) {
    volatile double a1 = 0.0, a2 = 0.0, s = 0.0;   /* volatile: no contraction / reassociation */
    for (int p = 0; p < n_ops; p++) {
        if (ops[p] == ORC_OP_MATCH) {
            s = s + 0.5 * a2;
            a2 = 0.5 * (a2 + a1);
            a1 = 0.5 * (a1 + 1.0);
        } else { a2 = 0.5 * a2; a1 = 0.5 * a1; }
    }
    return s;
}

/* edit_model.rs:2-11 */
int orc_edit_cut_off(int l) {
    double a = (double)l, v = ceil(0.5100 * a - 1.7312 * sqrt(a));
    return v > 0.0 ? (int)v : 0;
}

/* ---------------- interval.rs ---------------- */
static int is_overlap(const orc_row *a, const orc_row *b, float thr) {
    int64_t s = a->read_start_flank > b->read_start_flank ? a->read_start_flank : b->read_start_flank;
    int64_t e = a->read_end_flank < b->read_end_flank ? a->read_end_flank : b->read_end_flank;
    if (e <= s) return 0;
    int64_t la = a->read_end_flank - a->read_start_flank, lb = b->read_end_flank - b->read_start_flank;
    int64_t mn = la < lb ? la : lb;
    return ((float)(e - s) / (float)mn) >= thr;
}
/* is a strictly better than b under select_best_match's comparator (interval.rs:44-79)? */
static int better(const orc_row *a, const orc_row *b) {
    int pa = a->match_type <= ORC_RTAG ? 1 : 2, pb = b->match_type <= ORC_RTAG ? 1 : 2;
    if (pa != pb) return pa < pb;
    if (pa == 1) {
        if (a->barcode_cost != b->barcode_cost) return a->barcode_cost < b->barcode_cost;
        return a->flank_cost < b->flank_cost;
    }
    return (a->read_end_flank - a->read_start_flank) > (b->read_end_flank - b->read_start_flank);
}
int orc_collapse(orc_row *rows, int n, float thr) {
    if (n == 0) return 0;
    /* stable insertion sort by read_start_flank */
    for (int i = 1; i < n; i++) {
        orc_row x = rows[i]; int j = i - 1;
        while (j >= 0 && rows[j].read_start_flank > x.read_start_flank) { rows[j + 1] = rows[j]; j--; }
        rows[j + 1] = x;
    }
    int out = 0, g0 = 0;
    for (int i = 1; i <= n; i++) {
        int joins = 0;
        if (i < n) for (int q = g0; q < i; q++) if (is_overlap(&rows[q], &rows[i], thr)) { joins = 1; break; }
        if (!joins) {
            int best = g0;       /* stable sort + first  ==  first minimal element */
            for (int q = g0 + 1; q < i; q++) if (better(&rows[q], &rows[best])) best = q;
            orc_row b = rows[best];
            rows[out++] = b;
            g0 = i;
        }
    }
    return out;
}

/* ---------------- searcher.rs ---------------- */
static int64_t rel_dist_to_end(int64_t pos, int64_t read_len) {
    if (pos < 0) return 1;
    if (pos <= read_len / 2) return pos == 0 ? 1 : pos;
    if (pos == read_len) return -1;
    return -(read_len - pos);
}

typedef struct {
    int n_groups;
    pattern_t *flank;               /* per group (built with the overhang alpha) */
    pattern_t **bar_fwd, **bar_rc;  /* per group: n_barcodes pattern machines each (no overhang) */
    uint8_t **bar_codes_fwd, **bar_codes_rc;  /* per group: n_barcodes*bar_len codes */
    uint8_t **flank_codes;
    double *perfect;
} demux_state;

static demux_state *state_build(const orc_group *groups, int n_groups, float alpha) {
    init_tables();
    demux_state *S = (demux_state *)calloc(1, sizeof *S);
    S->n_groups = n_groups;
    S->flank = (pattern_t *)calloc((size_t)n_groups, sizeof(pattern_t));
    S->bar_fwd = (pattern_t **)calloc((size_t)n_groups, sizeof(pattern_t *));
    S->bar_rc = (pattern_t **)calloc((size_t)n_groups, sizeof(pattern_t *));
    S->bar_codes_fwd = (uint8_t **)calloc((size_t)n_groups, sizeof(uint8_t *));
    S->bar_codes_rc = (uint8_t **)calloc((size_t)n_groups, sizeof(uint8_t *));
    S->flank_codes = (uint8_t **)calloc((size_t)n_groups, sizeof(uint8_t *));
    S->perfect = (double *)calloc((size_t)n_groups, sizeof(double));
    for (int g = 0; g < n_groups; g++) {
        const orc_group *G = &groups[g];
        S->flank_codes[g] = (uint8_t *)malloc((size_t)G->flank_len + 1);
        encode((const uint8_t *)G->flank, G->flank_len, S->flank_codes[g]);
        pat_build(&S->flank[g].P, S->flank_codes[g], G->flank_len, alpha); S->flank[g].pc = S->flank_codes[g];
        size_t tot = (size_t)G->n_barcodes * G->bar_len;
        S->bar_codes_fwd[g] = (uint8_t *)malloc(tot + 1);
        S->bar_codes_rc[g] = (uint8_t *)malloc(tot + 1);
        S->bar_fwd[g] = (pattern_t *)calloc((size_t)G->n_barcodes, sizeof(pattern_t));
        S->bar_rc[g] = (pattern_t *)calloc((size_t)G->n_barcodes, sizeof(pattern_t));
        for (int b = 0; b < G->n_barcodes; b++) {
            for (int i = 0; i < G->bar_len; i++) {
                uint8_t ch = (uint8_t)G->barcodes[(size_t)b * G->bar_len + i];
                S->bar_codes_fwd[g][(size_t)b * G->bar_len + i] = CODE[ch];
                /* explicit reverse complement of the padded barcode (barcodes.rs:85-88) */
                S->bar_codes_rc[g][(size_t)b * G->bar_len + (G->bar_len - 1 - i)] = CODE[RCCHAR[ch]];
            }
            S->bar_fwd[g][b].pc = S->bar_codes_fwd[g] + (size_t)b * G->bar_len;
            S->bar_rc[g][b].pc = S->bar_codes_rc[g] + (size_t)b * G->bar_len;
            pat_build(&S->bar_fwd[g][b].P, S->bar_fwd[g][b].pc, G->bar_len, -1.0f);
            pat_build(&S->bar_rc[g][b].P, S->bar_rc[g][b].pc, G->bar_len, -1.0f);
        }
        /* get_perfect_match_score (searcher.rs:229-239): un-clamped pad width */
        int l = G->pad1 - G->pad0;
        uint8_t *ops = (uint8_t *)calloc((size_t)(l > 0 ? l : 1), 1);
        S->perfect[g] = orc_lodhi(ops, l > 0 ? l : 0);
        free(ops);
    }
    return S;
}
static void state_free(demux_state *S) {
    for (int g = 0; g < S->n_groups; g++) {
        free(S->bar_codes_fwd[g]); free(S->bar_codes_rc[g]); free(S->flank_codes[g]); free(S->bar_fwd[g]); free(S->bar_rc[g]);
    }
    free(S->bar_codes_fwd); free(S->bar_codes_rc); free(S->flank_codes); free(S->perfect);
    free(S->bar_fwd); free(S->bar_rc); free(S->flank); free(S);
}

static void flank_row(orc_row *r, uint32_t read_idx, int n, int g, const orc_group *G, const orc_match *fm) {
    memset(r, 0, sizeof *r);
    r->read_idx = read_idx; r->read_len = (uint32_t)n;
    r->read_start_bar = fm->text_start; r->read_end_bar = fm->text_end;
    r->read_start_flank = fm->text_start; r->read_end_flank = fm->text_end;
    r->bar_start = 0; r->bar_end = 0;
    r->match_type = (uint8_t)(G->match_type == ORC_FTAG ? ORC_FFLANK : ORC_RFLANK);
    r->flank_cost = fm->cost; r->barcode_cost = G->bar_len; r->label_idx = -1; r->group_idx = g;
    r->strand = (uint8_t)fm->strand;
    r->rel_dist_to_end = rel_dist_to_end(fm->text_start, n);
}

typedef struct { double norm, raw; orc_match m; int idx; } scored_t;

/* searcher.rs:267-426 for one flank match; appends exactly one row */
static void barcode_stage(const demux_state *S, const orc_group *G, int g, const orc_params *prm, uint32_t read_idx,
                          const uint8_t *tc, int n, const orc_match *fm, int64_t rs, int64_t re, orc_row *row) {
    const uint8_t *region = tc + rs; int rn = (int)(re - rs);
    int L = G->bar_len, nb = G->n_barcodes;
    int k1 = (int)((float)L * 0.4f);           /* searcher.rs:460 */
    const pattern_t *pats = fm->strand == ORC_FWD ? S->bar_fwd[g] : S->bar_rc[g];
    orc_match *best = (orc_match *)calloc((size_t)nb, sizeof(orc_match));
    uint8_t *have = (uint8_t *)calloc((size_t)nb, 1);
    int32_t *c = (int32_t *)malloc(sizeof(int32_t) * (size_t)(rn + L + 2) * (size_t)nb);
    int32_t **mats = (int32_t **)calloc((size_t)nb, sizeof(int32_t *));
    int *pos = (int *)malloc(sizeof(int) * (size_t)(rn + L + 2));
    int *bestpos = (int *)malloc(sizeof(int) * (size_t)nb);
    int matched = 0, kuse = k1;
    for (int b = 0; b < nb; b++) frame_costs(&pats[b], region, rn, c + (size_t)b * (rn + L + 2), &mats[b]);
    for (int pass = 0; pass < 2; pass++) {
        int k = pass == 0 ? k1 : L;
        if (pass == 1 && !(matched <= 1 && k1 < L)) break;      /* searcher.rs:303-306 */
        matched = 0; kuse = k;
        for (int b = 0; b < nb; b++) {
            /* search_encoded_patterns at this k, keep the lowest-cost match, first seen wins (searcher.rs:294-300) */
            const int32_t *cb = c + (size_t)b * (rn + L + 2);
            int np = local_minima(cb, rn + 1, k, pos), bi = -1;
            for (int q = 0; q < np; q++)      /* S5: the first (default) or the last of equal lowest-cost minima */
                if (bi < 0 || cb[pos[q]] < cb[pos[bi]] || ((g_policy.flags & ORC_POL_S5_LAST) && cb[pos[q]] == cb[pos[bi]])) bi = q;
            have[b] = bi >= 0; if (bi >= 0) { bestpos[b] = pos[bi]; matched++; }
        }
    }
    for (int b = 0; b < nb; b++) {
        if (have[b]) {
            const int32_t *cb = c + (size_t)b * (rn + L + 2);
            fmatch f = frame_trace(&pats[b], region, rn, kuse, bestpos[b], cb[bestpos[b]], mats[b]);
            best[b].text_start = f.ts; best[b].text_end = f.te; best[b].pattern_start = f.ps; best[b].pattern_end = f.pe;
            best[b].cost = f.cost; best[b].strand = ORC_FWD; best[b].n_ops = f.n_ops; best[b].ops = f.ops;
        }
        free(mats[b]);
    }
    free(c); free(mats); free(pos); free(bestpos);
    if (matched == 0) { flank_row(row, read_idx, n, g, G, fm); free(best); free(have); return; }
    scored_t *sc = (scored_t *)malloc(sizeof(scored_t) * (size_t)matched);
    int ns = 0;
    for (int b = 0; b < nb; b++) if (have[b]) {
        best[b].strand = fm->strand;            /* searcher.rs:333 */
        double s = orc_lodhi(best[b].ops, best[b].n_ops);
        sc[ns].raw = s; sc[ns].norm = S->perfect[g] > 0.0 ? s / S->perfect[g] : 0.0;
        sc[ns].m = best[b]; sc[ns].idx = b; ns++;
    }
    /* stable sort, high to low (searcher.rs:377) */
    for (int i = 1; i < ns; i++) {
        scored_t x = sc[i]; int j = i - 1;
        while (j >= 0 && sc[j].norm < x.norm) { sc[j + 1] = sc[j]; j--; }
        sc[j + 1] = x;
    }
    int64_t mp[5];
    int ok = orc_map_pat_to_text_with_cost(&sc[0].m, G->bar0 - G->pad0, G->bar1 - G->pad0, mp);
    double top = sc[0].norm;
    int valid = top >= prm->min_score;
    if (ns > 1) valid = valid && (top - sc[1].norm) >= prm->min_score_diff;
    if (ok && valid) {
        memset(row, 0, sizeof *row);
        row->read_idx = read_idx; row->read_len = (uint32_t)n;
        row->read_start_bar = rs + mp[2]; row->read_end_bar = rs + mp[3];
        row->read_start_flank = fm->text_start; row->read_end_flank = fm->text_end;
        row->bar_start = rs + mp[0]; row->bar_end = rs + mp[1];
        row->match_type = (uint8_t)G->match_type;
        row->flank_cost = fm->cost; row->barcode_cost = (int32_t)mp[4];
        row->label_idx = sc[0].idx; row->group_idx = g; row->strand = (uint8_t)sc[0].m.strand;
        row->rel_dist_to_end = rel_dist_to_end(fm->text_start, n);
    } else {
        /* !ok is the reference's `expect` panic (searcher.rs:388); the restatement degrades to a flank row */
        flank_row(row, read_idx, n, g, G, fm);
    }
    for (int q = 0; q < ns; q++) free(sc[q].m.ops);
    free(sc); free(best); free(have);
}

static int demux_codes(const demux_state *S, const orc_group *groups, const orc_params *prm, uint32_t read_idx,
                       const uint8_t *tc, const uint8_t *trc, int n, orc_row *out, int cap, int32_t *hits6, int64_t *nh,
                       int64_t hcap) {
    int nr = 0;
    for (int g = 0; g < S->n_groups; g++) {
        const orc_group *G = &groups[g];
        orc_match *fms = NULL;
        int nf = search_codes(&S->flank[g], tc, trc, n, G->k_flank, &fms);
        for (int f = 0; f < nf; f++) {
            orc_match *fm = &fms[f];
            if (hits6) {
                if (*nh < hcap) {
                    int32_t *h = hits6 + 6 * (*nh);
                    h[0] = (int32_t)read_idx; h[1] = g; h[2] = fm->strand; h[3] = fm->text_start; h[4] = fm->text_end; h[5] = fm->cost;
                }
                (*nh)++;
                continue;
            }
            int64_t rs, re;
            if (!orc_get_matching_region(fm, G->bar0, G->bar1, &rs, &re)) continue;   /* searcher.rs:445-449 */
            rs = rs > PADDING ? rs - PADDING : 0;                                       /* :453 */
            re = re + PADDING < n ? re + PADDING : n;                                   /* :454 */
            if (re < rs) re = rs;
            if (nr >= cap) { orc_free_matches(fms, nf); return -1; }
            barcode_stage(S, G, g, prm, read_idx, tc, n, fm, rs, re, &out[nr++]);
        }
        orc_free_matches(fms, nf);
    }
    return hits6 ? 0 : orc_collapse(out, nr, 0.8f);
}

int orc_demux(const orc_group *groups, int n_groups, const orc_params *prm, uint32_t read_idx, const uint8_t *read, int n,
              orc_row *out, int cap) {
    demux_state *S = state_build(groups, n_groups, prm->alpha);
    uint8_t *tc = (uint8_t *)malloc((size_t)n + 1), *trc = (uint8_t *)malloc((size_t)n + 1);
    encode(read, n, tc);
    for (int j = 0; j < n; j++) trc[j] = comp_code(tc[n - 1 - j]);
    int r = demux_codes(S, groups, prm, read_idx, tc, trc, n, out, cap, NULL, NULL, 0);
    free(tc); free(trc); state_free(S);
    return r;
}

/* per-read row / hit capacity inside the batch driver: 64, or ORC_PER_READ from the environment (fuzzing with repeat-heavy reads) */
static int per_read(void) { const char *e = getenv("ORC_PER_READ"); int v = e ? atoi(e) : 64; return v < 8 ? 8 : (v > 65536 ? 65536 : v); }
typedef struct {
    const demux_state *S; const orc_group *groups; const orc_params *prm; const uint8_t *bases; const uint64_t *offsets;
    uint32_t base, nb; orc_row *tmp; int32_t *htmp; int *cnt; volatile uint32_t *next; int per;
} job_t;
static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    const int PER = J->per;
    for (;;) {
        uint32_t q0 = __sync_fetch_and_add(J->next, 16u);
        if (q0 >= J->nb) break;
        uint32_t q1 = q0 + 16 < J->nb ? q0 + 16 : J->nb;
        for (uint32_t q = q0; q < q1; q++) {
            uint32_t r = J->base + q;
            int n = (int)(J->offsets[r + 1] - J->offsets[r]);
            uint8_t *tc = (uint8_t *)malloc((size_t)n + 1), *trc = (uint8_t *)malloc((size_t)n + 1);
            encode(J->bases + J->offsets[r], n, tc);
            for (int j = 0; j < n; j++) trc[j] = comp_code(tc[n - 1 - j]);
            if (J->htmp) {
                int64_t nh = 0;
                demux_codes(J->S, J->groups, J->prm, r, tc, trc, n, NULL, 0, J->htmp + 6 * (size_t)q * PER, &nh, PER);
                J->cnt[q] = nh > PER ? -1 : (int)nh;
            } else {
                J->cnt[q] = demux_codes(J->S, J->groups, J->prm, r, tc, trc, n, J->tmp + (size_t)q * PER, PER, NULL, NULL, 0);
            }
            free(tc); free(trc);
        }
    }
    return NULL;
}

static int64_t batch_impl(const orc_group *groups, int n_groups, const orc_params *prm, const uint8_t *bases,
                          const uint64_t *offsets, uint32_t n_reads, int n_threads, orc_row *out, int64_t cap,
                          int32_t *hits6, int64_t hcap) {
    demux_state *S = state_build(groups, n_groups, prm->alpha);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    /* reads are processed in parallel into per-read slots, then concatenated in input order */
    int64_t total = 0; int fail = 0;
    const int PER = per_read();
    uint32_t CH = (uint32_t)(8192 * 64 / PER); if (CH < 16) CH = 16;
    orc_row *tmp = hits6 ? NULL : (orc_row *)malloc(sizeof(orc_row) * (size_t)CH * PER);
    int32_t *htmp = hits6 ? (int32_t *)malloc(sizeof(int32_t) * 6 * (size_t)CH * PER) : NULL;
    int *cnt = (int *)malloc(sizeof(int) * CH);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (uint32_t base = 0; base < n_reads && !fail; base += CH) {
        uint32_t nb = n_reads - base < CH ? n_reads - base : CH;
        volatile uint32_t next = 0;
        job_t J = {S, groups, prm, bases, offsets, base, nb, tmp, htmp, cnt, &next, PER};
        for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, worker, &J);
        worker(&J);
        for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
        for (uint32_t q = 0; q < nb; q++) {
            if (cnt[q] < 0) { fail = 1; break; }
            if (hits6) {
                if (total + cnt[q] > hcap) { fail = 1; break; }
                memcpy(hits6 + 6 * total, htmp + 6 * (size_t)q * PER, sizeof(int32_t) * 6 * (size_t)cnt[q]);
            } else {
                if (total + cnt[q] > cap) { fail = 1; break; }
                memcpy(out + total, tmp + (size_t)q * PER, sizeof(orc_row) * (size_t)cnt[q]);
            }
            total += cnt[q];
        }
    }
    free(th); free(tmp); free(htmp); free(cnt); state_free(S);
    return fail ? -1 : total;
}

int64_t orc_demux_batch(const orc_group *groups, int n_groups, const orc_params *prm, const uint8_t *bases,
                        const uint64_t *offsets, uint32_t n_reads, int n_threads, orc_row *out, int64_t cap) {
    return batch_impl(groups, n_groups, prm, bases, offsets, n_reads, n_threads, out, cap, NULL, 0);
}
int64_t orc_flank_hits_batch(const orc_group *groups, int n_groups, const orc_params *prm, const uint8_t *bases,
                             const uint64_t *offsets, uint32_t n_reads, int n_threads, int32_t *out6, int64_t cap) {
    return batch_impl(groups, n_groups, prm, bases, offsets, n_reads, n_threads, NULL, 0, out6, cap);
}
