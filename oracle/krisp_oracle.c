/*
 * krisp_oracle.c — CPU restatement ("port") of krisp_fasta's diagnostic-region search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library, and
 * only as the checker / the timed CPU baseline.  The product (krisp_b200) never links it.
 *
 * It follows the reference's algorithm stage by stage, on the same data
 * representation (ASCII k-mers, bytewise ordering), so that letters outside ACGT
 * behave as in the reference.  File:line citations are into /root/reference/src/krisp.
 *
 *   stage A  kstream/kstream.py:617-642 (_kmers), :734-766 (_omitsoft/_mapsoft),
 *            :644-677 (_complements, COMP_MAP :11-18), :715-732 (_disallow "Nn"),
 *            :805-832 (_split [L,-R]; -0 is a positive split => quirk S9)
 *   stage B  kstream/kstream.py:83-119  LC_ALL=C sort -t, -k1,1 -k3,3 (+ whole-line tiebreak)
 *   stage C  krisp_fasta/shared.py:210-347,442-475 and intersectAmplicons.py:232-310
 *            (tree of pairwise intersections on (left,right); pairs popped from the list end)
 *   stage D  krisp_fasta/filterAlignments.py:4-28, Amplicon.py:495-521 (only when D > 0,
 *            krisp_fasta.py:265)
 *   rows     krisp_fasta/Amplicon.py:663-671, :550-558, :42-66, iupac_key :10-12
 *
 * Parity pinning: checked against tests/golden/golden.json (outputs of the unmodified
 * reference) by tests/test_oracle.py.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KO_OK 0
#define KO_EINVAL -1
#define KO_ENOMEM -2
#define KO_EKEY -3 /* the reference would raise KeyError (COMP_MAP / iupac_key) */

/* ---- stage A tables -------------------------------------------------------------------- */
static uint8_t COMP[256];
static int comp_ready = 0;
static void init_comp(void) {
    if (comp_ready) return;
    const char* a = "ATatGCgcRYryMKmkSWswBVbvDHdhNn";
    const char* b = "TAtaCGcgYRyrKMkmSWswVBvbHDhdNn";
    memset(COMP, 0, sizeof COMP);
    for (int i = 0; a[i]; i++) COMP[(uint8_t)a[i]] = (uint8_t)b[i];
    comp_ready = 1;
}

typedef struct {
    uint8_t* rec;   /* n records of `stride` bytes: left|right|mid (k bytes) + file id (2 bytes BE) */
    uint64_t n;
} table_t;

typedef struct {
    int k, L, D, R;       /* effective field lengths: left L, mid D, right R */
    int stride;
} layout_t;

/* Memory-lean runs of the checker on big panels: only k-mers whose (left,right) key hashes into part g_part of g_nparts are kept.
 * Every rule of the search is local to one key (S5-S8), so the union of the rows over all parts is the full answer. */
static int g_part = 0, g_nparts = 1;
void ko_set_key_partition(int part, int nparts) { g_nparts = nparts > 1 ? nparts : 1; g_part = (part >= 0 && part < g_nparts) ? part : 0; }
static uint32_t key_hash(const uint8_t* a, int na, const uint8_t* b, int nb) {
    uint32_t h = 2166136261u;
    for (int i = 0; i < na; i++) { h ^= a[i]; h *= 16777619u; }
    for (int i = 0; i < nb; i++) { h ^= b[i]; h *= 16777619u; }
    return h ^ (h >> 15);
}

static int cmp_stride;
#pragma omp threadprivate(cmp_stride)
static int cmp_rec(const void* a, const void* b) { return memcmp(a, b, (size_t)cmp_stride); }

/* Extract + sort one file.  Returns KO_* ; *count = number of emitted k-mer lines (kstream.write's return). */
static int extract_file(const uint8_t* bases, const uint64_t* rec_off, const int32_t* rec_file, int nrec, int file,
                        const layout_t* lo, int omit_soft, table_t* out, uint64_t* count) {
    const int k = lo->k, L = lo->L, D = lo->D, R = lo->R, S = lo->stride;
    uint64_t cap = 0;
    for (int r = 0; r < nrec; r++) {
        if (rec_file[r] != file) continue;
        uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len >= (uint64_t)k) cap += 2 * (len - k + 1);
    }
    uint8_t* buf = (uint8_t*)malloc(cap * S + 1);
    if (!buf) return KO_ENOMEM;
    uint64_t n = 0;
    uint8_t* w = (uint8_t*)malloc(2 * (size_t)k + 2);
    if (!w) { free(buf); return KO_ENOMEM; }
    uint8_t* x = w + k;
    int rc_err = KO_OK;
    for (int r = 0; r < nrec && rc_err == KO_OK; r++) {
        if (rec_file[r] != file) continue;
        const uint8_t* s = bases + rec_off[r];
        int64_t len = (int64_t)(rec_off[r + 1] - rec_off[r]);
        /* running counts over the current window: lowercase letters, uppercase letters, N/n */
        int64_t n_lower = 0, n_upper = 0, n_n = 0;
        for (int64_t i = 0; i < len; i++) {
            uint8_t c = s[i];
            n_lower += (c >= 'a' && c <= 'z'); n_upper += (c >= 'A' && c <= 'Z'); n_n += (c == 'N' || c == 'n');
            if (i >= k) {
                uint8_t d = s[i - k];
                n_lower -= (d >= 'a' && d <= 'z'); n_upper -= (d >= 'A' && d <= 'Z'); n_n -= (d == 'N' || d == 'n');
            }
            if (i < k - 1) continue;
            const uint8_t* p = s + (i - k + 1);
            if (omit_soft) {                       /* str.isupper(): >=1 cased char, none lowercase (kstream.py:749) */
                if (n_lower > 0 || n_upper == 0) continue;
                memcpy(w, p, (size_t)k);
            } else {                               /* str.upper() (kstream.py:766) */
                for (int j = 0; j < k; j++) { uint8_t c2 = p[j]; w[j] = (c2 >= 'a' && c2 <= 'z') ? (uint8_t)(c2 - 32) : c2; }
            }
            /* reverse complement (kstream.py:658); KeyError on anything outside COMP_MAP */
            for (int j = 0; j < k; j++) {
                uint8_t c2 = COMP[w[k - 1 - j]];
                if (!c2) { rc_err = KO_EKEY; break; }
                x[j] = c2;
            }
            if (rc_err != KO_OK) break;
            if (n_n > 0) continue;                  /* _disallow("Nn") after the complement; N<->N so both strands drop */
            for (int strand = 0; strand < 2; strand++) {
                const uint8_t* q = strand ? x : w;
                if (g_nparts > 1 && (int)(key_hash(q, L, q + L + D, R) % (uint32_t)g_nparts) != g_part) continue;
                uint8_t* o = buf + n * S;
                memcpy(o, q, (size_t)L);                        /* left  */
                memcpy(o + L, q + L + D, (size_t)R);            /* right */
                memcpy(o + L + R, q + L, (size_t)D);            /* mid   */
                o[k] = (uint8_t)(file >> 8); o[k + 1] = (uint8_t)(file & 255);
                n++;
            }
        }
    }
    free(w);
    if (rc_err != KO_OK) { free(buf); return rc_err; }
    if (g_nparts > 1) { uint8_t* sh = (uint8_t*)realloc(buf, n * S + 1); if (sh) buf = sh; }
    cmp_stride = S;
    qsort(buf, n, (size_t)S, cmp_rec);              /* stage B */
    out->rec = buf; out->n = n; *count = n;
    return KO_OK;
}

/* stage C: intersection of two sorted tables on (left,right) = first L+R bytes (shared.py:321-347). */
static int intersect(const table_t* a, const table_t* b, const layout_t* lo, table_t* out) {
    const int S = lo->stride, K = lo->L + lo->R;
    uint8_t* buf = (uint8_t*)malloc((a->n + b->n) * S + 1);
    if (!buf) return KO_ENOMEM;
    uint64_t i = 0, j = 0, n = 0;
    while (i < a->n && j < b->n) {
        const uint8_t* pa = a->rec + i * S; const uint8_t* pb = b->rec + j * S;
        int c = memcmp(pa, pb, (size_t)K);
        uint64_t ie = i, je = j;
        if (c <= 0) { while (ie < a->n && memcmp(a->rec + ie * S, pa, (size_t)K) == 0) ie++; }
        if (c >= 0) { while (je < b->n && memcmp(b->rec + je * S, pb, (size_t)K) == 0) je++; }
        if (c == 0) {                               /* key in both: union of the amplicons, kept sorted */
            uint64_t x = i, y = j;
            while (x < ie || y < je) {
                int take_a = (y >= je) || (x < ie && memcmp(a->rec + x * S, b->rec + y * S, (size_t)S) <= 0);
                memcpy(buf + n * S, take_a ? a->rec + x * S : b->rec + y * S, (size_t)S);
                n++; if (take_a) x++; else y++;
            }
        }
        if (c <= 0) i = ie;
        if (c >= 0) j = je;
    }
    out->rec = buf; out->n = n;
    return KO_OK;
}

/* ---- rows ------------------------------------------------------------------------------- */
static int iupac_letter(unsigned mask) { /* bit0 A, bit1 C, bit2 G, bit3 T ; Amplicon.py:10-12 */
    static const char t[16] = {0, 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};
    return t[mask & 15];
}
static int base_bit(uint8_t c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 0; }

typedef struct { char* p; size_t n, cap; } sbuf_t;
static int sb_put(sbuf_t* b, const void* s, size_t len) {
    if (b->n + len + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 4096; while (nc < b->n + len + 1) nc *= 2;
        char* q = (char*)realloc(b->p, nc); if (!q) return KO_ENOMEM; b->p = q; b->cap = nc;
    }
    memcpy(b->p + b->n, s, len); b->n += len; b->p[b->n] = 0; return KO_OK;
}
static int cmp_str(const void* a, const void* b) { return strcmp(*(const char* const*)a, *(const char* const*)b); }

void ko_free(void* p) { free(p); }

int ko_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * bases/rec_off/rec_file : nrec FASTA records (already parsed like kstream._parse_FASTA), record r belongs to file rec_file[r]
 * is_ingroup[f]          : simplename(file f) is in the ingroup label set (membership is by LABEL, krisp_fasta.py:267)
 * labels[f]              : simplename(file f) (only for the interchange text; may be NULL)
 * have_outgroup          : --outgroup was given (krisp_fasta.py:282) => consensus over ingroup sequences
 * rows_out               : malloc'd, sorted rows joined by '\n' (no header); groups_out: malloc'd interchange text of the
 *                          surviving groups (filtered.txt content, one blank line between groups), may be NULL
 * counts[f]              : stage-A line count per file ("Extracted and sorted N k-kmers", krisp_fasta.py:61)
 */
int ko_search(const uint8_t* bases, const uint64_t* rec_off, const int32_t* rec_file, int nrec,
              int nfiles, const uint8_t* is_ingroup, const char* const* labels, int have_outgroup,
              int L0, int D0, int R0, int omit_soft, int threads,
              char** rows_out, uint64_t* nrows_out, char** groups_out, uint64_t* counts) {
    if (L0 < 0 || D0 < 0 || R0 < 0 || L0 + D0 + R0 <= 0 || nfiles <= 0 || nfiles > 65535) return KO_EINVAL;
    init_comp();
    layout_t lo;
    lo.k = L0 + D0 + R0;
    /* quirk S9 (kstream.py:824-830): split=[L,-0] => fields (left, '', rest) */
    if (R0 == 0) { lo.L = L0; lo.D = 0; lo.R = D0; } else { lo.L = L0; lo.D = D0; lo.R = R0; }
    lo.stride = lo.k + 2;
    const int run_filter = D0 > 0;                  /* krisp_fasta.py:265 uses the command-line values */
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    table_t* tabs = (table_t*)calloc((size_t)nfiles, sizeof(table_t));
    if (!tabs) return KO_ENOMEM;
    int err = KO_OK;
#pragma omp parallel for schedule(dynamic, 1)
    for (int f = 0; f < nfiles; f++) {
        uint64_t c = 0;
        int e = extract_file(bases, rec_off, rec_file, nrec, f, &lo, omit_soft, &tabs[f], &c);
        if (counts) counts[f] = c;
        if (e != KO_OK) {
#pragma omp critical
            err = e;
        }
    }
    /* stage C: mergeFiles (intersectAmplicons.py:256-310): rounds of pairs popped from the list end */
    int* list = (int*)malloc(sizeof(int) * (size_t)nfiles);
    int nlist = nfiles;
    for (int i = 0; i < nfiles; i++) list[i] = i;
    while (err == KO_OK && nlist > 1) {
        int jobs = nlist / 2;
        int* res = (int*)malloc(sizeof(int) * (size_t)jobs);
#pragma omp parallel for schedule(dynamic, 1)
        for (int j = 0; j < jobs; j++) {
            int f0 = list[nlist - 1 - 2 * j], f1 = list[nlist - 2 - 2 * j];
            table_t t; t.rec = NULL; t.n = 0;
            int e = intersect(&tabs[f0], &tabs[f1], &lo, &t);
            free(tabs[f0].rec); free(tabs[f1].rec); tabs[f1].rec = NULL;
            tabs[f0] = t; res[j] = f0;
            if (e != KO_OK) {
#pragma omp critical
                err = e;
            }
        }
        int rem = nlist - 2 * jobs;                 /* files = results + files */
        int keep = rem ? list[0] : -1;
        for (int j = 0; j < jobs; j++) list[j] = res[j];
        if (rem) list[jobs] = keep;
        nlist = jobs + rem;
        free(res);
    }
    sbuf_t rows = {0, 0, 0}, groups = {0, 0, 0};
    char** rowv = NULL; uint64_t nrow = 0, rowcap = 0;
    if (err == KO_OK) {
        const table_t* T = &tabs[list[0]];
        const int S = lo.stride, K = lo.L + lo.R, L = lo.L, D = lo.D, R = lo.R, k = lo.k;
        uint8_t* in_m = (uint8_t*)calloc((size_t)(D + 1), 32), *out_m = (uint8_t*)calloc((size_t)(D + 1), 32);
        char* row = (char*)malloc((size_t)k + 8);
        uint64_t i = 0;
        while (i < T->n && err == KO_OK) {
            const uint8_t* g = T->rec + i * S;
            uint64_t e = i;
            while (e < T->n && memcmp(T->rec + e * S, g, (size_t)K) == 0) e++;
            /* distinct sequences = runs of equal first k bytes */
            uint64_t ndist = 0;
            for (uint64_t x = i; x < e; x++) if (x == i || memcmp(T->rec + x * S, T->rec + (x - 1) * S, (size_t)k) != 0) ndist++;
            int keep = 1;
            if (run_filter) {                        /* ingroupUniqueColumns, Amplicon.py:495-521 */
                memset(in_m, 0, (size_t)(D + 1) * 32); memset(out_m, 0, (size_t)(D + 1) * 32);
                for (uint64_t x = i; x < e; x++) {
                    const uint8_t* r = T->rec + x * S; int f = (r[k] << 8) | r[k + 1];
                    uint8_t* m = is_ingroup[f] ? in_m : out_m;
                    for (int c = 0; c < D; c++) m[c * 32 + (r[K + c] >> 3)] |= (uint8_t)(1u << (r[K + c] & 7));
                }
                keep = 0;
                for (int c = 0; c < D && !keep; c++) {
                    int dis = 1;
                    for (int b = 0; b < 32; b++) if (in_m[c * 32 + b] & out_m[c * 32 + b]) { dis = 0; break; }
                    keep = dis;
                }
            }
            if (keep) {
                /* render_csv: consensus over all if 1 amplicon or no ingroup passed, else over amplicons whose labels
                   are all ingroup (Amplicon.py:550-558, 663-671) */
                unsigned colmask[3] = {0, 0, 0}; (void)colmask;
                int pos = 0, nsel = 0;
                unsigned* cm = (unsigned*)calloc((size_t)k, sizeof(unsigned));
                uint64_t x = i;
                while (x < e) {
                    uint64_t y = x; int all_in = 1;
                    while (y < e && memcmp(T->rec + y * S, T->rec + x * S, (size_t)k) == 0) {
                        int f = (T->rec[y * S + k] << 8) | T->rec[y * S + k + 1];
                        if (!is_ingroup[f]) all_in = 0;
                        y++;
                    }
                    if (ndist == 1 || !have_outgroup || all_in) {
                        nsel++;
                        for (int c = 0; c < k; c++) {
                            uint8_t ch = T->rec[x * S + c];
                            int bit = base_bit(ch);
                            if (ch == 'N' || ch == '*' || ch == '?') bit = 15;
                            else if (!bit) bit = 16;   /* iupac_key KeyError */
                            cm[c] |= (unsigned)bit;
                        }
                    }
                    x = y;
                }
                if (nsel == 0) err = KO_EKEY;       /* max([]) ValueError in collapse_to_iupac */
                for (int c = 0; c < k && err == KO_OK; c++) if (cm[c] & 16) err = KO_EKEY;
                if (err == KO_OK) {
                    for (int c = 0; c < L; c++) row[pos++] = (char)iupac_letter(cm[c]);
                    row[pos++] = ',';
                    for (int c = 0; c < D; c++) row[pos++] = (char)iupac_letter(cm[K + c]);
                    row[pos++] = ',';
                    for (int c = 0; c < R; c++) row[pos++] = (char)iupac_letter(cm[L + c]);
                    row[pos] = 0;
                    if (nrow == rowcap) { rowcap = rowcap ? rowcap * 2 : 1024; rowv = (char**)realloc(rowv, rowcap * sizeof(char*)); }
                    rowv[nrow++] = strdup(row);
                    if (groups_out) {                /* interchange text: left,mid,right,label(n);... (Amplicon.py:330-348) */
                        x = i;
                        while (x < e && err == KO_OK) {
                            uint64_t y = x;
                            const uint8_t* r = T->rec + x * S;
                            sb_put(&groups, r, (size_t)L); sb_put(&groups, ",", 1);
                            sb_put(&groups, r + K, (size_t)D); sb_put(&groups, ",", 1);
                            sb_put(&groups, r + L, (size_t)R); sb_put(&groups, ",", 1);
                            /* label multiset: counts per label string, sorted by name */
                            int nl = 0; const char** ln = (const char**)malloc(sizeof(char*) * (size_t)(e - x + 1));
                            while (y < e && memcmp(T->rec + y * S, r, (size_t)k) == 0) {
                                int f = (T->rec[y * S + k] << 8) | T->rec[y * S + k + 1];
                                ln[nl++] = labels ? labels[f] : "";
                                y++;
                            }
                            qsort(ln, (size_t)nl, sizeof(char*), cmp_str);
                            for (int a = 0; a < nl;) {
                                int b = a; while (b < nl && strcmp(ln[b], ln[a]) == 0) b++;
                                char tmp[64];
                                if (a) sb_put(&groups, ";", 1);
                                sb_put(&groups, ln[a], strlen(ln[a]));
                                if (b - a > 1) { snprintf(tmp, sizeof tmp, "(%d)", b - a); sb_put(&groups, tmp, strlen(tmp)); }
                                a = b;
                            }
                            free(ln);
                            sb_put(&groups, "\n", 1);
                            x = y;
                        }
                        sb_put(&groups, "\n", 1);
                    }
                }
                free(cm);
            }
            i = e;
        }
        free(in_m); free(out_m); free(row);
    }
    for (int f = 0; f < nfiles; f++) free(tabs[f].rec);
    free(tabs); free(list);
    if (err == KO_OK) {
        qsort(rowv, nrow, sizeof(char*), cmp_str);
        for (uint64_t r = 0; r < nrow; r++) { sb_put(&rows, rowv[r], strlen(rowv[r])); sb_put(&rows, "\n", 1); }
        if (!rows.p) sb_put(&rows, "", 0);
        if (groups_out && !groups.p) sb_put(&groups, "", 0);
        *rows_out = rows.p; *nrows_out = nrow;
        if (groups_out) *groups_out = groups.p;
    } else {
        free(rows.p); free(groups.p);
    }
    for (uint64_t r = 0; r < nrow; r++) free(rowv[r]);
    free(rowv);
    return err;
}

/* Stage A+B of ONE file as the reference's text table (left,mid,right lines, GNU-sort order). */
int ko_table(const uint8_t* bases, const uint64_t* rec_off, int nrec, int L0, int D0, int R0, int omit_soft,
             char** text_out, uint64_t* nlines) {
    if (L0 < 0 || D0 < 0 || R0 < 0 || L0 + D0 + R0 <= 0) return KO_EINVAL;
    init_comp();
    layout_t lo; lo.k = L0 + D0 + R0;
    if (R0 == 0) { lo.L = L0; lo.D = 0; lo.R = D0; } else { lo.L = L0; lo.D = D0; lo.R = R0; }
    lo.stride = lo.k + 2;
    int32_t* rf = (int32_t*)calloc((size_t)nrec + 1, sizeof(int32_t));
    table_t t; uint64_t c = 0;
    int e = extract_file(bases, rec_off, rf, nrec, 0, &lo, omit_soft, &t, &c);
    free(rf);
    if (e != KO_OK) return e;
    size_t ll = (size_t)lo.k + 3;
    char* out = (char*)malloc(t.n * ll + 1);
    if (!out) { free(t.rec); return KO_ENOMEM; }
    for (uint64_t i = 0; i < t.n; i++) {
        const uint8_t* r = t.rec + i * lo.stride; char* o = out + i * ll;
        memcpy(o, r, (size_t)lo.L); o[lo.L] = ',';
        memcpy(o + lo.L + 1, r + lo.L + lo.R, (size_t)lo.D); o[lo.L + 1 + lo.D] = ',';
        memcpy(o + lo.L + 2 + lo.D, r + lo.L, (size_t)lo.R); o[ll - 1] = '\n';
    }
    out[t.n * ll] = 0;
    free(t.rec);
    *text_out = out; *nlines = t.n;
    return KO_OK;
}
