/*
 * cq_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A CPU restatement, in plain C on libc, of krow89/cq's query hot path, exposed with the
 * same argument structs as include/cq_gpu.h but under the `cqo_` prefix. It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA path
 * on inputs too large for the real reference (its grouping is O(N*G) and its join O(L*R)).
 * Nothing in cq_b200/ may link, import or call it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py runs this restatement and the
 * unmodified reference (compiled from /root/reference into oracle/_ref/ by oracle/Makefile)
 * on the same SQL over the reference's own fixtures plus a quirk corpus, and
 * tests/golden/ holds the reference's outputs so the comparison also runs where
 * /root/reference is absent.
 *
 * Every function cites the reference code it follows (paths relative to the cq tree).
 * The arithmetic is libc's, as in the reference: strtoll, strtod, sscanf, strcmp, snprintf.
 * The only deliberate difference is algorithmic: groups and join matches are found through
 * hash maps instead of linear scans; results (values, order) are the same.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cq_gpu.h"

#define CQO_EXPORT __attribute__((visibility("default")))

static __thread char g_err[512];
static void set_err(const char* m) { snprintf(g_err, sizeof g_err, "%s", m); }
CQO_EXPORT const char* cqo_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------ */
/* tables                                                                               */
/* ------------------------------------------------------------------------------------ */

typedef struct cqo_table {
    char* data;   /* size bytes + zero padding (the reference's strtoll/strtod read past a
                     field's end up to the next non-numeric byte: src/csv_reader.c:207-210) */
    size_t size;
    cqg_csv_config_t cfg;
    int ncols;
    char** names;
    size_t data_start; /* offset of the first byte after the header LINE (not terminators) */
    int shard_index, shard_count;
} cqo_table_t;

typedef struct {
    const char* p;
    size_t len;
} field_t;

/* parse_line (src/csv_reader.c:278-338): quote-aware field split of [ls, le). Fields are
 * written to `out` up to `cap`; the return value is the number of fields found. */
static int split_line(const cqo_table_t* t, const char* ls, const char* le, field_t* out, int cap) {
    const char* ptr = ls;
    int n = 0;
    const char quote = t->cfg.quote, delim = t->cfg.delimiter;
    while (ptr < le) {
        while (ptr < le && isspace((unsigned char)*ptr) && *ptr != '\n' && *ptr != '\r') ptr++; /* :287 */
        if (ptr >= le) break;                                                                   /* :289 */
        const char* fs = ptr;
        size_t flen = 0;
        if (*ptr == quote) { /* :295 */
            ptr++;
            fs = ptr;
            while (ptr < le) {
                if (*ptr == quote) {
                    if (ptr + 1 < le && *(ptr + 1) == quote) { /* :302 doubled quote: skipped, not un-escaped */
                        ptr += 2;
                        flen += 2;
                    } else {
                        flen = (size_t)(ptr - fs); /* :307 */
                        ptr++;
                        break;
                    }
                } else {
                    ptr++;
                }
            }
            while (ptr < le && *ptr != delim && *ptr != '\n' && *ptr != '\r') ptr++; /* :317 */
        } else {
            while (ptr < le && *ptr != delim && *ptr != '\n' && *ptr != '\r') ptr++; /* :320 */
            flen = (size_t)(ptr - fs);
        }
        if (n < cap) {
            out[n].p = fs;
            out[n].len = flen;
        }
        n++;
        if (ptr < le && *ptr == delim) ptr++; /* :335 */
    }
    return n;
}

/* trim_whitespace (src/csv_reader.c:27-42) on a malloc'ed NUL-terminated copy */
static void trim_ws(char* str) {
    char* s = str;
    while (*s && isspace((unsigned char)*s)) s++;
    char* end = s + strlen(s) - 1;
    while (end > s && isspace((unsigned char)*end)) *end-- = '\0';
    if (s != str) memmove(str, s, strlen(s) + 1);
}

static int parse_header(cqo_table_t* t) {
    const char* ptr = t->data;
    const char* end = t->data + t->size;
    /* csv_load line loop (src/csv_reader.c:404-427): first NON-EMPTY line is the header */
    while (ptr < end) {
        const char* ls = ptr;
        while (ptr < end && *ptr != '\n' && *ptr != '\r') ptr++;
        const char* le = ptr;
        if (le > ls) {
            int cap = 16, n;
            field_t* f = malloc(sizeof(field_t) * cap);
            n = split_line(t, ls, le, f, cap);
            if (n > cap) {
                cap = n;
                f = realloc(f, sizeof(field_t) * cap);
                n = split_line(t, ls, le, f, cap);
            }
            t->ncols = n;
            t->names = calloc(n > 0 ? n : 1, sizeof(char*));
            for (int i = 0; i < n; i++) { /* :346-357 */
                if (t->cfg.has_header && f[i].len > 0) {
                    t->names[i] = strndup(f[i].p, f[i].len);
                    trim_ws(t->names[i]);
                } else {
                    char b[16];
                    snprintf(b, sizeof b, "$%d", i);
                    t->names[i] = strdup(b);
                }
            }
            free(f);
            /* :417 without a header the first line is also data */
            t->data_start = t->cfg.has_header ? (size_t)(le - t->data) : (size_t)(ls - t->data);
            return 0;
        }
        while (ptr < end && (*ptr == '\n' || *ptr == '\r')) ptr++;
    }
    t->ncols = 0;
    t->names = calloc(1, sizeof(char*));
    t->data_start = t->size;
    return 0;
}

CQO_EXPORT int cqo_table_open_buffer(const void* data, size_t size, int pinned, cqg_csv_config_t cfg,
                                     cqg_table_t** out) {
    (void)pinned;
    cqo_table_t* t = calloc(1, sizeof *t);
    t->data = calloc(size + 64, 1);
    memcpy(t->data, data, size);
    t->size = size;
    t->cfg = cfg;
    t->shard_index = 0;
    t->shard_count = 1;
    parse_header(t);
    *out = (cqg_table_t*)t;
    return CQG_OK;
}

/* portable_mmap (src/mmap.c:78-108): open + fstat; an empty file is an error */
CQO_EXPORT int cqo_table_open(const char* path, cqg_csv_config_t cfg, cqg_table_t** out) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_err("Error loading file");
        return CQG_ERR_IO;
    }
    struct stat sb;
    if (fstat(fd, &sb) < 0 || sb.st_size == 0) {
        close(fd);
        set_err("Error loading file");
        return CQG_ERR_IO;
    }
    size_t size = (size_t)sb.st_size;
    char* buf = malloc(size);
    size_t got = 0;
    while (got < size) {
        ssize_t r = read(fd, buf + got, size - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    int rc = cqo_table_open_buffer(buf, got, 0, cfg, out);
    free(buf);
    return rc;
}

CQO_EXPORT int cqo_table_set_shard(cqg_table_t* tt, int index, int count) {
    cqo_table_t* t = (cqo_table_t*)tt;
    if (count < 1 || index < 0 || index >= count) return CQG_ERR_ARG;
    t->shard_index = index;
    t->shard_count = count;
    return CQG_OK;
}

CQO_EXPORT void cqo_table_close(cqg_table_t* tt) {
    cqo_table_t* t = (cqo_table_t*)tt;
    if (!t) return;
    for (int i = 0; i < t->ncols; i++) free(t->names[i]);
    free(t->names);
    free(t->data);
    free(t);
}

CQO_EXPORT int cqo_table_column_count(const cqg_table_t* t) { return ((const cqo_table_t*)t)->ncols; }
CQO_EXPORT const char* cqo_table_column_name(const cqg_table_t* tt, int c) {
    const cqo_table_t* t = (const cqo_table_t*)tt;
    return (c >= 0 && c < t->ncols) ? t->names[c] : NULL;
}
/* csv_get_column_index (src/csv_reader.c:500-509) */
CQO_EXPORT int cqo_table_column_index(const cqg_table_t* tt, const char* name) {
    const cqo_table_t* t = (const cqo_table_t*)tt;
    if (!name) return -1;
    for (int i = 0; i < t->ncols; i++)
        if (strcasecmp(t->names[i], name) == 0) return i;
    return -1;
}
CQO_EXPORT size_t cqo_table_size(const cqg_table_t* t) { return ((const cqo_table_t*)t)->size; }

/* ------------------------------------------------------------------------------------ */
/* values                                                                               */
/* ------------------------------------------------------------------------------------ */

/* is_leap_year / days_in_month / is_valid_date (src/date_utils.c:8-24) */
static int leap(int y) { return (y % 4 == 0 && y % 100 != 0) || (y % 400 == 0); }
static int dim(int y, int m) {
    static const int d[] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    if (m < 1 || m > 12) return 0;
    if (m == 2 && leap(y)) return 29;
    return d[m - 1];
}
static int valid_date(int y, int m, int d) {
    if (y < 1000 || y > 9999) return 0;
    if (m < 1 || m > 12) return 0;
    if (d < 1) return 0;
    return d <= dim(y, m);
}

/* parse_date (src/date_utils.c:26-100): ISO, US, EU, COMPACT tried in that order with the
 * reference's own sscanf formats. */
static int parse_date_str(const char* s, cqg_date_t* out) {
    int y = 0, m = 0, d = 0;
    if (sscanf(s, "%d-%d-%d", &y, &m, &d) == 3 && valid_date(y, m, d)) goto ok;
    y = m = d = 0;
    if (sscanf(s, "%d/%d/%d", &m, &d, &y) == 3 && valid_date(y, m, d)) goto ok;
    y = m = d = 0;
    if (sscanf(s, "%d/%d/%d", &d, &m, &y) == 3 && valid_date(y, m, d)) goto ok;
    y = m = d = 0;
    if (sscanf(s, "%8d", &y) == 1) {
        d = y % 100;
        y /= 100;
        m = y % 100;
        y /= 100;
        if (valid_date(y, m, d)) goto ok;
    }
    return 0;
ok:
    out->year = y;
    out->month = m;
    out->day = d;
    return 1;
}

static int try_date(const char* str, size_t len, cqg_date_t* out) {
    char buf[32];
    memcpy(buf, str, len);
    buf[len] = '\0';
    char* tr = buf; /* src/csv_reader.c:143-149 */
    while (*tr && isspace((unsigned char)*tr)) tr++;
    size_t tl = strlen(tr);
    while (tl > 0 && isspace((unsigned char)tr[tl - 1])) tr[--tl] = '\0';
    return parse_date_str(tr, out);
}

/* infer_type (src/csv_reader.c:133-193) */
static int infer_type(const char* str, size_t len, cqg_date_t* date) {
    if (len == 0) return CQG_TYPE_NULL;
    if (len >= 8 && len <= 10) {
        if (try_date(str, len, date)) return CQG_TYPE_DATE;
    }
    bool has_dot = false, is_number = true, has_digit = false;
    size_t i = 0;
    while (i < len && isspace((unsigned char)str[i])) i++;
    if (i < len && (str[i] == '+' || str[i] == '-')) i++;
    if (i >= len) return CQG_TYPE_STRING;
    while (i < len && !isspace((unsigned char)str[i])) {
        if (isdigit((unsigned char)str[i])) {
            has_digit = true;
        } else if (str[i] == '.' && !has_dot) {
            has_dot = true;
        } else {
            is_number = false;
            break;
        }
        i++;
    }
    while (i < len && isspace((unsigned char)str[i])) i++;
    if (is_number && has_digit && i == len) return has_dot ? CQG_TYPE_DOUBLE : CQG_TYPE_INTEGER;
    return CQG_TYPE_STRING;
}

/* internal value: strings are (pointer,len) views, trimmed as parse_value would */
typedef struct {
    int type;          /* cqg_type_t, or VT_BOOL on the predicate stack */
    long long i;
    double d;
    const char* s;     /* STRING: trimmed view, NOT NUL-terminated */
    size_t slen;
    cqg_date_t date;
} val_t;
#define VT_BOOL 100

/* parse_value (src/csv_reader.c:195-240). strtoll/strtod run on the un-terminated field
 * pointer exactly as the reference does (:207,:210). */
static val_t parse_val(const char* str, size_t len) {
    val_t v;
    memset(&v, 0, sizeof v);
    v.type = infer_type(str, len, &v.date);
    switch (v.type) {
        case CQG_TYPE_INTEGER:
            v.i = strtoll(str, NULL, 10);
            break;
        case CQG_TYPE_DOUBLE:
            v.d = strtod(str, NULL);
            break;
        case CQG_TYPE_STRING: {
            /* cq_strndup + trim_whitespace (:234-235), as a view */
            const char* s = str;
            size_t l = len;
            /* a NUL inside the field ends the reference's C string */
            const char* z = memchr(s, 0, l);
            if (z) l = (size_t)(z - s);
            while (l > 0 && isspace((unsigned char)*s)) {
                s++;
                l--;
            }
            /* trim_whitespace: `while (end > s && isspace(*end))` never removes the first byte */
            while (l > 1 && isspace((unsigned char)s[l - 1])) l--;
            v.s = s;
            v.slen = l;
            break;
        }
        default:
            break;
    }
    return v;
}

/* strcmp on two views (the reference compares NUL-terminated copies: src/csv_reader.c:125) */
static int view_cmp(const char* a, size_t al, const char* b, size_t bl) {
    size_t n = al < bl ? al : bl;
    for (size_t i = 0; i < n; i++) {
        unsigned char ca = (unsigned char)a[i], cb = (unsigned char)b[i];
        if (ca != cb) return (int)ca - (int)cb;
    }
    if (al == bl) return 0;
    return al < bl ? -(int)(unsigned char)b[n] : (int)(unsigned char)a[n];
}

/* value_compare (src/csv_reader.c:98-130) */
static int val_compare(const val_t* a, const val_t* b) {
    if (a->type == CQG_TYPE_NULL && b->type == CQG_TYPE_NULL) return 0;
    if (a->type == CQG_TYPE_NULL) return -1;
    if (b->type == CQG_TYPE_NULL) return 1;
    if (a->type == CQG_TYPE_DATE && b->type == CQG_TYPE_DATE) { /* compare_dates date_utils.c:195-199 */
        if (a->date.year != b->date.year) return a->date.year - b->date.year;
        if (a->date.month != b->date.month) return a->date.month - b->date.month;
        return a->date.day - b->date.day;
    }
    bool an = a->type == CQG_TYPE_INTEGER || a->type == CQG_TYPE_DOUBLE;
    bool bn = b->type == CQG_TYPE_INTEGER || b->type == CQG_TYPE_DOUBLE;
    if (an && bn) {
        double av = a->type == CQG_TYPE_INTEGER ? (double)a->i : a->d;
        double bv = b->type == CQG_TYPE_INTEGER ? (double)b->i : b->d;
        if (av < bv) return -1;
        if (av > bv) return 1;
        return 0;
    }
    if (a->type == CQG_TYPE_STRING && b->type == CQG_TYPE_STRING) return view_cmp(a->s, a->slen, b->s, b->slen);
    return 0;
}

static val_t val_from_const(const cqg_value_t* c) {
    val_t v;
    memset(&v, 0, sizeof v);
    v.type = c->type;
    switch (c->type) {
        case CQG_TYPE_INTEGER: v.i = c->int_value; break;
        case CQG_TYPE_DOUBLE: v.d = c->double_value; break;
        case CQG_TYPE_STRING:
            v.s = c->string_value ? c->string_value : "";
            v.slen = strlen(v.s);
            break;
        case CQG_TYPE_DATE: v.date = c->date_value; break;
        default: break;
    }
    return v;
}

/* match_pattern (src/evaluator/evaluator_conditions.c:16-59) on views */
static bool like_match(const char* str, size_t sl, const char* pat, size_t pl, bool cs) {
    size_t s = 0, p = 0, star = (size_t)-1, ss = 0;
    while (s < sl) {
        char pc = p < pl ? pat[p] : '\0';
        if (pc == '%') {
            star = p++;
            ss = s;
        } else if (pc == '_') {
            s++;
            p++;
        } else {
            bool m = cs ? (str[s] == pc) : (tolower((unsigned char)str[s]) == tolower((unsigned char)pc));
            if (m) {
                s++;
                p++;
            } else if (star != (size_t)-1) {
                p = star + 1;
                s = ++ss;
            } else {
                return false;
            }
        }
    }
    while (p < pl && pat[p] == '%') p++;
    return p == pl;
}

/* ------------------------------------------------------------------------------------ */
/* row access                                                                           */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    const cqo_table_t* t;
    field_t* f;
    int nf, cap;
} rowview_t;

static void row_parse(rowview_t* r, const char* ls, const char* le) {
    int n = split_line(r->t, ls, le, r->f, r->cap);
    if (n > r->cap) {
        r->cap = n * 2;
        r->f = realloc(r->f, sizeof(field_t) * r->cap);
        n = split_line(r->t, ls, le, r->f, r->cap);
    }
    r->nf = n;
}

/* ragged rows: a missing column is NULL (what build_result does,
 * evaluator_utils.c:240) */
static val_t row_col(const rowview_t* r, int col) {
    if (col < 0 || col >= r->nf) {
        val_t v;
        memset(&v, 0, sizeof v);
        return v;
    }
    return parse_val(r->f[col].p, r->f[col].len);
}

/* joined row: left columns then right columns (evaluator_joins.c:73-77) */
typedef struct {
    const rowview_t* l;
    const rowview_t* r; /* NULL when no join */
    int nleft;
} jrow_t;

static val_t jrow_col(const jrow_t* j, int col) {
    if (col < 0) {
        val_t v;
        memset(&v, 0, sizeof v);
        return v;
    }
    if (!j->r || col < j->nleft) return row_col(j->l, col);
    return row_col(j->r, col - j->nleft);
}

/* ------------------------------------------------------------------------------------ */
/* predicate interpreter                                                                */
/* ------------------------------------------------------------------------------------ */

/* the x86-64 result of `(long long)double` for out-of-range / NaN inputs (cvttsd2si) */
static long long d2ll_x86(double x) {
    if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) {
        if (x == -9223372036854775808.0) return (long long)x;
        return (long long)0x8000000000000000ULL;
    }
    return (long long)x;
}

/* BINARY_OP arm of evaluate_expression (evaluator_expressions.c:156-262) */
static val_t arith(int op, const val_t* l, const val_t* r) {
    val_t res;
    memset(&res, 0, sizeof res);
    double lv, rv;
    long long li = 0, ri = 0;
    bool lint = false, rint = false;
    if (l->type == CQG_TYPE_INTEGER) {
        lv = (double)l->i;
        li = l->i;
        lint = true;
    } else if (l->type == CQG_TYPE_DOUBLE) {
        lv = l->d;
    } else {
        return res;
    }
    if (r->type == CQG_TYPE_INTEGER) {
        rv = (double)r->i;
        ri = r->i;
        rint = true;
    } else if (r->type == CQG_TYPE_DOUBLE) {
        rv = r->d;
    } else {
        return res;
    }
    double out = 0;
    long long outi = 0;
    bool is_int = false;
    switch (op) {
        case CQG_OP_ADD: out = lv + rv; break;
        case CQG_OP_SUB: out = lv - rv; break;
        case CQG_OP_MUL: out = lv * rv; break;
        case CQG_OP_DIV:
            if (rv == 0) return res;
            out = lv / rv;
            break;
        case CQG_OP_MOD:
            if (lint && rint) {
                if (ri == 0) return res;
                /* LLONG_MIN % -1 traps on x86; the reference would crash. Defined here as 0. */
                outi = (ri == -1) ? 0 : li % ri;
                is_int = true;
            } else {
                if (rv == 0) return res;
                out = fmod(lv, rv);
            }
            break;
        case CQG_OP_BAND:
            if (!(lint && rint)) return res;
            outi = li & ri;
            is_int = true;
            break;
        case CQG_OP_BOR:
            if (!(lint && rint)) return res;
            outi = li | ri;
            is_int = true;
            break;
        case CQG_OP_BXOR:
            if (!(lint && rint)) return res;
            outi = li ^ ri;
            is_int = true;
            break;
        default: break; /* CQG_OP_ARITH_NULL: out stays 0 */
    }
    if (is_int) {
        res.type = CQG_TYPE_INTEGER;
        res.i = outi;
    } else if (lint && rint && out == (double)d2ll_x86(out)) { /* :253-256 */
        res.type = CQG_TYPE_INTEGER;
        res.i = d2ll_x86(out);
    } else {
        res.type = CQG_TYPE_DOUBLE;
        res.d = out;
    }
    return res;
}

#define STACK_MAX 64

/* evaluate_condition (evaluator_conditions.c:62-164) over postfix code */
static bool eval_pred(const cqg_predicate_t* p, const jrow_t* row) {
    if (!p->code || p->n_code == 0) return true;
    val_t st[STACK_MAX];
    int sp = 0;
    for (int pc = 0; pc < p->n_code; pc++) {
        int op = p->code[pc].op, a = p->code[pc].a;
        switch (op) {
            case CQG_OP_COL: st[sp++] = jrow_col(row, a); break;
            case CQG_OP_CONST: st[sp++] = val_from_const(&p->consts[a]); break;
            case CQG_OP_ADD: case CQG_OP_SUB: case CQG_OP_MUL: case CQG_OP_DIV: case CQG_OP_MOD:
            case CQG_OP_BAND: case CQG_OP_BOR: case CQG_OP_BXOR: case CQG_OP_ARITH_NULL: {
                val_t r = st[--sp], l = st[--sp];
                st[sp++] = arith(op, &l, &r);
                break;
            }
            case CQG_OP_NEG: { /* expressions.c:112-128 */
                val_t o = st[--sp], r;
                memset(&r, 0, sizeof r);
                if (o.type == CQG_TYPE_INTEGER) {
                    r.type = CQG_TYPE_INTEGER;
                    r.i = (long long)(0ULL - (unsigned long long)o.i);
                } else if (o.type == CQG_TYPE_DOUBLE) {
                    r.type = CQG_TYPE_DOUBLE;
                    r.d = -o.d;
                }
                st[sp++] = r;
                break;
            }
            case CQG_OP_POS: break;
            case CQG_OP_EQ: case CQG_OP_NE: case CQG_OP_GT: case CQG_OP_LT: case CQG_OP_GE: case CQG_OP_LE: {
                val_t r = st[--sp], l = st[--sp];
                int c = val_compare(&l, &r);
                bool b = op == CQG_OP_EQ ? c == 0 : op == CQG_OP_NE ? c != 0 : op == CQG_OP_GT ? c > 0
                       : op == CQG_OP_LT ? c < 0 : op == CQG_OP_GE ? c >= 0 : c <= 0;
                val_t o;
                memset(&o, 0, sizeof o);
                o.type = VT_BOOL;
                o.i = b;
                st[sp++] = o;
                break;
            }
            case CQG_OP_IN: case CQG_OP_NOT_IN: { /* conditions.c:134-148 */
                bool found = false;
                val_t* left = &st[sp - a - 1];
                for (int k = 0; k < a; k++)
                    if (val_compare(left, &st[sp - a + k]) == 0) {
                        found = true;
                        break;
                    }
                sp -= a + 1;
                val_t o;
                memset(&o, 0, sizeof o);
                o.type = VT_BOOL;
                o.i = (op == CQG_OP_IN) ? found : !found;
                st[sp++] = o;
                break;
            }
            case CQG_OP_LIKE: case CQG_OP_ILIKE: { /* conditions.c:152-161 */
                val_t r = st[--sp], l = st[--sp];
                val_t o;
                memset(&o, 0, sizeof o);
                o.type = VT_BOOL;
                o.i = (l.type == CQG_TYPE_STRING && r.type == CQG_TYPE_STRING)
                          ? like_match(l.s, l.slen, r.s, r.slen, op == CQG_OP_LIKE)
                          : 0;
                st[sp++] = o;
                break;
            }
            case CQG_OP_AND: {
                bool r = st[--sp].i, l = st[--sp].i;
                st[sp].type = VT_BOOL;
                st[sp++].i = l && r;
                break;
            }
            case CQG_OP_OR: {
                bool r = st[--sp].i, l = st[--sp].i;
                st[sp].type = VT_BOOL;
                st[sp++].i = l || r;
                break;
            }
            case CQG_OP_NOT: st[sp - 1].i = !st[sp - 1].i; break;
            case CQG_OP_TRUE: st[sp].type = VT_BOOL; st[sp++].i = 1; break;
            case CQG_OP_FALSE: st[sp].type = VT_BOOL; st[sp++].i = 0; break;
            case CQG_OP_POP: sp--; break;
            default: return false;
        }
    }
    return sp > 0 && st[sp - 1].i != 0;
}

/* ------------------------------------------------------------------------------------ */
/* grouping                                                                             */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    bool have;
    val_t v; /* strings: view into the table data */
} ext_t;

typedef struct {
    char* key;
    uint64_t first_off, first_off_right;
    long long count;
    double* sum;
    long long* ncount;
    ext_t* mn;
    ext_t* mx;
} group_t;

typedef struct {
    group_t* g;
    long long n, cap;
    long long* slots; /* open addressing: index+1 */
    long long nslots;
} gmap_t;

static uint64_t fnv1a(const char* s) {
    uint64_t h = 1469598103934665603ULL;
    for (; *s; s++) {
        h ^= (unsigned char)*s;
        h *= 1099511628211ULL;
    }
    return h;
}

static void gmap_init(gmap_t* m) {
    memset(m, 0, sizeof *m);
    m->cap = 64;
    m->g = calloc((size_t)m->cap, sizeof(group_t));
    m->nslots = 256;
    m->slots = calloc((size_t)m->nslots, sizeof(long long));
}

static void gmap_rehash(gmap_t* m) {
    free(m->slots);
    m->nslots *= 4;
    m->slots = calloc((size_t)m->nslots, sizeof(long long));
    for (long long i = 0; i < m->n; i++) {
        uint64_t h = fnv1a(m->g[i].key) & (uint64_t)(m->nslots - 1);
        while (m->slots[h]) h = (h + 1) & (uint64_t)(m->nslots - 1);
        m->slots[h] = i + 1;
    }
}

/* find-or-create by rendered key: same identity as the strcmp scan at
 * evaluator_aggregates.c:145-150 / src/evaluator.c:183-188 */
static group_t* gmap_get(gmap_t* m, const char* key, int naggs, bool* created) {
    uint64_t h = fnv1a(key) & (uint64_t)(m->nslots - 1);
    while (m->slots[h]) {
        group_t* g = &m->g[m->slots[h] - 1];
        if (strcmp(g->key, key) == 0) {
            *created = false;
            return g;
        }
        h = (h + 1) & (uint64_t)(m->nslots - 1);
    }
    if (m->n == m->cap) {
        m->cap *= 2;
        m->g = realloc(m->g, sizeof(group_t) * (size_t)m->cap);
    }
    group_t* g = &m->g[m->n];
    memset(g, 0, sizeof *g);
    g->key = strdup(key);
    g->sum = calloc((size_t)(naggs > 0 ? naggs : 1), sizeof(double));
    g->ncount = calloc((size_t)(naggs > 0 ? naggs : 1), sizeof(long long));
    g->mn = calloc((size_t)(naggs > 0 ? naggs : 1), sizeof(ext_t));
    g->mx = calloc((size_t)(naggs > 0 ? naggs : 1), sizeof(ext_t));
    m->slots[h] = ++m->n;
    if (m->n * 2 > m->nslots) {
        gmap_rehash(m);
        g = &m->g[m->n - 1];
    }
    *created = true;
    return g;
}

static void gmap_free(gmap_t* m) {
    for (long long i = 0; i < m->n; i++) {
        free(m->g[i].key);
        free(m->g[i].sum);
        free(m->g[i].ncount);
        free(m->g[i].mn);
        free(m->g[i].mx);
    }
    free(m->g);
    free(m->slots);
}

/* key rendering (evaluator_aggregates.c:121-141; src/evaluator.c:155-173) */
static void render_key_part(const val_t* v, char* out /* 256 */) {
    switch (v->type) {
        case CQG_TYPE_NULL: strcpy(out, "NULL"); break;
        case CQG_TYPE_INTEGER: snprintf(out, 256, "%lld", v->i); break;
        case CQG_TYPE_DOUBLE: snprintf(out, 256, "%.6f", v->d); break;
        case CQG_TYPE_DATE:
            snprintf(out, 256, "%04d-%02d-%02d", v->date.year, v->date.month, v->date.day);
            break;
        case CQG_TYPE_STRING: {
            size_t n = v->slen < 255 ? v->slen : 255; /* strncpy(..., 255) */
            memcpy(out, v->s, n);
            out[n] = '\0';
            break;
        }
        default: out[0] = '\0';
    }
}

/* ------------------------------------------------------------------------------------ */
/* result building                                                                      */
/* ------------------------------------------------------------------------------------ */

typedef struct arena_blk {
    struct arena_blk* next;
} arena_blk_t;

static void* arena_alloc(cqg_result_t* r, size_t n) {
    arena_blk_t* b = calloc(1, sizeof(arena_blk_t) + n + 8);
    b->next = (arena_blk_t*)r->arena;
    r->arena = b;
    return (void*)(b + 1);
}

CQO_EXPORT void cqo_result_free(cqg_result_t* r) {
    if (!r) return;
    arena_blk_t* b = (arena_blk_t*)r->arena;
    while (b) {
        arena_blk_t* n = b->next;
        free(b);
        b = n;
    }
    free(r);
}

static cqg_value_t export_val(cqg_result_t* r, const val_t* v) {
    cqg_value_t o;
    memset(&o, 0, sizeof o);
    o.type = v->type;
    switch (v->type) {
        case CQG_TYPE_INTEGER: o.int_value = v->i; break;
        case CQG_TYPE_DOUBLE: o.double_value = v->d; break;
        case CQG_TYPE_DATE: o.date_value = v->date; break;
        case CQG_TYPE_STRING: {
            char* s = arena_alloc(r, v->slen + 1);
            memcpy(s, v->s, v->slen);
            s[v->slen] = '\0';
            o.string_value = s;
            break;
        }
        default: break;
    }
    return o;
}

/* one filtered (possibly joined) row enters the aggregation: create_groups +
 * evaluate_aggregate streamed in row order (evaluator_aggregates.c:118-173, 286-326) */
static void agg_row(gmap_t* m, const cqg_query_t* q, const jrow_t* jr, uint64_t loff, uint64_t roff) {
    char key[2048];
    if (q->n_group_cols == 0) {
        strcpy(key, "_all_");
    } else {
        key[0] = '\0';
        size_t kl = 0;
        for (int g = 0; g < q->n_group_cols; g++) {
            char part[256];
            val_t v = jrow_col(jr, q->group_cols[g]);
            render_key_part(&v, part);
            if (g > 0 && kl + 1 < sizeof key) key[kl++] = '\t'; /* src/evaluator.c:124 */
            size_t pl = strlen(part);
            if (kl + pl >= sizeof key) pl = sizeof key - 1 - kl;
            memcpy(key + kl, part, pl);
            kl += pl;
            key[kl] = '\0';
        }
    }
    bool created;
    group_t* g = gmap_get(m, key, q->n_aggs, &created);
    if (created) {
        g->first_off = loff;
        g->first_off_right = roff;
    }
    g->count++;
    for (int a = 0; a < q->n_aggs; a++) {
        int f = q->aggs[a].func;
        if (f == CQG_AGG_COUNT_STAR || f == CQG_AGG_COUNT || q->aggs[a].col < 0) continue;
        val_t v = jrow_col(jr, q->aggs[a].col);
        if (f == CQG_AGG_SUM || f == CQG_AGG_AVG) {
            if (v.type == CQG_TYPE_INTEGER) {
                g->sum[a] += v.i; /* :293 */
                g->ncount[a]++;
            } else if (v.type == CQG_TYPE_DOUBLE) {
                g->sum[a] += v.d;
                g->ncount[a]++;
            }
        } else if (v.type != CQG_TYPE_NULL) { /* :316-321 first value wins ties */
            if (f == CQG_AGG_MIN) {
                if (!g->mn[a].have || val_compare(&v, &g->mn[a].v) < 0) {
                    g->mn[a].have = true;
                    g->mn[a].v = v;
                }
            } else {
                if (!g->mx[a].have || val_compare(&v, &g->mx[a].v) > 0) {
                    g->mx[a].have = true;
                    g->mx[a].v = v;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* scan driver                                                                          */
/* ------------------------------------------------------------------------------------ */

typedef void (*row_cb)(void* ctx, const char* ls, const char* le);

/* csv_load's line loop (src/csv_reader.c:404-427) over the data rows of this shard:
 * rows whose first byte lies in [lo, hi) of the shard's byte range. */
static void for_each_row(const cqo_table_t* t, row_cb cb, void* ctx) {
    const char* base = t->data;
    const char* end = base + t->size;
    size_t lo = (size_t)((__uint128_t)t->size * (unsigned)t->shard_index / (unsigned)t->shard_count);
    size_t hi = (size_t)((__uint128_t)t->size * (unsigned)(t->shard_index + 1) / (unsigned)t->shard_count);
    const char* ptr = base + t->data_start;
    while (ptr < end) {
        const char* ls = ptr;
        while (ptr < end && *ptr != '\n' && *ptr != '\r') ptr++;
        const char* le = ptr;
        if (le > ls) {
            size_t off = (size_t)(ls - base);
            if (off >= hi) return;
            if (off >= lo) cb(ctx, ls, le);
        }
        while (ptr < end && (*ptr == '\n' || *ptr == '\r')) ptr++;
    }
}

/* join index over the right table: rows bucketed by comparison class */
typedef struct {
    uint64_t off;  /* row start offset */
    uint64_t len;  /* row length */
    val_t key;
} rrow_t;

typedef struct {
    rrow_t* rows;
    long long n, cap;
} rlist_t;

typedef struct {
    const cqo_table_t* lt;
    const cqo_table_t* rt;
    const cqg_query_t* q;
    gmap_t gm;
    rowview_t lrow, rrow;
    /* select */
    cqg_result_t* res;
    long long n_selected;
    uint64_t* sel_off;
    uint64_t* sel_off_r;
    long long sel_cap;
    long long rows_scanned;
    /* join */
    rlist_t right;
    long long* bucket_head; /* hash on class+value -> chain through next[] */
    long long* next;
    long long nbuckets;
    long long class_count[5];
    unsigned char* matched; /* RIGHT / FULL joins: right row i matched some left row (evaluator_joins.c:146-157) */
} run_t;

/* offsets that stand for "no row on this side" in outer joins (include/cq_gpu.h, cqg_result_t::row_offset) */
#define CQO_RIGHT_ONLY (1ULL << 45) /* | right offset: a right row without a match has no left row */
#define CQO_NO_RIGHT (~0ULL)        /* a left row without a match */

/* parse the row at `off`, or make the view an empty row (every column NULL) when that side has none */
static void parse_side(rowview_t* rv, const cqo_table_t* t, uint64_t off, bool none);

static int key_class(const val_t* v) {
    switch (v->type) {
        case CQG_TYPE_INTEGER: case CQG_TYPE_DOUBLE: return 1;
        case CQG_TYPE_STRING: return 2;
        case CQG_TYPE_DATE: return 3;
        default: return 0;
    }
}

static uint64_t key_hash(const val_t* v) {
    switch (key_class(v)) {
        case 1: {
            double d = v->type == CQG_TYPE_INTEGER ? (double)v->i : v->d;
            if (d == 0) d = 0; /* -0.0 == 0.0 */
            uint64_t b;
            memcpy(&b, &d, 8);
            b ^= b >> 33;
            b *= 0xff51afd7ed558ccdULL;
            b ^= b >> 33;
            return b ^ 0x1111;
        }
        case 2: {
            uint64_t h = 1469598103934665603ULL;
            for (size_t i = 0; i < v->slen; i++) {
                h ^= (unsigned char)v->s[i];
                h *= 1099511628211ULL;
            }
            return h ^ 0x2222;
        }
        case 3: return ((uint64_t)v->date.year * 512 + (uint64_t)v->date.month * 32 + (uint64_t)v->date.day) * 0x9E3779B97F4A7C15ULL;
        default: return 0x4444;
    }
}

static void emit_row(run_t* R, const jrow_t* jr, uint64_t loff, uint64_t roff) {
    if (!eval_pred(&R->q->where, jr)) return;
    if (R->q->mode == CQG_MODE_AGGREGATE) {
        agg_row(&R->gm, R->q, jr, loff, roff);
    } else {
        if (R->n_selected == R->sel_cap) {
            R->sel_cap = R->sel_cap ? R->sel_cap * 2 : 1024;
            R->sel_off = realloc(R->sel_off, sizeof(uint64_t) * (size_t)R->sel_cap);
            R->sel_off_r = realloc(R->sel_off_r, sizeof(uint64_t) * (size_t)R->sel_cap);
        }
        R->sel_off[R->n_selected] = loff;
        R->sel_off_r[R->n_selected] = roff;
        R->n_selected++;
    }
}

static void right_row_cb(void* ctx, const char* ls, const char* le) {
    run_t* R = ctx;
    row_parse(&R->rrow, ls, le);
    if (R->right.n == R->right.cap) {
        R->right.cap = R->right.cap ? R->right.cap * 2 : 1024;
        R->right.rows = realloc(R->right.rows, sizeof(rrow_t) * (size_t)R->right.cap);
    }
    rrow_t* rr = &R->right.rows[R->right.n++];
    rr->off = (uint64_t)(ls - R->rt->data);
    rr->len = (uint64_t)(le - ls);
    rr->key = row_col(&R->rrow, R->q->join.right_col);
    R->class_count[key_class(&rr->key)]++;
}

static const char* row_end(const cqo_table_t* t, const char* ls);
static void parse_side(rowview_t* rv, const cqo_table_t* t, uint64_t off, bool none) {
    if (none) {
        rv->nf = 0;
        return;
    }
    const char* ls = t->data + off;
    row_parse(rv, ls, row_end(t, ls));
}

static int cmp_ll(const void* a, const void* b) {
    long long x = *(const long long*)a, y = *(const long long*)b;
    return x < y ? -1 : x > y;
}

static void left_row_cb(void* ctx, const char* ls, const char* le) {
    run_t* R = ctx;
    R->rows_scanned++;
    row_parse(&R->lrow, ls, le);
    uint64_t loff = (uint64_t)(ls - R->lt->data);
    if (!R->rt) {
        jrow_t jr = {&R->lrow, NULL, R->lt->ncols};
        emit_row(R, &jr, loff, 0);
        return;
    }
    /* perform_join's inner loop (evaluator_joins.c:96-126): all right rows r, in file
     * order, with value_compare(left_key, right_key) == 0 */
    const int jtype = R->q->join.type;
    const bool keep_left = jtype == CQG_JOIN_LEFT || jtype == CQG_JOIN_FULL;
    if (R->q->join.left_col < 0 || R->q->join.right_col < 0) { /* resolve_column NULL -> false for every pair */
        if (keep_left) { /* :128-139: the left row with NULL right columns */
            R->rrow.nf = 0;
            jrow_t jr = {&R->lrow, &R->rrow, R->lt->ncols};
            emit_row(R, &jr, loff, CQO_NO_RIGHT);
        }
        return;
    }
    val_t lk = row_col(&R->lrow, R->q->join.left_col);
    int lc = key_class(&lk);
    long long* matches = NULL;
    long long nm = 0, mcap = 0;
    bool cross = false;
    if (lc != 0)
        for (int c = 1; c <= 3; c++)
            if (c != lc && R->class_count[c] > 0) cross = true;
    if (cross) {
        /* value_compare returns 0 for every mixed non-NULL class pair: rare, scan */
        for (long long i = 0; i < R->right.n; i++) {
            if (val_compare(&lk, &R->right.rows[i].key) == 0) {
                if (nm == mcap) {
                    mcap = mcap ? mcap * 2 : 16;
                    matches = realloc(matches, sizeof(long long) * (size_t)mcap);
                }
                matches[nm++] = i;
            }
        }
    } else {
        uint64_t h = key_hash(&lk) & (uint64_t)(R->nbuckets - 1);
        for (long long i = R->bucket_head[h]; i >= 0; i = R->next[i]) {
            if (val_compare(&lk, &R->right.rows[i].key) == 0) {
                if (nm == mcap) {
                    mcap = mcap ? mcap * 2 : 16;
                    matches = realloc(matches, sizeof(long long) * (size_t)mcap);
                }
                matches[nm++] = i;
            }
        }
        qsort(matches, (size_t)nm, sizeof(long long), cmp_ll);
    }
    for (long long k = 0; k < nm; k++) {
        rrow_t* rr = &R->right.rows[matches[k]];
        if (R->matched) R->matched[matches[k]] = 1;
        row_parse(&R->rrow, R->rt->data + rr->off, R->rt->data + rr->off + rr->len);
        jrow_t jr = {&R->lrow, &R->rrow, R->lt->ncols};
        emit_row(R, &jr, loff, rr->off);
    }
    if (nm == 0 && keep_left) { /* :128-139 */
        R->rrow.nf = 0;
        jrow_t jr = {&R->lrow, &R->rrow, R->lt->ncols};
        emit_row(R, &jr, loff, CQO_NO_RIGHT);
    }
    free(matches);
}

static int cmp_group_first(const void* a, const void* b) {
    const group_t* x = a;
    const group_t* y = b;
    if (x->first_off != y->first_off) return x->first_off < y->first_off ? -1 : 1;
    if (x->first_off_right != y->first_off_right) return x->first_off_right < y->first_off_right ? -1 : 1;
    return 0;
}

static const char* row_end(const cqo_table_t* t, const char* ls) {
    const char* end = t->data + t->size;
    const char* p = ls;
    while (p < end && *p != '\n' && *p != '\r') p++;
    return p;
}

CQO_EXPORT int cqo_execute(const cqg_table_t* tt, const cqg_query_t* q, cqg_result_t** out) {
    const cqo_table_t* t = (const cqo_table_t*)tt;
    run_t R;
    memset(&R, 0, sizeof R);
    R.lt = t;
    R.rt = (const cqo_table_t*)q->join.right;
    R.q = q;
    R.lrow.t = t;
    R.lrow.cap = 32;
    R.lrow.f = malloc(sizeof(field_t) * 32);
    R.rrow.t = R.rt;
    R.rrow.cap = 32;
    R.rrow.f = malloc(sizeof(field_t) * 32);
    gmap_init(&R.gm);
    cqg_result_t* res = calloc(1, sizeof *res);
    R.res = res;

    if (R.rt && (q->join.type < CQG_JOIN_INNER || q->join.type > CQG_JOIN_FULL)) {
        free(R.lrow.f);
        free(R.rrow.f);
        free(res);
        set_err("unknown join type");
        return CQG_ERR_ARG;
    }
    if (R.rt && q->join.type >= CQG_JOIN_RIGHT && (q->join.left_col < 0 || q->join.right_col < 0 || t->shard_count > 1)) {
        /* the GPU library declines these too (empty join table / right rows are unmatched only over ALL left rows) */
        free(R.lrow.f);
        free(R.rrow.f);
        free(res);
        set_err("RIGHT / FULL JOIN on an unknown key column or on a shard");
        return CQG_ERR_UNSUPPORTED;
    }
    if (R.rt) {
        /* the right table is always read whole, whatever the left shard */
        cqo_table_t rt_all = *R.rt;
        rt_all.shard_index = 0;
        rt_all.shard_count = 1;
        for_each_row(&rt_all, right_row_cb, &R);
        R.nbuckets = 1024;
        while (R.nbuckets < R.right.n * 2) R.nbuckets *= 2;
        R.bucket_head = malloc(sizeof(long long) * (size_t)R.nbuckets);
        for (long long i = 0; i < R.nbuckets; i++) R.bucket_head[i] = -1;
        R.next = malloc(sizeof(long long) * (size_t)(R.right.n + 1));
        for (long long i = R.right.n - 1; i >= 0; i--) { /* reverse: chains come out ascending-ish */
            uint64_t h = key_hash(&R.right.rows[i].key) & (uint64_t)(R.nbuckets - 1);
            R.next[i] = R.bucket_head[h];
            R.bucket_head[h] = i;
        }
    }

    if (R.rt && q->join.type >= CQG_JOIN_RIGHT) R.matched = calloc((size_t)(R.right.n + 1), 1);
    for_each_row(t, left_row_cb, &R);
    if (R.matched) {
        /* evaluator_joins.c:142-171: behind everything else, the right rows no left row matched, in file order,
         * with NULL left columns */
        for (long long i = 0; i < R.right.n; i++) {
            if (R.matched[i]) continue;
            rrow_t* rr = &R.right.rows[i];
            row_parse(&R.rrow, R.rt->data + rr->off, R.rt->data + rr->off + rr->len);
            R.lrow.nf = 0;
            jrow_t jr = {&R.lrow, &R.rrow, R.lt->ncols};
            emit_row(&R, &jr, CQO_RIGHT_ONLY | rr->off, rr->off);
        }
        free(R.matched);
    }

    res->rows_scanned = R.rows_scanned;
    res->n_aggs = q->n_aggs;
    res->n_out_cols = q->n_out_cols;

    if (q->mode == CQG_MODE_AGGREGATE) {
        bool zero_groups = false;
        /* create_groups returns no group at all when its single key column is unknown
         * (evaluator_aggregates.c:114-116) */
        if (q->n_group_cols == 1 && q->group_cols[0] < 0) zero_groups = true;
        /* no GROUP BY: one `_all_` group even over zero rows (src/evaluator.c:232-247) */
        if (q->n_group_cols == 0 && R.gm.n == 0) {
            bool created;
            gmap_get(&R.gm, "_all_", q->n_aggs, &created);
        }
        long long G = zero_groups ? 0 : R.gm.n;
        qsort(R.gm.g, (size_t)R.gm.n, sizeof(group_t), cmp_group_first);
        res->n_groups = G;
        size_t gn = (size_t)(G > 0 ? G : 1), an = (size_t)(q->n_aggs > 0 ? q->n_aggs : 1),
               on = (size_t)(q->n_out_cols > 0 ? q->n_out_cols : 1);
        res->first_offset = arena_alloc(res, sizeof(uint64_t) * gn);
        res->count = arena_alloc(res, sizeof(int64_t) * gn);
        res->sum = arena_alloc(res, sizeof(double) * gn * an);
        res->ncount = arena_alloc(res, sizeof(int64_t) * gn * an);
        res->value = arena_alloc(res, sizeof(cqg_value_t) * gn * an);
        res->out = arena_alloc(res, sizeof(cqg_value_t) * gn * on);
        for (long long gi = 0; gi < G; gi++) {
            group_t* g = &R.gm.g[gi];
            res->first_offset[gi] = g->first_off;
            res->count[gi] = g->count;
            for (int a = 0; a < q->n_aggs; a++) {
                size_t ix = (size_t)a * (size_t)G + (size_t)gi;
                res->sum[ix] = g->sum[a];
                res->ncount[ix] = g->ncount[a];
                cqg_value_t v;
                memset(&v, 0, sizeof v);
                int f = q->aggs[a].func;
                if (f == CQG_AGG_COUNT_STAR) { /* :268-272 */
                    v.type = CQG_TYPE_INTEGER;
                    v.int_value = g->count;
                } else if (q->aggs[a].col < 0) { /* :276-278 */
                    v.type = CQG_TYPE_NULL;
                } else if (f == CQG_AGG_COUNT) {
                    v.type = CQG_TYPE_INTEGER;
                    v.int_value = g->count;
                } else if (f == CQG_AGG_SUM) {
                    v.type = CQG_TYPE_DOUBLE;
                    v.double_value = g->sum[a];
                } else if (f == CQG_AGG_AVG) { /* :306 */
                    v.type = CQG_TYPE_DOUBLE;
                    v.double_value = g->ncount[a] > 0 ? g->sum[a] / (double)g->ncount[a] : 0;
                } else if (f == CQG_AGG_MIN) {
                    if (g->mn[a].have) v = export_val(res, &g->mn[a].v);
                } else if (f == CQG_AGG_MAX) {
                    if (g->mx[a].have) v = export_val(res, &g->mx[a].v);
                }
                res->value[ix] = v;
            }
            /* bare columns: the group's first row (evaluator_aggregates.c:679-689) */
            if (q->n_out_cols > 0) {
                if (g->count > 0) {
                    parse_side(&R.lrow, t, g->first_off, (g->first_off & CQO_RIGHT_ONLY) != 0);
                    if (R.rt) parse_side(&R.rrow, R.rt, g->first_off_right, g->first_off_right == CQO_NO_RIGHT);
                    jrow_t jr = {&R.lrow, R.rt ? &R.rrow : NULL, t->ncols};
                    for (int c = 0; c < q->n_out_cols; c++) {
                        val_t v = jrow_col(&jr, q->out_cols[c]);
                        res->out[(size_t)c * (size_t)G + (size_t)gi] = export_val(res, &v);
                    }
                } /* else: NULLs (calloc) */
            }
        }
    } else {
        res->n_selected = R.n_selected;
        long long nout = R.n_selected;
        if (q->max_rows >= 0 && nout > q->max_rows) nout = q->max_rows;
        res->n_rows_out = nout;
        size_t rn = (size_t)(nout > 0 ? nout : 1), on = (size_t)(q->n_out_cols > 0 ? q->n_out_cols : 1);
        res->row_offset = arena_alloc(res, sizeof(uint64_t) * rn);
        res->row_offset_right = R.rt ? arena_alloc(res, sizeof(uint64_t) * rn) : NULL;
        res->rows = arena_alloc(res, sizeof(cqg_value_t) * rn * on);
        for (long long i = 0; i < nout; i++) {
            res->row_offset[i] = R.sel_off[i];
            if (R.rt) res->row_offset_right[i] = R.sel_off_r[i];
            parse_side(&R.lrow, t, R.sel_off[i], (R.sel_off[i] & CQO_RIGHT_ONLY) != 0);
            if (R.rt) parse_side(&R.rrow, R.rt, R.sel_off_r[i], R.sel_off_r[i] == CQO_NO_RIGHT);
            jrow_t jr = {&R.lrow, R.rt ? &R.rrow : NULL, t->ncols};
            for (int c = 0; c < q->n_out_cols; c++) {
                val_t v = jrow_col(&jr, q->out_cols[c]);
                res->rows[(size_t)i * (size_t)q->n_out_cols + (size_t)c] = export_val(res, &v);
            }
        }
    }

    gmap_free(&R.gm);
    free(R.lrow.f);
    free(R.rrow.f);
    free(R.sel_off);
    free(R.sel_off_r);
    free(R.right.rows);
    free(R.bucket_head);
    free(R.next);
    *out = res;
    return CQG_OK;
}

static void count_cb(void* ctx, const char* ls, const char* le) {
    (void)ls;
    (void)le;
    (*(int64_t*)ctx)++;
}

CQO_EXPORT int cqo_table_row_count(const cqg_table_t* t, int64_t* out) {
    *out = 0;
    for_each_row((const cqo_table_t*)t, count_cb, out);
    return CQG_OK;
}

CQO_EXPORT int cqo_parse_value(const char* str, size_t len, cqg_value_t* out) {
    /* the reference's literal strings are NUL-terminated (evaluator_expressions.c:31) */
    char* buf = calloc(len + 16, 1);
    memcpy(buf, str, len);
    val_t v = parse_val(buf, len);
    memset(out, 0, sizeof *out);
    out->type = v.type;
    switch (v.type) {
        case CQG_TYPE_INTEGER: out->int_value = v.i; break;
        case CQG_TYPE_DOUBLE: out->double_value = v.d; break;
        case CQG_TYPE_DATE: out->date_value = v.date; break;
        case CQG_TYPE_STRING: out->string_value = strndup(v.s, v.slen); break;
        default: break;
    }
    free(buf);
    return CQG_OK;
}

CQO_EXPORT void cqo_value_release(cqg_value_t* v) {
    if (v && v->type == CQG_TYPE_STRING) {
        free(v->string_value);
        v->string_value = NULL;
    }
}

/* ------------------------------------------------------------------------------------ */
/* seeded generator: restatement of utils/generate_big_dataset.py:9-19                   */
/* ------------------------------------------------------------------------------------ */

/* splitmix64 keyed by (seed,row): the draws of row i do not depend on any other row */
static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

CQO_EXPORT size_t cqo_generate_bigdata_bound(int64_t rows, int64_t key_card) {
    return 64 + (size_t)rows * (size_t)(31 + (key_card > 0 ? 21 : 0));
}

/* name = one of A..P x10, surname = A..P x8, age U{10..80}, gender f|m,
 * height U{100..200}/100 printed like Python's repr(float): 1.0, 1.5, 1.84, 2.0 */
static size_t gen_row(char* o, uint64_t seed, uint64_t i, int64_t key_card) {
    uint64_t r = mix64(seed * 0xD1342543DE82EF95ULL + i);
    uint64_t r2 = mix64(r);
    char* p = o;
    char nm = (char)('A' + (r & 15));
    char sn = (char)('A' + ((r >> 4) & 15));
    unsigned age = 10 + (unsigned)(((r >> 8) & 0xffffff) % 71);
    char gd = ((r >> 32) & 1) ? 'm' : 'f';
    unsigned h = 100 + (unsigned)(((r >> 33) & 0xffffff) % 101);
    memset(p, nm, 10); p += 10; *p++ = ',';
    memset(p, sn, 8); p += 8; *p++ = ',';
    *p++ = (char)('0' + age / 10); *p++ = (char)('0' + age % 10); *p++ = ',';
    *p++ = gd; *p++ = ',';
    *p++ = (char)('0' + h / 100); *p++ = '.';
    unsigned frac = h % 100;
    *p++ = (char)('0' + frac / 10);
    if (frac % 10) *p++ = (char)('0' + frac % 10);
    if (key_card > 0) {
        *p++ = ',';
        uint64_t uid = r2 % (uint64_t)key_card;
        char tmp[24];
        int n = 0;
        do { tmp[n++] = (char)('0' + uid % 10); uid /= 10; } while (uid);
        while (n) *p++ = tmp[--n];
    }
    *p++ = '\n';
    return (size_t)(p - o);
}

CQO_EXPORT int cqo_generate_bigdata(void* host_ptr, size_t capacity, int64_t rows, uint64_t seed,
                                    int64_t key_card, size_t* size_out) {
    char* o = host_ptr;
    const char* hdr = key_card > 0 ? "name,surname,age,gender,height,uid\n" : "name,surname,age,gender,height\n";
    size_t n = strlen(hdr);
    if (capacity < cqo_generate_bigdata_bound(rows, key_card)) return CQG_ERR_ARG;
    memcpy(o, hdr, n);
    for (int64_t i = 0; i < rows; i++) n += gen_row(o + n, seed, (uint64_t)i, key_card);
    *size_out = n;
    return CQG_OK;
}
