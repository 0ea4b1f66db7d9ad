/*
 * cq_dispatch.c — the drop-in seam between cq's host-side C (tokenizer, parser, HAVING /
 * ORDER BY / DISTINCT / LIMIT, output) and the GPU hot path.
 *
 * This translation unit DEFINES `evaluate_query` and `evaluate_query_internal`
 * (include/evaluator.h:32, include/evaluator/evaluator_internal.h:11). It is linked with
 * the reference's own objects, compiled unmodified except that src/evaluator.c is built
 * with -Devaluate_query=cq_ref_evaluate_query
 *      -Devaluate_query_internal=cq_ref_evaluate_query_internal
 * so every caller in cq (main.c, the tests, sub-query re-entry, CREATE TABLE AS) lands
 * here. A supported query shape is planned into a cqg_query_t and run
 * through the C-ABI of include/cq_gpu.h; anything else (DML/DDL, sub-queries, CASE, scalar
 * and window functions, STDDEV/MEDIAN, outer joins, correlated context) is forwarded to
 * the reference's evaluator untouched — those are not operators of the GPU path.
 *
 * The backend prefix is `cqg_` (libcqgpu.so). Building with -DCQ_BACKEND_ORACLE binds the
 * same planner to the CPU restatement in oracle/ (`cqo_`): that variant exists only so the
 * planner + oracle can be checked against the real reference; it is never shipped.
 */
#include <ctype.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "csv_reader.h"
#include "evaluator.h"
#include "evaluator/evaluator_aggregates.h"
#include "evaluator/evaluator_internal.h"
#include "evaluator/evaluator_utils.h"
#include "parser.h"
#include "string_utils.h"

#include "cq_gpu.h"

#ifdef CQ_BACKEND_ORACLE
#define BE(name) cqo_##name
const char* cqo_last_error(void);
int cqo_table_open(const char*, cqg_csv_config_t, cqg_table_t**);
void cqo_table_close(cqg_table_t*);
int cqo_table_column_count(const cqg_table_t*);
const char* cqo_table_column_name(const cqg_table_t*, int);
int cqo_execute(const cqg_table_t*, const cqg_query_t*, cqg_result_t**);
void cqo_result_free(cqg_result_t*);
#else
#define BE(name) cqg_##name
#endif

/* the reference's own evaluator, renamed at compile time (see header comment) */
ResultSet* cq_ref_evaluate_query(ASTNode* query_ast);
ResultSet* cq_ref_evaluate_query_internal(ASTNode* query_ast, Row* outer_row, CsvTable* outer_table);

/* CQ_GPU=0 routes everything to the reference evaluator (oracle mode of the same binary);
 * CQ_GPU_TRACE=1 prints which route a query took. */
static bool gpu_enabled(void) {
    const char* e = getenv("CQ_GPU");
    return !(e && e[0] == '0');
}
static bool trace_enabled(void) {
    const char* e = getenv("CQ_GPU_TRACE");
    return e && e[0] == '1';
}

/* ------------------------------------------------------------------------------------ */
/* plan-time column binding                                                             */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    /* the working table the evaluator would see as ctx->tables[0].table: either the FROM
     * table, or the joined table whose columns are `alias.col`, left then right
     * (evaluator_joins.c:30-37,73-77) */
    int ncols;
    char** names;
    const char* base_alias; /* ctx->tables[0].alias (src/evaluator.c:49) */
    ASTNode* select;
    bool unsupported; /* set when binding needs something only the reference can do */
} binder_t;

/* csv_get_column_index (src/csv_reader.c:500-509) */
static int bind_index(const binder_t* b, const char* name) {
    if (!name) return -1;
    for (int i = 0; i < b->ncols; i++)
        if (strcasecmp(b->names[i], name) == 0) return i;
    return -1;
}

/* find_column_index_with_fallback (evaluator_aggregates.c:20-36) */
static int bind_index_fallback(const binder_t* b, const char* name) {
    if (!name) return -1;
    int i = bind_index(b, name);
    if (i < 0) {
        const char* dot = strchr(name, '.');
        if (dot) i = bind_index(b, dot + 1);
    }
    return i;
}

/* resolve_column (evaluator_core.c:70-167) for table_index 0 with no outer row, done once
 * instead of per row. Returns the column index, or -1 for "NULL value". */
static int bind_resolve(binder_t* b, const char* name) {
    const char* dot = strchr(name, '.');
    if (dot) {
        int i = bind_index(b, name);
        if (i >= 0) return i;
        size_t al = (size_t)(dot - name);
        /* context_get_table: the only table in the context is tables[0] */
        if (strlen(b->base_alias) == al && strncasecmp(b->base_alias, name, al) == 0)
            return bind_index(b, dot + 1);
        return -1;
    }
    int i = bind_index(b, name);
    if (i >= 0) return i;
    /* the SELECT-alias extension (evaluator_core.c:131-159) evaluates an expression per row:
     * leave such queries to the reference */
    if (b->select && b->select->type == NODE_TYPE_SELECT && b->select->select.column_nodes) {
        for (int k = 0; k < b->select->select.column_count; k++) {
            const char* cs = b->select->select.columns[k];
            if (!cs) continue;
            const char* as = cq_strcasestr(cs, " AS ");
            if (as) {
                const char* a = as + 4;
                while (*a && isspace((unsigned char)*a)) a++;
                if (strcasecmp(a, name) == 0) {
                    b->unsupported = true;
                    return -1;
                }
            }
        }
    }
    return -1;
}

/* ------------------------------------------------------------------------------------ */
/* predicate compiler: AST -> postfix cqg_insn_t                                        */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    cqg_insn_t* code;
    int n, cap;
    cqg_value_t* consts;
    int nc, ccap;
    binder_t* b;
    bool ok;
} pc_t;

static void emit(pc_t* p, int op, int a) {
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 32;
        p->code = realloc(p->code, sizeof(cqg_insn_t) * (size_t)p->cap);
    }
    p->code[p->n].op = op;
    p->code[p->n].a = a;
    p->n++;
}

/* literals are folded once with the reference's own parse_value (Q8;
 * evaluator_expressions.c:30-31) */
static int add_const(pc_t* p, const char* lit) {
    if (p->nc == p->ccap) {
        p->ccap = p->ccap ? p->ccap * 2 : 16;
        p->consts = realloc(p->consts, sizeof(cqg_value_t) * (size_t)p->ccap);
    }
    Value v = parse_value(lit, strlen(lit));
    cqg_value_t c;
    memset(&c, 0, sizeof c);
    c.type = (int32_t)v.type;
    switch (v.type) {
        case VALUE_TYPE_INTEGER: c.int_value = v.int_value; break;
        case VALUE_TYPE_DOUBLE: c.double_value = v.double_value; break;
        case VALUE_TYPE_STRING: c.string_value = v.string_value; break; /* ownership moves */
        case VALUE_TYPE_DATE:
            c.date_value.year = v.date_value.year;
            c.date_value.month = v.date_value.month;
            c.date_value.day = v.date_value.day;
            break;
        default: break;
    }
    p->consts[p->nc] = c;
    return p->nc++;
}

static void compile_expr(pc_t* p, ASTNode* e) {
    if (!p->ok) return;
    if (!e) { /* evaluate_expression(NULL) -> NULL value */
        emit(p, CQG_OP_COL, -1);
        return;
    }
    switch (e->type) {
        case NODE_TYPE_LITERAL:
            emit(p, CQG_OP_CONST, add_const(p, e->literal));
            break;
        case NODE_TYPE_IDENTIFIER: {
            int c = bind_resolve(p->b, e->identifier);
            if (p->b->unsupported) p->ok = false;
            emit(p, CQG_OP_COL, c);
            break;
        }
        case NODE_TYPE_BINARY_OP: {
            const char* op = e->binary_op.operator;
            if (!op) {
                p->ok = false;
                return;
            }
            if (!e->binary_op.left || !e->binary_op.right) { /* unary: expressions.c:103-154 */
                ASTNode* operand = e->binary_op.left ? e->binary_op.left : e->binary_op.right;
                if (!operand) {
                    emit(p, CQG_OP_COL, -1);
                    return;
                }
                compile_expr(p, operand);
                if (strcmp(op, "-") == 0) emit(p, CQG_OP_NEG, 0);
                else if (strcmp(op, "+") == 0) emit(p, CQG_OP_POS, 0);
                else { /* any other unary operator yields NULL */
                    emit(p, CQG_OP_POP, 0);
                    emit(p, CQG_OP_COL, -1);
                }
                return;
            }
            compile_expr(p, e->binary_op.left);
            compile_expr(p, e->binary_op.right);
            int code = strcmp(op, "+") == 0 ? CQG_OP_ADD : strcmp(op, "-") == 0 ? CQG_OP_SUB
                     : strcmp(op, "*") == 0 ? CQG_OP_MUL : strcmp(op, "/") == 0 ? CQG_OP_DIV
                     : strcmp(op, "%") == 0 ? CQG_OP_MOD : strcmp(op, "&") == 0 ? CQG_OP_BAND
                     : strcmp(op, "|") == 0 ? CQG_OP_BOR : strcmp(op, "^") == 0 ? CQG_OP_BXOR
                     : CQG_OP_ARITH_NULL;
            emit(p, code, 0);
            break;
        }
        default: /* FUNCTION, SUBQUERY, CASE, WINDOW_FUNCTION: host route */
            p->ok = false;
    }
}

static void compile_cond(pc_t* p, ASTNode* c) {
    if (!p->ok) return;
    if (!c) { /* conditions.c:63 */
        emit(p, CQG_OP_TRUE, 0);
        return;
    }
    if (c->type != NODE_TYPE_CONDITION) { /* conditions.c:65 */
        emit(p, CQG_OP_FALSE, 0);
        return;
    }
    const char* op = c->condition.operator;
    if (!op) {
        p->ok = false;
        return;
    }
    if (strcasecmp(op, "NOT") == 0) {
        compile_cond(p, c->condition.left);
        emit(p, CQG_OP_NOT, 0);
        return;
    }
    if (strcasecmp(op, "AND") == 0 || strcasecmp(op, "OR") == 0) {
        compile_cond(p, c->condition.left);
        compile_cond(p, c->condition.right);
        emit(p, strcasecmp(op, "AND") == 0 ? CQG_OP_AND : CQG_OP_OR, 0);
        return;
    }
    int cmp = strcmp(op, "=") == 0 ? CQG_OP_EQ : (strcmp(op, "!=") == 0 || strcmp(op, "<>") == 0) ? CQG_OP_NE
            : strcmp(op, ">") == 0 ? CQG_OP_GT : strcmp(op, "<") == 0 ? CQG_OP_LT
            : strcmp(op, ">=") == 0 ? CQG_OP_GE : strcmp(op, "<=") == 0 ? CQG_OP_LE : 0;
    if (cmp) {
        /* the reference evaluates the right operand with evaluate_expression even when it is
         * a LIST/SUBQUERY node (-> NULL value); only expression nodes are planned here */
        ASTNode* r = c->condition.right;
        if (r && (r->type == NODE_TYPE_LIST || r->type == NODE_TYPE_SUBQUERY)) {
            p->ok = false;
            return;
        }
        compile_expr(p, c->condition.left);
        compile_expr(p, r);
        emit(p, cmp, 0);
        return;
    }
    if (strcasecmp(op, "IN") == 0 || strcasecmp(op, "NOT IN") == 0) {
        bool not_in = strcasecmp(op, "NOT IN") == 0;
        ASTNode* r = c->condition.right;
        if (!r || r->type != NODE_TYPE_LIST) { /* sub-query IN: host route */
            p->ok = false;
            return;
        }
        compile_expr(p, c->condition.left);
        for (int i = 0; i < r->list.node_count; i++) compile_expr(p, r->list.nodes[i]);
        emit(p, not_in ? CQG_OP_NOT_IN : CQG_OP_IN, r->list.node_count);
        return;
    }
    if (strcasecmp(op, "LIKE") == 0 || strcasecmp(op, "ILIKE") == 0) {
        ASTNode* r = c->condition.right;
        if (r && (r->type == NODE_TYPE_LIST || r->type == NODE_TYPE_SUBQUERY)) {
            p->ok = false;
            return;
        }
        compile_expr(p, c->condition.left);
        compile_expr(p, r);
        emit(p, strcasecmp(op, "LIKE") == 0 ? CQG_OP_LIKE : CQG_OP_ILIKE, 0);
        return;
    }
    /* unknown operator: both operands are evaluated, then `return false` (conditions.c:163) */
    {
        ASTNode* l = c->condition.left;
        ASTNode* r = c->condition.right;
        if ((l && l->type != NODE_TYPE_LITERAL && l->type != NODE_TYPE_IDENTIFIER && l->type != NODE_TYPE_BINARY_OP) ||
            (r && r->type != NODE_TYPE_LITERAL && r->type != NODE_TYPE_IDENTIFIER && r->type != NODE_TYPE_BINARY_OP)) {
            p->ok = false;
            return;
        }
        emit(p, CQG_OP_FALSE, 0);
    }
}

static void pc_free(pc_t* p) {
    for (int i = 0; i < p->nc; i++)
        if (p->consts[i].type == CQG_TYPE_STRING) free(p->consts[i].string_value);
    free(p->consts);
    free(p->code);
}

/* ------------------------------------------------------------------------------------ */
/* ResultSet construction                                                               */
/* ------------------------------------------------------------------------------------ */

/* ownership rules of csv_free (src/csv_reader.c:467-490): every STRING individually malloc'ed */
static Value to_value(const cqg_value_t* v) {
    Value o;
    memset(&o, 0, sizeof o);
    o.type = (ValueType)v->type;
    switch (v->type) {
        case CQG_TYPE_INTEGER: o.int_value = v->int_value; break;
        case CQG_TYPE_DOUBLE: o.double_value = v->double_value; break;
        case CQG_TYPE_STRING: o.string_value = strdup(v->string_value ? v->string_value : ""); break;
        case CQG_TYPE_DATE:
            o.date_value.year = v->date_value.year;
            o.date_value.month = v->date_value.month;
            o.date_value.day = v->date_value.day;
            break;
        default: break;
    }
    return o;
}

static ResultSet* new_result(int ncols) {
    ResultSet* r = calloc(1, sizeof(ResultSet));
    r->filename = strdup("query_result"); /* evaluator_aggregates.c:535 */
    r->has_header = true;
    r->delimiter = ',';
    r->quote = '"';
    r->fd = -1; /* csv_free -> portable_munmap would close(0) otherwise (src/mmap.c:114) */
    r->column_count = ncols;
    r->columns = malloc(sizeof(Column) * (size_t)(ncols > 0 ? ncols : 1));
    return r;
}

/* display name rules of build_aggregated_result (evaluator_aggregates.c:546-593) */
static char* agg_display_name(const char* col_spec) {
    char* alias = extract_column_alias(col_spec);
    if (alias) return alias;
    const char* paren = strchr(col_spec, '(');
    if (paren) {
        char func_buf[256], arg_buf[128], display[512];
        const char* close = strchr(paren, ')');
        size_t fl = (size_t)(paren - col_spec);
        if (fl >= sizeof func_buf) fl = sizeof func_buf - 1;
        memcpy(func_buf, col_spec, fl);
        func_buf[fl] = '\0';
        size_t al = close ? (size_t)(close - (paren + 1)) : strlen(paren + 1);
        if (al >= sizeof arg_buf) al = sizeof arg_buf - 1;
        memcpy(arg_buf, paren + 1, al);
        arg_buf[al] = '\0';
        const char* dot = strchr(arg_buf, '.');
        snprintf(display, sizeof display, "%s(%s)", func_buf, dot ? dot + 1 : arg_buf);
        return strdup(display);
    }
    const char* dot = strchr(col_spec, '.');
    return strdup(dot ? dot + 1 : col_spec);
}

/* display name rules of build_result (evaluator_utils.c:427-457) */
static char* plain_display_name(const char* col_spec) {
    char* alias = extract_column_alias(col_spec);
    if (alias) return alias;
    if (strchr(col_spec, '(')) return strdup(col_spec);
    const char* dot = strchr(col_spec, '.');
    return strdup(dot ? dot + 1 : col_spec);
}

/* ------------------------------------------------------------------------------------ */
/* the planner                                                                          */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int kind; /* 0 = aggregate value[agg], 1 = first-row out[out] */
    int index;
} sel_slot_t;

static bool is_plain_agg(const char* f) {
    return strcasecmp(f, "COUNT") == 0 || strcasecmp(f, "SUM") == 0 || strcasecmp(f, "AVG") == 0 ||
           strcasecmp(f, "MIN") == 0 || strcasecmp(f, "MAX") == 0;
}

static cqg_csv_config_t current_cfg(void) {
    cqg_csv_config_t c;
    c.delimiter = global_csv_config.delimiter;
    c.quote = global_csv_config.quote;
    c.has_header = global_csv_config.has_header ? 1 : 0;
    c.reserved = 0;
    return c;
}

/* Returns a ResultSet, or NULL with *fallback=true when the shape is not on the GPU path,
 * or NULL with *fallback=false on a real error (already reported on stderr). */
static ResultSet* run_on_backend(ASTNode* q, bool* fallback) {
    *fallback = true;
    ASTNode* from = q->query.from;
    ASTNode* select = q->query.select;
    if (!from || from->type != NODE_TYPE_FROM || from->from.subquery || !from->from.table) return NULL;
    if (!select || select->type != NODE_TYPE_SELECT || select->select.column_count <= 0) return NULL;
    if (select->select.column_count > CQG_MAX_OUT_COLS) return NULL;
    if (q->query.join_count > 1) return NULL;

    ASTNode* join = NULL;
    if (q->query.join_count == 1) {
        join = q->query.joins[0];
        if (!join || join->type != NODE_TYPE_JOIN) return NULL;
        if (join->join.join_type != JOIN_TYPE_INNER && join->join.join_type != JOIN_TYPE_LEFT &&
            join->join.join_type != JOIN_TYPE_RIGHT && join->join.join_type != JOIN_TYPE_FULL)
            return NULL;
        ASTNode* on = join->join.condition;
        /* evaluate_join_condition (evaluator_joins.c:40-60) only ever matches `ident = ident` */
        if (!on || on->type != NODE_TYPE_CONDITION || !on->condition.operator ||
            strcmp(on->condition.operator, "=") != 0 || !on->condition.left || !on->condition.right ||
            on->condition.left->type != NODE_TYPE_IDENTIFIER || on->condition.right->type != NODE_TYPE_IDENTIFIER)
            return NULL;
    }

    /* has_aggregate_functions (evaluator_aggregates.c:55-81) */
    ASTNode* gb = q->query.group_by;
    bool grouped = gb && gb->type == NODE_TYPE_GROUP_BY && gb->group_by.columns && gb->group_by.column_count > 0;
    bool aggregated = grouped || has_aggregate_functions(select);
    if (grouped && gb->group_by.column_count > CQG_MAX_GROUP_COLS) return NULL;

    /* ---- open tables ---- */
    cqg_table_t* lt = NULL;
    cqg_table_t* rt = NULL;
    int orc = BE(table_open)(from->from.table, current_cfg(), &lt);
    if (orc == CQG_ERR_UNSUPPORTED_PLAN) return NULL; /* a dialect the kernels do not cover: the reference's route */
    if (orc != CQG_OK) {
        /* same messages as csv_load + load_from_table (src/csv_reader.c:382, joins.c:222) */
        *fallback = false;
        fprintf(stderr, "Error loading file: %s\n", BE(last_error)());
        fprintf(stderr, "Failed to load table from '%s'\n", from->from.table);
        return NULL;
    }
    const char* base_alias = from->from.alias ? from->from.alias : "main"; /* joins.c:226 */
    if (join) {
        if (BE(table_open)(join->join.table, current_cfg(), &rt) != CQG_OK) {
            /* process_joins skips a join whose table fails to load (joins.c:251-254): rare,
             * keep the reference's behaviour by taking its route */
            BE(table_close)(lt);
            return NULL;
        }
    }

    binder_t b;
    memset(&b, 0, sizeof b);
    b.select = select;
    b.base_alias = base_alias;
    int nl = BE(table_column_count)(lt);
    int nr = rt ? BE(table_column_count)(rt) : 0;
    b.ncols = nl + nr;
    b.names = calloc((size_t)(b.ncols > 0 ? b.ncols : 1), sizeof(char*));
    ResultSet* result = NULL;
    pc_t pc;
    memset(&pc, 0, sizeof pc);
    pc.b = &b;
    pc.ok = true;
    sel_slot_t* slots = NULL;
    cqg_result_t* res = NULL;

    if (!join) {
        for (int i = 0; i < nl; i++) b.names[i] = strdup(BE(table_column_name)(lt, i));
    } else {
        const char* ralias = join->join.alias ? join->join.alias : "right"; /* joins.c:256 */
        char buf[256];
        for (int i = 0; i < nl; i++) { /* copy_columns_with_prefix, joins.c:30-37 */
            snprintf(buf, sizeof buf, "%s.%s", base_alias, BE(table_column_name)(lt, i));
            b.names[i] = strdup(buf);
        }
        for (int i = 0; i < nr; i++) {
            snprintf(buf, sizeof buf, "%s.%s", ralias, BE(table_column_name)(rt, i));
            b.names[nl + i] = strdup(buf);
        }
    }

    cqg_query_t plan;
    memset(&plan, 0, sizeof plan);
    plan.max_rows = -1;

    if (join) {
        /* ON operands are resolved positionally: left identifier against the LEFT table,
         * right identifier against the RIGHT table (evaluator_joins.c:49-52 with a two-table
         * context, evaluator_core.c:79-119). */
        ASTNode* on = join->join.condition;
        const char* ralias = join->join.alias ? join->join.alias : "right";
        binder_t lb, rb;
        memset(&lb, 0, sizeof lb);
        memset(&rb, 0, sizeof rb);
        lb.ncols = nl;
        lb.names = calloc((size_t)(nl > 0 ? nl : 1), sizeof(char*));
        for (int i = 0; i < nl; i++) lb.names[i] = (char*)BE(table_column_name)(lt, i);
        rb.ncols = nr;
        rb.names = calloc((size_t)(nr > 0 ? nr : 1), sizeof(char*));
        for (int i = 0; i < nr; i++) rb.names[i] = (char*)BE(table_column_name)(rt, i);
        const char* ids[2] = {on->condition.left->identifier, on->condition.right->identifier};
        binder_t* side[2] = {&lb, &rb};
        int cols[2] = {-1, -1};
        for (int s = 0; s < 2; s++) {
            const char* name = ids[s];
            const char* dot = strchr(name, '.');
            if (dot) {
                int ix = bind_index(side[s], name); /* exact match first */
                if (ix < 0) {
                    /* alias lookup over BOTH tables of the temporary context, then the column
                     * is looked up in THAT table but indexed into THIS side's row */
                    size_t al = (size_t)(dot - name);
                    const binder_t* tb = NULL;
                    if (strlen(base_alias) == al && strncasecmp(base_alias, name, al) == 0) tb = &lb;
                    else if (strlen(ralias) == al && strncasecmp(ralias, name, al) == 0) tb = &rb;
                    if (tb) {
                        ix = bind_index(tb, dot + 1);
                        if (tb != side[s] && ix >= 0) {
                            /* reference indexes the other table's column position into this
                             * row (UB when out of range): not reproduced */
                            free(lb.names);
                            free(rb.names);
                            goto done_fallback;
                        }
                    }
                }
                cols[s] = ix;
            } else {
                cols[s] = bind_index(side[s], name);
                /* SELECT-alias extension would kick in for unknown names: keep it simple */
            }
        }
        free(lb.names);
        free(rb.names);
        plan.join.right = rt;
        plan.join.left_col = cols[0];
        plan.join.right_col = cols[1];
        /* perform_join's join_type (evaluator_joins.c:128-171): unmatched rows of the left / right / both tables */
        plan.join.type = join->join.join_type == JOIN_TYPE_LEFT ? CQG_JOIN_LEFT
                       : join->join.join_type == JOIN_TYPE_RIGHT ? CQG_JOIN_RIGHT
                       : join->join.join_type == JOIN_TYPE_FULL ? CQG_JOIN_FULL : CQG_JOIN_INNER;
        /* an unresolved key column matches nothing: RIGHT / FULL then return the whole right table, which the
         * backend's (empty) join table cannot enumerate - the reference's route */
        if (plan.join.type >= CQG_JOIN_RIGHT && (cols[0] < 0 || cols[1] < 0)) goto done_fallback;
    }

    /* ---- WHERE ---- */
    if (q->query.where) {
        compile_cond(&pc, q->query.where);
        if (!pc.ok) goto done_fallback;
    }
    plan.where.code = pc.code;
    plan.where.n_code = pc.n;
    plan.where.consts = pc.consts;
    plan.where.n_consts = pc.nc;

    int ncols_sel = select->select.column_count;

    if (aggregated) {
        plan.mode = CQG_MODE_AGGREGATE;
        /* ---- GROUP BY keys (src/evaluator.c:78-112) ---- */
        if (grouped) {
            int ng = gb->group_by.column_count;
            for (int g = 0; g < ng; g++) {
                const char* gc = gb->group_by.columns[g];
                ASTNode* gexpr = NULL;
                if (gc && select->select.column_nodes) {
                    for (int i = 0; i < ncols_sel; i++) {
                        const char* cs = select->select.columns[i];
                        if (!cs) continue;
                        const char* as = cq_strcasestr(cs, " AS ");
                        if (as) {
                            const char* a = as + 4;
                            while (*a && isspace((unsigned char)*a)) a++;
                            if (strcasecmp(a, gc) == 0) {
                                gexpr = select->select.column_nodes[i];
                                break;
                            }
                        }
                    }
                }
                if (gexpr) {
                    /* grouping by a SELECT alias evaluates that expression per row; a bare
                     * identifier is just a column, anything else stays on the host */
                    if (gexpr->type != NODE_TYPE_IDENTIFIER) goto done_fallback;
                    plan.group_cols[g] = bind_resolve(&b, gexpr->identifier);
                    if (b.unsupported) goto done_fallback;
                } else if (ng == 1) {
                    plan.group_cols[g] = bind_index_fallback(&b, gc); /* aggregates.c:114 */
                } else {
                    plan.group_cols[g] = bind_index(&b, gc); /* src/evaluator.c:152 */
                }
            }
            plan.n_group_cols = ng;
        }
        /* ---- SELECT list, parsed textually as build_aggregated_result does (Q13) ---- */
        slots = calloc((size_t)ncols_sel, sizeof(sel_slot_t));
        for (int c = 0; c < ncols_sel; c++) {
            const char* spec = select->select.columns[c];
            if (!spec) goto done_fallback;
            char col_name[512];
            const char* as = cq_strcasestr(spec, " AS ");
            size_t cl = as ? (size_t)(as - spec) : strlen(spec);
            if (cl >= sizeof col_name) goto done_fallback;
            memcpy(col_name, spec, cl);
            col_name[cl] = '\0';
            trim_trailing_spaces(col_name);
            char* paren = strchr(col_name, '(');
            if (paren) {
                char func[64];
                size_t fl = (size_t)(paren - col_name);
                if (fl >= sizeof func) goto done_fallback;
                memcpy(func, col_name, fl);
                func[fl] = '\0';
                if (!is_aggregate_function(func)) goto done_fallback; /* scalar fn on first row */
                if (!is_plain_agg(func)) goto done_fallback;          /* STDDEV / MEDIAN */
                char* arg = paren + 1;
                char* close = strchr(arg, ')');
                if (close) *close = '\0'; /* else: the reference keeps the whole col_name */
                else arg = col_name;
                if (plan.n_aggs >= CQG_MAX_AGGS) goto done_fallback;
                cqg_agg_t* a = &plan.aggs[plan.n_aggs];
                if (strcasecmp(func, "COUNT") == 0 && strcmp(arg, "*") == 0) {
                    a->func = CQG_AGG_COUNT_STAR;
                    a->col = -1;
                } else {
                    a->func = strcasecmp(func, "COUNT") == 0 ? CQG_AGG_COUNT
                            : strcasecmp(func, "SUM") == 0 ? CQG_AGG_SUM
                            : strcasecmp(func, "AVG") == 0 ? CQG_AGG_AVG
                            : strcasecmp(func, "MIN") == 0 ? CQG_AGG_MIN : CQG_AGG_MAX;
                    a->col = bind_index_fallback(&b, arg); /* aggregates.c:274 */
                }
                slots[c].kind = 0;
                slots[c].index = plan.n_aggs++;
            } else {
                ASTNode* node = select->select.column_nodes ? select->select.column_nodes[c] : NULL;
                if (node && node->type != NODE_TYPE_IDENTIFIER) goto done_fallback; /* expr on first row */
                slots[c].kind = 1;
                slots[c].index = plan.n_out_cols;
                plan.out_cols[plan.n_out_cols++] = bind_index_fallback(&b, col_name); /* :680 */
            }
        }
    } else {
        plan.mode = CQG_MODE_SELECT;
        /* build_result (evaluator_utils.c:249-549) for `*` and bare identifiers */
        for (int c = 0; c < ncols_sel; c++) {
            const char* spec = select->select.columns[c];
            if (!spec) goto done_fallback;
            if (strcmp(spec, "*") == 0) {
                if (plan.n_out_cols + b.ncols > CQG_MAX_OUT_COLS) goto done_fallback;
                for (int j = 0; j < b.ncols; j++) plan.out_cols[plan.n_out_cols++] = j;
                continue;
            }
            ASTNode* node = select->select.column_nodes ? select->select.column_nodes[c] : NULL;
            if (!node || node->type != NODE_TYPE_IDENTIFIER) goto done_fallback;
            if (plan.n_out_cols >= CQG_MAX_OUT_COLS) goto done_fallback;
            plan.out_cols[plan.n_out_cols++] = bind_resolve(&b, node->identifier);
            if (b.unsupported) goto done_fallback;
        }
        /* LIMIT can stop the fetch early only when nothing reorders or dedups afterwards */
        bool ordered = q->query.order_by && q->query.order_by->type == NODE_TYPE_ORDER_BY && q->query.order_by->order_by.column;
        if (!ordered && !select->select.distinct && q->query.limit >= 0)
            plan.max_rows = (int64_t)q->query.limit + (q->query.offset > 0 ? q->query.offset : 0);
    }

    /* ---- run ---- */
    *fallback = false;
    int rc = BE(execute)(lt, &plan, &res);
    if (rc == CQG_ERR_UNSUPPORTED_PLAN) {
        /* declined at plan time (predicate depth, column count ...), before any table byte was uploaded: not an
         * operator of this path, the reference evaluates it like every other unsupported shape */
        *fallback = true;
        goto done;
    }
    if (rc == CQG_ERR_UNSUPPORTED) {
        /* The operators of this shape belong to the GPU path: there is no silent CPU route for
         * them. The statement fails with the reason, unless the operator explicitly opted into
         * the reference evaluator for such inputs (CQ_GPU_FALLBACK=1). */
        const char* fb = getenv("CQ_GPU_FALLBACK");
        if (fb && fb[0] == '1') {
            fprintf(stderr, "[cq-gpu] backend declined (%s): CQ_GPU_FALLBACK=1, using the reference evaluator\n",
                    BE(last_error)());
            *fallback = true;
        } else {
            fprintf(stderr, "GPU query execution declined: %s\n", BE(last_error)());
        }
        goto done;
    }
    if (rc != CQG_OK) {
        fprintf(stderr, "GPU query execution failed: %s\n", BE(last_error)());
        goto done;
    }

    /* ---- ResultSet ---- */
    if (aggregated) {
        result = new_result(ncols_sel);
        for (int c = 0; c < ncols_sel; c++) {
            result->columns[c].name = agg_display_name(select->select.columns[c]);
            result->columns[c].inferred_type = VALUE_TYPE_STRING; /* aggregates.c:592 */
        }
        int64_t G = res->n_groups;
        result->row_count = (int)G;
        result->row_capacity = (int)G;
        result->rows = malloc(sizeof(Row) * (size_t)(G > 0 ? G : 1));
        for (int64_t g = 0; g < G; g++) {
            result->rows[g].column_count = ncols_sel;
            result->rows[g].values = malloc(sizeof(Value) * (size_t)ncols_sel);
            for (int c = 0; c < ncols_sel; c++) {
                const cqg_value_t* v = slots[c].kind == 0
                                           ? &res->value[(size_t)slots[c].index * (size_t)G + (size_t)g]
                                           : &res->out[(size_t)slots[c].index * (size_t)G + (size_t)g];
                result->rows[g].values[c] = to_value(v);
            }
        }
        /* the reference's own post passes (src/evaluator.c:223-231, 249-258) */
        if (q->query.having) apply_having_filter(result, q->query.having, select);
        ASTNode* ob = q->query.order_by;
        if (ob && ob->type == NODE_TYPE_ORDER_BY && ob->order_by.column)
            sort_result(result, select, ob->order_by.column, ob->order_by.descending);
    } else {
        int nout = plan.n_out_cols;
        result = new_result(nout);
        int k = 0;
        for (int c = 0; c < ncols_sel; c++) {
            const char* spec = select->select.columns[c];
            if (strcmp(spec, "*") == 0) {
                for (int j = 0; j < b.ncols; j++) {
                    result->columns[k].name = strdup(b.names[j]);
                    result->columns[k].inferred_type = VALUE_TYPE_STRING;
                    k++;
                }
            } else {
                result->columns[k].name = plain_display_name(spec);
                result->columns[k].inferred_type = VALUE_TYPE_STRING;
                k++;
            }
        }
        int64_t N = res->n_rows_out;
        result->row_count = (int)N;
        result->row_capacity = (int)N;
        result->rows = malloc(sizeof(Row) * (size_t)(N > 0 ? N : 1));
        for (int64_t i = 0; i < N; i++) {
            result->rows[i].column_count = nout;
            result->rows[i].values = malloc(sizeof(Value) * (size_t)(nout > 0 ? nout : 1));
            for (int c = 0; c < nout; c++)
                result->rows[i].values[c] = to_value(&res->rows[(size_t)i * (size_t)nout + (size_t)c]);
        }
        ASTNode* ob = q->query.order_by;
        if (ob && ob->type == NODE_TYPE_ORDER_BY && ob->order_by.column)
            sort_result(result, select, ob->order_by.column, ob->order_by.descending);
    }
    if (select->select.distinct) apply_distinct(result);          /* src/evaluator.c:279-281 */
    apply_limit_offset(result, q->query.limit, q->query.offset);  /* :284 */
    goto done;

done_fallback:
    *fallback = true;
done:
    if (res) BE(result_free)(res);
    free(slots);
    pc_free(&pc);
    for (int i = 0; i < b.ncols; i++) free(b.names[i]);
    free(b.names);
    if (rt) BE(table_close)(rt);
    BE(table_close)(lt);
    return result;
}

/* ------------------------------------------------------------------------------------ */
/* csv_load for the callers that need the array of structs                  */
/* ------------------------------------------------------------------------------------ */
/* Every shape that keeps the reference's route (CASE, scalar functions, sub-queries, multi-join chains, window
 * functions) and every DML statement starts with csv_load (evaluator_joins.c:219, :249; evaluator_statements.c), the
 * single-threaded parse that dominates the reference's run time. src/csv_reader.c is compiled with that one
 * name changed to cq_ref_csv_load and this csv_load stands in: split + typed decode on the GPU (a `SELECT *` projection
 * and cqg_table_field_counts), then the same CsvTable csv_load builds - Row::column_count per row as parse_line counts
 * fields, names and `$n` names, inferred types from the first 20 rows (src/csv_reader.c:429-462). Anything it cannot
 * hand back identically goes to the reference's own csv_load: a file that does not open (so that the messages are the
 * reference's), more columns than a projection carries, a row with MORE fields than the header (their values are not
 * part of a projection), a dialect or datum the kernels decline. CQ_GPU=0 and CQ_GPU_LOAD=0 switch it off. */
#ifndef CQ_BACKEND_ORACLE
CsvTable* cq_ref_csv_load(const char* filename, CsvConfig config);

static CsvTable* gpu_csv_load(const char* filename, CsvConfig config) {
    cqg_csv_config_t cfg;
    cfg.delimiter = config.delimiter;
    cfg.quote = config.quote;
    cfg.has_header = config.has_header ? 1 : 0;
    cfg.reserved = 0;
    cqg_table_t* t = NULL;
    if (cqg_table_open(filename, cfg, &t) != CQG_OK) return NULL;
    CsvTable* table = NULL;
    cqg_result_t* res = NULL;
    int32_t* fc = NULL;
    int ncols = cqg_table_column_count(t);
    if (ncols < 1 || ncols > CQG_MAX_OUT_COLS) goto out;
    cqg_query_t q;
    memset(&q, 0, sizeof q);
    q.mode = CQG_MODE_SELECT;
    q.n_out_cols = ncols;
    for (int c = 0; c < ncols; c++) q.out_cols[c] = c;
    q.max_rows = -1;
    if (cqg_execute(t, &q, &res) != CQG_OK) goto out;
    int64_t N = res->n_rows_out;
    if (N > 0x7fffffff) goto out;
    fc = malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    if (!fc || cqg_table_field_counts(t, res->row_offset, N, fc) != CQG_OK) goto out;
    for (int64_t i = 0; i < N; i++)
        if (fc[i] > ncols) goto out; /* fields beyond the header: only the reference's loader keeps them */

    table = calloc(1, sizeof(CsvTable));
    table->filename = strdup(filename);
    table->data = NULL; /* csv_free: nothing to unmap */
    table->file_size = cqg_table_size(t);
    table->fd = -1;
    table->delimiter = config.delimiter;
    table->quote = config.quote;
    table->has_header = config.has_header;
    table->column_count = ncols;
    table->columns = malloc(sizeof(Column) * (size_t)ncols);
    for (int c = 0; c < ncols; c++) {
        table->columns[c].name = strdup(cqg_table_column_name(t, c));
        table->columns[c].inferred_type = VALUE_TYPE_STRING;
    }
    table->row_count = (int)N;
    table->row_capacity = (int)N;
    table->rows = N > 0 ? malloc(sizeof(Row) * (size_t)N) : NULL;
    for (int64_t i = 0; i < N; i++) {
        const int n = fc[i];
        table->rows[i].column_count = n;
        table->rows[i].values = malloc(sizeof(Value) * (size_t)(n > 0 ? n : 1));
        for (int c = 0; c < n; c++) table->rows[i].values[c] = to_value(&res->rows[(size_t)i * (size_t)ncols + (size_t)c]);
    }
    /* inferred column types: src/csv_reader.c:429-462 */
    if (table->row_count > 0) {
        const int sample = table->row_count < 20 ? table->row_count : 20;
        for (int c = 0; c < ncols; c++) {
            int counts[5] = {0, 0, 0, 0, 0};
            for (int r = 0; r < sample; r++)
                if (c < table->rows[r].column_count) {
                    const ValueType ty = table->rows[r].values[c].type;
                    if (ty >= 0 && ty < 5) counts[ty]++;
                }
            table->columns[c].inferred_type = counts[VALUE_TYPE_DATE] > 0      ? VALUE_TYPE_DATE
                                              : counts[VALUE_TYPE_DOUBLE] > 0  ? VALUE_TYPE_DOUBLE
                                              : counts[VALUE_TYPE_INTEGER] > 0 ? VALUE_TYPE_INTEGER
                                                                               : VALUE_TYPE_STRING;
        }
    }
out:
    free(fc);
    if (res) cqg_result_free(res);
    cqg_table_close(t);
    return table;
}

CsvTable* csv_load(const char* filename, CsvConfig config) {
    const char* e = getenv("CQ_GPU_LOAD");
    if (filename && gpu_enabled() && !(e && e[0] == '0')) {
        CsvTable* table = gpu_csv_load(filename, config);
        if (table) {
            if (trace_enabled()) fprintf(stderr, "[cq-gpu] csv_load=gpu %s\n", filename);
            return table;
        }
    }
    if (trace_enabled()) fprintf(stderr, "[cq-gpu] csv_load=reference %s\n", filename ? filename : "(null)");
    return cq_ref_csv_load(filename, config);
}
#endif

/* ------------------------------------------------------------------------------------ */
/* the two symbols cq binds to                                                          */
/* ------------------------------------------------------------------------------------ */

ResultSet* evaluate_query_internal(ASTNode* query_ast, Row* outer_row, CsvTable* outer_table) {
    if (!query_ast || query_ast->type != NODE_TYPE_QUERY) {
        fprintf(stderr, "Invalid query AST\n"); /* src/evaluator.c:28 */
        return NULL;
    }
    if (gpu_enabled() && !outer_row && !outer_table) {
        bool fallback = true;
        ResultSet* r = run_on_backend(query_ast, &fallback);
        if (r || !fallback) {
            if (trace_enabled()) fprintf(stderr, "[cq-gpu] route=gpu\n");
            return r;
        }
    }
    if (trace_enabled()) fprintf(stderr, "[cq-gpu] route=reference\n");
    return cq_ref_evaluate_query_internal(query_ast, outer_row, outer_table);
}

ResultSet* evaluate_query(ASTNode* query_ast) {
    if (!query_ast) return NULL;
    if (query_ast->type == NODE_TYPE_SET_OP) {
        /* same dispatch as src/evaluator.c:307-345, but each side re-enters THIS
         * evaluate_query so both arms can take the GPU route */
        ResultSet* left = evaluate_query(query_ast->set_op.left);
        if (!left) return NULL;
        ResultSet* right = evaluate_query(query_ast->set_op.right);
        if (!right) {
            csv_free(left);
            return NULL;
        }
        if (left->column_count != right->column_count) {
            fprintf(stderr, "Error: SET operation queries must have the same number of columns\n");
            csv_free(left);
            csv_free(right);
            return NULL;
        }
        ResultSet* result = NULL;
        switch (query_ast->set_op.op_type) {
            case SET_OP_UNION: result = set_union(left, right, false); break;
            case SET_OP_UNION_ALL: result = set_union(left, right, true); break;
            case SET_OP_INTERSECT: result = set_intersect(left, right); break;
            case SET_OP_EXCEPT: result = set_except(left, right); break;
        }
        csv_free(left);
        csv_free(right);
        return result;
    }
    if (query_ast->type != NODE_TYPE_QUERY) return cq_ref_evaluate_query(query_ast); /* DML / DDL */
    return evaluate_query_internal(query_ast, NULL, NULL);
}
