/*
 * cq_dump.c — run one SQL statement through parse() + evaluate_query() and print every
 * Value exactly (type tag, int64, hex-float double, string bytes, date), because the CLI
 * rounds doubles to %.2f (src/csv_reader.c:88).
 *
 * The same source is linked three ways (oracle/Makefile, cq_b200/build.py):
 *   oracle/_ref/ref_dump        reference objects only            -> the true oracle
 *   oracle/_ref/oracle_dump     cq_dispatch.c + oracle backend    -> pins the restatement
 *   build/cq_gpu_dump           cq_dispatch.c + libcqgpu.so       -> the product
 *
 * usage: dump [-s delimiter] [-n] "SQL"      (-n: has_header = false)
 * output:
 *   #rows <n> cols <m>
 *   #col <name>            (m lines)
 *   one line per row, values separated by \t:
 *     N | I:<lld> | D:<%a> | S:<len>:<bytes> | T:<y>-<m>-<d>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csv_reader.h"
#include "evaluator.h"
#include "parser.h"

int main(int argc, char** argv) {
    int i = 1;
    while (i < argc && argv[i][0] == '-' && argv[i][1] && !argv[i][2]) {
        if (argv[i][1] == 's' && i + 1 < argc) {
            global_csv_config.delimiter = argv[i + 1][0];
            i += 2;
        } else if (argv[i][1] == 'n') {
            global_csv_config.has_header = false;
            i++;
        } else {
            break;
        }
    }
    if (i >= argc) {
        fprintf(stderr, "usage: %s [-s delim] [-n] \"SQL\"\n", argv[0]);
        return 2;
    }
    ASTNode* ast = parse(argv[i]);
    if (!ast) {
        printf("#error parse\n");
        return 1;
    }
    ResultSet* r = evaluate_query(ast);
    if (!r) {
        printf("#error eval\n");
        releaseNode(ast);
        return 1;
    }
    printf("#rows %d cols %d\n", r->row_count, r->column_count);
    for (int c = 0; c < r->column_count; c++) printf("#col %s\n", r->columns[c].name ? r->columns[c].name : "");
    for (int row = 0; row < r->row_count; row++) {
        Row* rw = &r->rows[row];
        for (int c = 0; c < rw->column_count; c++) {
            Value* v = &rw->values[c];
            if (c) putchar('\t');
            switch (v->type) {
                case VALUE_TYPE_NULL: putchar('N'); break;
                case VALUE_TYPE_INTEGER: printf("I:%lld", v->int_value); break;
                case VALUE_TYPE_DOUBLE: printf("D:%a", v->double_value); break;
                case VALUE_TYPE_STRING: {
                    const char* s = v->string_value ? v->string_value : "";
                    printf("S:%zu:", strlen(s));
                    fwrite(s, 1, strlen(s), stdout);
                    break;
                }
                case VALUE_TYPE_DATE:
                    printf("T:%d-%d-%d", v->date_value.year, v->date_value.month, v->date_value.day);
                    break;
                default: printf("?%d", (int)v->type);
            }
        }
        putchar('\n');
    }
    csv_free(r);
    releaseNode(ast);
    return 0;
}
