"""Plan-level parity cases shared by the GPU tests, smoke() and bench.py's self-check.

Columns of the generated table: 0 name, 1 surname, 2 age, 3 gender, 4 height [, 5 uid].
"""
from cq_b200 import _abi as A
from cq_b200.engine import Plan

NAME, SURNAME, AGE, GENDER, HEIGHT, UID = range(6)


def col(i):
    return ("col", i)


def const(v):
    return ("const", v)


def plans():
    P = {}
    P["count_age_gt_40"] = dict(where=(">", col(AGE), const(40)), aggs=[(A.AGG_COUNT_STAR, -1)])
    P["count_height_gt_1_5"] = dict(where=(">", col(HEIGHT), const(1.5)), aggs=[(A.AGG_COUNT_STAR, -1)])
    P["scalar_aggs"] = dict(where=("and", (">", col(AGE), const(25)), ("=", col(GENDER), const("f"))),
                            aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE), (A.AGG_AVG, HEIGHT), (A.AGG_MIN, HEIGHT)])
    P["scalar_minmax"] = dict(aggs=[(A.AGG_MIN, AGE), (A.AGG_MAX, AGE), (A.AGG_MIN, NAME), (A.AGG_MAX, SURNAME)])
    P["scalar_many"] = dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE), (A.AGG_AVG, AGE), (A.AGG_MIN, AGE),
                                  (A.AGG_MAX, AGE), (A.AGG_SUM, HEIGHT), (A.AGG_COUNT, NAME)])
    P["group_name"] = dict(where=(">", col(AGE), const(25)), group_by=[NAME], out_cols=[NAME],
                           aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, HEIGHT), (A.AGG_SUM, AGE)])
    P["group_gender_minmax"] = dict(group_by=[GENDER], out_cols=[GENDER, NAME],
                                    aggs=[(A.AGG_MIN, HEIGHT), (A.AGG_MAX, HEIGHT), (A.AGG_MIN, SURNAME), (A.AGG_MAX, AGE)])
    P["group_age"] = dict(group_by=[AGE], out_cols=[AGE], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, HEIGHT)])
    P["group_height"] = dict(where=("<", col(AGE), const(50)), group_by=[HEIGHT], out_cols=[HEIGHT],
                             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE)])
    P["group_name_surname"] = dict(group_by=[NAME, SURNAME], out_cols=[NAME, SURNAME],
                                   aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE), (A.AGG_MIN, HEIGHT)])
    P["group_high_card"] = dict(group_by=[NAME, SURNAME, AGE, HEIGHT], out_cols=[NAME, SURNAME, AGE, HEIGHT],
                                aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE), (A.AGG_MIN, HEIGHT), (A.AGG_MAX, HEIGHT),
                                      (A.AGG_AVG, HEIGHT)])
    # lean GROUP BY kernel shapes: few groups (stays lean), and more groups than a CTA numbers (aborts to general)
    P["lean_group_two_keys"] = dict(where=("<=", col(HEIGHT), const(1.9)), group_by=[NAME, GENDER], out_cols=[NAME, GENDER],
                                    aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE), (A.AGG_AVG, HEIGHT), (A.AGG_COUNT, SURNAME)])
    P["lean_group_gender_height"] = dict(group_by=[GENDER, HEIGHT], out_cols=[GENDER, HEIGHT],
                                         aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, HEIGHT)])
    P["lean_group_abort_many"] = dict(where=(">", col(AGE), const(15)), group_by=[NAME, SURNAME, AGE],
                                      out_cols=[NAME, SURNAME, AGE], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, HEIGHT)])
    P["filter_between_in"] = dict(where=("and", ("and", (">=", col(AGE), const(20)), ("<=", col(AGE), const(60))),
                                         ("in", col(NAME), [const("AAAAAAAAAA"), const("CCCCCCCCCC"), const(5)])),
                                  aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE)])
    P["filter_like"] = dict(where=("or", ("like", col(NAME), const("%B%")), ("ilike", col(SURNAME), const("c_c%"))),
                            group_by=[GENDER], out_cols=[GENDER], aggs=[(A.AGG_COUNT_STAR, -1)])
    P["filter_arith"] = dict(where=(">", ("+", ("*", col(AGE), const(2)), col(HEIGHT)), const(100.5)),
                             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, HEIGHT)])
    P["filter_mod_bits"] = dict(where=("and", ("=", ("%", col(AGE), const(3)), const(1)),
                                       ("!=", ("&", col(AGE), const(4)), const(0))),
                                aggs=[(A.AGG_COUNT_STAR, -1)])
    P["filter_cross_type"] = dict(where=(">=", col(NAME), const(15)), aggs=[(A.AGG_COUNT_STAR, -1)])
    P["filter_not"] = dict(where=("not", ("or", ("<", col(HEIGHT), const(1.2)), ("=", col(GENDER), const("m")))),
                           aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, AGE)])
    P["select_rows"] = dict(mode="select", where=("and", ("=", col(AGE), const(33)), ("=", col(HEIGHT), const(1.5))),
                            out_cols=[NAME, AGE, HEIGHT, GENDER, 9])
    P["select_limit"] = dict(mode="select", where=(">", col(AGE), const(70)), out_cols=[SURNAME, AGE], max_rows=25)
    return P


def plans_uid():
    P = {}
    P["group_uid"] = dict(group_by=[UID], out_cols=[UID],
                          aggs=[(A.AGG_SUM, AGE), (A.AGG_MIN, HEIGHT), (A.AGG_MAX, HEIGHT), (A.AGG_AVG, HEIGHT)])
    # results of many groups are finished on the device: every cell kind (string / numeric MIN and MAX, COUNT(col),
    # AVG, NULL for an unknown column, bare columns incl. text) and the general kernel's entries
    P["group_uid_cell_kinds"] = dict(group_by=[UID], out_cols=[UID, NAME, HEIGHT, 11],
                                     aggs=[(A.AGG_MIN, NAME), (A.AGG_MAX, SURNAME), (A.AGG_COUNT, GENDER), (A.AGG_MIN, AGE),
                                           (A.AGG_AVG, HEIGHT), (A.AGG_SUM, 12), (A.AGG_MAX, HEIGHT)])
    P["group_uid_keys_only"] = dict(where=("<", col(AGE), const(60)), group_by=[UID, GENDER], out_cols=[GENDER, UID])
    P["lean_group_uid_sum"] = dict(where=(">=", col(AGE), const(18)), group_by=[UID], out_cols=[UID],
                                   aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE), (A.AGG_AVG, HEIGHT)])
    return P


def build(spec, join=None):
    spec = dict(spec)
    if join is not None:
        spec["join"] = join
    return Plan(**spec)


def compare_results(got, want, rel=1e-12):
    """Result dicts of engine.decode_result: integers, keys, MIN/MAX and row sets bit-exact;
    SUM/AVG (DOUBLE aggregates) within `rel` relative (BASELINE.json north_star)."""
    def veq(x, y):
        if x[0] != y[0]:
            return False
        if x[0] == "D":
            a, b = x[1], y[1]
            return a == b or abs(a - b) <= rel * max(abs(a), abs(b))
        return x == y

    if "groups" in want:
        assert got["rows_scanned"] == want["rows_scanned"], (got["rows_scanned"], want["rows_scanned"])
        assert len(got["groups"]) == len(want["groups"]), (len(got["groups"]), len(want["groups"]))
        for i, (g, w) in enumerate(zip(got["groups"], want["groups"])):
            assert g["first_offset"] == w["first_offset"], (i, g, w)
            assert g["count"] == w["count"], (i, g, w)
            assert g["out"] == w["out"], (i, g["out"], w["out"])
            for a, (x, y) in enumerate(zip(g["aggs"], w["aggs"])):
                assert veq(x, y), (i, a, x, y)
            for a, (x, y) in enumerate(zip(g["ncount"], w["ncount"])):
                assert x == y, (i, a, x, y)
    else:
        assert got["n_selected"] == want["n_selected"]
        assert got["row_offset"] == want["row_offset"]
        assert got["row_offset_right"] == want["row_offset_right"]
        assert got["rows"] == want["rows"]
