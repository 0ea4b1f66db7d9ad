#!/usr/bin/env python
"""Generate tests/fixtures/*.csv and tests/golden/sql_golden.json.

Runs HERE only (needs /root/reference compiled into oracle/_ref/ by `make -C oracle ref`).
Every expected output is what the UNMODIFIED reference prints through oracle/_ref/ref_dump
(cq_b200/host/cq_dump.c linked against the reference's own objects): exact int64, hex-float
doubles, string bytes, dates. The fixtures are this repo's own data (nothing is copied from
the reference's data/ directory); cases tagged "refdata" run on the reference's shipped
fixtures in place and are only replayed where /root/reference exists.

usage: python tests/golden/make_golden.py
"""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FIX = os.path.join(ROOT, "tests", "fixtures")
REF = os.environ.get("CQ_REF", "/root/reference")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
ORACLE_DUMP = os.path.join(ROOT, "oracle", "_ref", "oracle_dump")


def w(name, text, binary=False):
    with open(os.path.join(FIX, name), "wb") as f:
        f.write(text if binary else text.encode("utf-8"))


def make_fixtures():
    os.makedirs(FIX, exist_ok=True)
    # a delimiter that is a letter (strtod in the reference reads through it: the GPU dialect check declines it)
    w("alpha.csv", "idxval\n1x10\n2x20\n3x30\n")
    # ---- people: the shape of the reference's users table, own content ----
    w("people.csv", """id,name,age,role,height,active,email,city,joined,tail
1,Ada,36,admin,170.25,1,ada@example.com,London,2021-03-04,x
2,Brook,29,user,181.5,1,brook@gmail.com,Oslo,2020-11-30,x
3,Cyrus,41,moderator,175.0,0,cyrus@company.org,London,2019-01-15,x
4,Dana,25,user,162.75,1,dana@example.com,Quito,2022-07-01,x
5,Eli,33,user,190.1,0,eli@gmail.com,Oslo,2021-03-04,x
6,Fay,52,admin,158.0,1,fay@company.org,Lima,2018-05-20,x
7,Gus,27,moderator,177.3,1,gus@example.com,Quito,2023-02-28,x
8,Hana,19,user,166.6,1,hana@mail.net,London,2024-01-09,x
9,Ivo,45,user,172.4,0,ivo@gmail.com,Lima,2017-12-31,x
10,Jude,30,user,180.0,1,jude@example.com,Oslo,2020-02-29,x
11,Kim,26,guest,,1,kim@mail.net,,2022-10-10,x
12,Lev,,user,169.9,0,,Quito,,x
13,Mona,38,admin,171.0,1,mona@company.org,London,2019-06-06,x
14,Ned,24,user,183.25,1,ned@gmail.com,Oslo,2023-08-15,x
15,Oda,31,moderator,160.5,0,oda@example.com,Lima,2021-09-09,x
""")
    # ---- typing quirks: one value per row in column `val` ----
    vals = ["1", "007", "-5", "+8", "3.25", ".5", "5.", "-0.0", "0.0", "1e5", "-", "+", ".", "abc", "NULL", "",
            " 12 ", " 1.5 ", "12 3", "20240115", "2010010100", "2023-1-5x", "555-0001", "2024-01-15", "01/02/2024",
            "31/12/2023", "13/13/2023", "2023-02-29", "2024-02-29", "9223372036854775807", "9223372036854775808",
            "-9223372036854775809", "123456789012345678901234", "0.1", "0.30000000000000004", "1.7976931348623157",
            "123456789.123456789", "99999999999999999999.5", "0.000001", "0.0000005", "1.0000001", "1.0000002",
            "2.5", "2.50", "-2.5", "1234567", "12345678", "123456789", "1234567890", "12345678901", "+20240115",
            "-20240115", "2024/01/15", "1/2/3", "10000101", "99991231", "00010101", "19000229", "20000229",
            "x2024-01-15", "2024-01-15x", "Hello World", "hello world", "  padded  ", "tab\there", "a%b", "a_b",
            "100%", "4.0", "4", "4.000000", "-4", "1.5e3", "0x10", "inf", "nan", "1,5"]
    rows = ["id,val,tag"]
    for i, v in enumerate(vals):
        if "," in v:
            v = '"' + v + '"'
        rows.append(f"{i},{v},t{i % 3}")
    w("types.csv", "\n".join(rows) + "\n")
    # ---- quoting / splitting quirks (Q1, Q2). Columns touched by queries exist in every row. ----
    w("quoted.csv", '''id,name,role,note
1,"Last, First",admin,"plain"
2,"say ""hi""",user,"x,y,z"
3,  spaced  ,user,  lead
4,"quoted" tail,moderator,after
5,"  inner pad  ",admin,""
6,unquoted "mid" quote,user,q
7,"a,b",admin,"c,d"
8,xylophone,user,"has x inside"
9,"comma, and ""quote""",moderator,n9
10,last,admin,"unterminated, rest of line
11,after,user,fine
12,"",guest,emptyq
13,"tab\tin",user,"t"
''')
    w("crlf.csv", "id,role,age\r\n1,admin,30\r\n2,user,25\r\n\r\n3,user,41\r\n4,admin,52", binary=False)
    w("blank_lines.csv", "\n\nid,role,age\n\n1,admin,30\n\n\n2,user,25\n3,user,41\n\n")
    w("noheader.csv", "1,admin,30\n2,user,25\n3,user,41\n4,guest,19\n")
    w("semi.csv", "id;name;score;z\n1;ann;3.5;a\n2;bob;4;b\n3;\"c;d\";5.25;c\n4;eve;;d\n")
    w("tabs.csv", "id\tname\tscore\n1\tann\t3.5\n2\tbob\t4\n3\teve\t5.25\n")
    # ---- join tables ----
    w("orders.csv", """id,price,tax,quantity,customer_id,note
1,100.00,10.00,2,1,n1
2,50.50,5.05,1,2,n2
3,200.00,20.00,3,3,n3
4,75.25,7.53,1,1,n4
5,10.00,1.00,5,4,n5
6,300.10,30.01,1,2,n6
7,12.34,1.23,2,9,n7
8,99.99,10.00,1,,n8
9,45.00,4.50,4,3,n9
10,5.00,0.50,10,1,n10
11,60.00,6.00,1,2.0,n11
12,70.00,7.00,2,5,n12
""")
    w("customers.csv", """id,name,email,year
1,Ada,ada@example.com,2023
2,Brook,brook@example.com,2023
3,Cyrus,cyrus@example.com,2024
4,Dana,dana@example.com,2024
5,Eli,eli@example.com,2022
6,Fay,fay@example.com,2022
,Ghost,ghost@example.com,2021
3,Cyrus II,cyrus2@example.com,2025
""")
    w("tags.csv", "code,label\nadmin,Administrators\nuser,Users\nmoderator,Mods\nuser,Users again\nnobody,Nobody\n")
    # ---- LIKE corpus (the shapes tests/test_like.c exercises, own data) ----
    w("like.csv", """id,name,email,product
1,Alma,alma@example.com,USB-001
2,Alec,alec@example.org,USB-002
3,Beate,beate@example.com,USBX
4,alva,alva@Example.com,HDMI-10
5,Al,al@test.io,USB-12
6,Clive,clive@example.com,usb-003
7,Olive,olive@mail.com,DP-1
8,Ali_,ali@x.y,100%
""")
    # ---- a generated table for GROUP BY at a few thousand rows ----
    rnd = random.Random(20261018)
    lines = ["name,surname,age,gender,height,dept,bonus,end"]
    depts = ["eng", "ops", "sales", "NULL", "", "hr"]
    for i in range(3000):
        nm = rnd.choice("ABCDEFGHIJKLMNOP") * 10
        sn = rnd.choice("ABCDEFGHIJKLMNOP") * 8
        age = rnd.randint(10, 80)
        g = rnd.choice("fm")
        h = rnd.randint(100, 200) / 100
        dept = rnd.choice(depts)
        bonus = rnd.choice(["", "0", "10", "2.5", "-3", "7", "7.0", "1000000", "0.1", "0.2"])
        lines.append(f"{nm},{sn},{age},{g},{h},{dept},{bonus},e")
    w("big3k.csv", "\n".join(lines) + "\n")
    # ---- doubles as group keys (%.6f identity) and mixed-type aggregates ----
    w("keys.csv", """k,v,d
1.0000001,1,2024-01-01
1.0000002,2,2024-01-02
1.0000004,3,2023-12-31
1.0000006,4,2024-01-01
-0.0,5,2024-03-01
0.0,6,2024-03-01
0.0000004,7,2022-02-02
-0.0000004,8,2022-02-02
2.5,9,2021-01-01
2.50,10,2021-01-01
2.5000005,11,2021-01-01
2.5000015,12,2020-01-01
1,13,2020-01-01
1.0,14,2020-01-01
abc,15,2020-06-06
NULL,16,2020-06-06
,17,2020-06-06
2024-01-15,18,2020-06-06
20240115,19,2020-06-06
123456.1234565,20,2020-06-06
123456.1234575,21,2020-06-06
""")
    # rows with FEWER fields than the header (csv_load keeps Row::column_count per row; queries below touch column 0 only,
    # which every row has: the reference reads out of bounds otherwise) and a trailing delimiter (no field)
    w("short.csv", "id,name,score\n1,ann,3.5\n2,bob\n3\n4,eve,5,\n5,\"q,x\",6\n")
    # a row with MORE fields than the header: only the reference's own csv_load keeps the extra value
    w("long.csv", "id,name,score\n1,ann,3.5\n2,bob,4,extra\n3,eve,5\n")
    w("mixed.csv", """g,v,z
a,5,1
a,3.5,1
a,abc,1
a,,1
a,2024-01-01,1
b,zeta,1
b,alpha,1
b,10,1
b,Alpha,1
c,2024-05-05,1
c,2023-05-05,1
c,7,1
d,,1
d,,1
e,-1,1
e,1.5,1
e,-1.5,1
f,2,1
f,10,1
f,9.99,1
""")


P = "people.csv"
QUERIES = []


def q(sql, args=(), kind="fix", route=None, load=None):
    """route="any": the GPU planner may decline the shape at plan time (the reference then evaluates it); the
    result must be the reference's either way. load="gpu" / "reference": which csv_load the drop-in binary must have
    used for a statement that keeps the reference's evaluator (cq_dispatch.c: csv_load)."""
    QUERIES.append({"sql": sql, "args": list(args), "kind": kind, "route": route, "load": load})


def make_queries():
    # --- config 1 shape + basic aggregates ---
    q(f"SELECT role, COUNT(*), AVG(age) FROM '{P}' WHERE age > 25 GROUP BY role")
    q(f"SELECT role, COUNT(*), SUM(age), MIN(age), MAX(age) FROM '{P}' GROUP BY role")
    q(f"SELECT city, COUNT(*), AVG(height), MIN(height), MAX(height) FROM '{P}' GROUP BY city")
    q(f"SELECT COUNT(*) FROM '{P}'")
    q(f"SELECT COUNT(*), SUM(age), AVG(age), MIN(age), MAX(age) FROM '{P}'")
    q(f"SELECT COUNT(age), COUNT(city), COUNT(nosuch) FROM '{P}'")
    q(f"SELECT SUM(height), AVG(height), MIN(name), MAX(name) FROM '{P}'")
    q(f"SELECT MIN(joined), MAX(joined), MIN(email), MAX(city) FROM '{P}'")
    q(f"SELECT COUNT(*), SUM(age) FROM '{P}' WHERE age > 1000")
    q(f"SELECT role, city, COUNT(*), SUM(age) FROM '{P}' GROUP BY role, city")
    q(f"SELECT active, role, COUNT(*) FROM '{P}' WHERE age >= 25 GROUP BY active, role")
    q(f"SELECT joined, COUNT(*) FROM '{P}' GROUP BY joined")
    q(f"SELECT height, COUNT(*) FROM '{P}' GROUP BY height")
    q(f"SELECT role, COUNT(*) AS n, AVG(height) AS avg_height FROM '{P}' GROUP BY role")
    q(f"SELECT role, name, COUNT(*) FROM '{P}' GROUP BY role")
    q(f"SELECT nosuch, COUNT(*) FROM '{P}' GROUP BY nosuch")
    q(f"SELECT role, COUNT(*) FROM '{P}' GROUP BY role, nosuch")
    q(f"SELECT role, SUM(nosuch), MIN(nosuch) FROM '{P}' GROUP BY role")
    q(f"SELECT role, COUNT(*) FROM '{P}' GROUP BY role ORDER BY role")
    q(f"SELECT role, COUNT(*) AS n FROM '{P}' GROUP BY role HAVING n > 2")
    q(f"SELECT role, COUNT(*) FROM '{P}' WHERE city = 'Oslo' GROUP BY role LIMIT 1")
    # --- WHERE: comparisons and Q7 ---
    for cond in ["age > 30", "age >= 30", "age < 30", "age <= 30", "age = 30", "age != 30", "age <> 30", "30 < age",
                 "height > 170.5", "height = 175", "height = 175.0", "name = 'Ada'", "name > 'Fay'", "name <= 'Dana'",
                 "city = ''", "city != ''", "age = ''", "email = ''", "joined > '2021-01-01'", "joined = '2021-03-04'",
                 "joined = '20210304'", "joined < 20210304", "name = 15", "name >= 15", "name > 15", "name != 15",
                 "age = 'abc'", "age > 'abc'", "nosuch = 1", "nosuch < 1", "nosuch > 1", "nosuch = ''", "age > height",
                 "id = active", "'Ada' = name", "age > 25 AND role = 'user'", "age > 40 OR role = 'admin'",
                 "age > 25 AND role = 'user' OR city = 'Lima'", "NOT age > 30", "NOT (age > 30 AND active = 1)",
                 "age BETWEEN 25 AND 30", "height BETWEEN 160.5 AND 172.4", "name BETWEEN 'B' AND 'G'",
                 "joined BETWEEN '2020-01-01' AND '2021-12-31'", "age NOT BETWEEN 25 AND 30",
                 "role IN ('admin', 'guest')", "role NOT IN ('admin', 'guest')", "age IN (25, 30, 41)",
                 "age IN ('x', 30)", "age NOT IN (25, 'x')", "city IN ('', 'Oslo')", "height IN (175, 180.0)",
                 "age IN (25)", "age + 5 > 35", "age - 5 >= 25", "age * 2 = 60", "age / 2 > 15", "age / 0 > 1",
                 "age % 2 = 0", "age % 2 = 1", "height % 2 > 1", "age & 1 = 1", "age | 1 = 31", "age ^ 1 = 31",
                 "height & 1 = 1", "age * height > 5000", "-age < -40", "age + height / 2 > 120", "age + name > 1",
                 "(age + 1) * 2 > 70", "age * 2 / 2 = age", "height * 2 = 350", "age / 4 = 7.5", "age + 0.5 > 30",
                 "email LIKE '%@example.com'", "email LIKE '%gmail%'", "name LIKE 'A%'", "name LIKE '_da'",
                 "name LIKE '%a'", "name ILIKE 'a%'", "name LIKE 'ada'", "name ILIKE 'ADA'", "age LIKE '3%'",
                 "name LIKE '%'", "city LIKE ''", "email NOT LIKE '%gmail%'", "name LIKE 'J_d_'", "name LIKE '%%a%%'"]:
        q(f"SELECT COUNT(*) FROM '{P}' WHERE {cond}")
    # --- projection / select mode ---
    q(f"SELECT name, age FROM '{P}' WHERE age > 35")
    q(f"SELECT * FROM '{P}' WHERE role = 'admin'")
    q(f"SELECT name, nosuch, joined, height FROM '{P}' WHERE active = 0")
    q(f"SELECT name FROM '{P}' WHERE age > 20 LIMIT 3")
    q(f"SELECT name, age FROM '{P}' WHERE age > 20 ORDER BY age DESC LIMIT 4")
    q(f"SELECT name FROM '{P}' LIMIT 2 OFFSET 3")
    q(f"SELECT DISTINCT role FROM '{P}'")
    q(f"SELECT id, name FROM '{P}'")
    # --- typing quirks ---
    q("SELECT id, val FROM 'types.csv'")
    q("SELECT val, COUNT(*) FROM 'types.csv' GROUP BY val")
    q("SELECT tag, COUNT(*), SUM(val), AVG(val) FROM 'types.csv' GROUP BY tag")
    for cond in ["val > 0", "val < 0", "val = 4", "val = 4.0", "val = 2.5", "val >= 9223372036854775807", "val = ''",
                 "val = 'abc'", "val > 'a'", "val = '2024-01-15'", "val > '2024-01-01'", "val = 20240115",
                 "val LIKE '%e%'", "val LIKE '1%'", "val ILIKE 'HELLO%'", "val LIKE 'a\\%b'", "val LIKE 'a_b'",
                 "val LIKE '100%'", "val = 'NULL'", "val IN (1, 4, 'abc')", "val + 1 > 2", "val * 1 = val",
                 "val = 0.1", "val = 0.30000000000000004", "val = '  padded  '", "val = 'padded'", "val % 2 = 1",
                 "val = '12 3'", "val = 12", "val = 1.5"]:
        q(f"SELECT COUNT(*) FROM 'types.csv' WHERE {cond}")
        q(f"SELECT id FROM 'types.csv' WHERE {cond}")
    q("SELECT MIN(val), MAX(val) FROM 'types.csv' WHERE id < 10")
    q("SELECT MIN(val), MAX(val) FROM 'types.csv' WHERE id > 59 AND id < 67")
    q("SELECT MIN(val), MAX(val) FROM 'types.csv' WHERE id > 18 AND id < 29")
    # --- quoting ---
    q("SELECT id, name, role FROM 'quoted.csv'")
    q("SELECT role, COUNT(*) FROM 'quoted.csv' GROUP BY role")
    q("SELECT name, COUNT(*) FROM 'quoted.csv' GROUP BY name")
    for cond in ["role = 'admin'", "name LIKE '%x%'", "name LIKE '%,%'", "name = 'Last, First'", "name LIKE '%\"\"%'",
                 "name = 'spaced'", "name = 'quoted'", "name = ''", "role != 'user' AND name LIKE '%a%'", "id > 9"]:
        q(f"SELECT COUNT(*) FROM 'quoted.csv' WHERE {cond}")
        q(f"SELECT id, name FROM 'quoted.csv' WHERE {cond}")
    q("SELECT COUNT(*) FROM 'crlf.csv' WHERE age > 26")
    q("SELECT role, COUNT(*), SUM(age) FROM 'crlf.csv' GROUP BY role")
    q("SELECT * FROM 'crlf.csv'")
    q("SELECT role, COUNT(*), SUM(age) FROM 'blank_lines.csv' GROUP BY role")
    q("SELECT * FROM 'blank_lines.csv' WHERE age < 41")
    q("SELECT $1, COUNT(*), SUM($2) FROM 'noheader.csv' GROUP BY $1", args=["-n"])
    q("SELECT COUNT(*) FROM 'noheader.csv' WHERE $2 > 20", args=["-n"])
    q("SELECT * FROM 'noheader.csv' WHERE $0 >= 3", args=["-n"])
    q("SELECT name, score FROM 'semi.csv' WHERE score > 3.9", args=["-s", ";"])
    q("SELECT COUNT(*), SUM(score), MIN(name) FROM 'semi.csv'", args=["-s", ";"])
    q("SELECT name, score FROM 'tabs.csv' WHERE score > 3.9", args=["-s", "\t"])
    q("SELECT COUNT(*), SUM(score) FROM 'tabs.csv'", args=["-s", "\t"])
    # --- LIKE corpus ---
    for pat in ["A%", "%e", "%li%", "A_", "a%", "%@example.com", "USB-___", "USB%", "Al", "%", "_", "____", "%_%_%",
                "Ali\\_", "100%", "%%", "A%a", "%l%a%"]:
        for col in ["name", "email", "product"]:
            q(f"SELECT COUNT(*) FROM 'like.csv' WHERE {col} LIKE '{pat}'")
        q(f"SELECT COUNT(*) FROM 'like.csv' WHERE name ILIKE '{pat}'")
    q("SELECT COUNT(*) FROM 'like.csv' WHERE name LIKE 'A%' AND email LIKE '%.com'")
    q("SELECT COUNT(*) FROM 'like.csv' WHERE name LIKE 'A%' OR product LIKE 'USB%'")
    # --- GROUP BY on the generated table ---
    B = "big3k.csv"
    q(f"SELECT name, COUNT(*), SUM(age), AVG(height), MAX(height) FROM '{B}' GROUP BY name")
    q(f"SELECT name, surname, COUNT(*), SUM(age), MIN(height) FROM '{B}' WHERE age > 25 GROUP BY name, surname")
    q(f"SELECT age, gender, COUNT(*), AVG(height) FROM '{B}' GROUP BY age, gender")
    q(f"SELECT name, surname, age, height, COUNT(*) FROM '{B}' GROUP BY name, surname, age, height")
    q(f"SELECT SUM(age), MIN(height), MAX(height), AVG(height), COUNT(*) FROM '{B}' GROUP BY name, surname, age, height")
    q(f"SELECT dept, COUNT(*), SUM(bonus), AVG(bonus), MIN(bonus) FROM '{B}' GROUP BY dept")
    q(f"SELECT dept, MAX(bonus), MIN(age), MAX(age), COUNT(bonus) FROM '{B}' GROUP BY dept")
    q(f"SELECT height, COUNT(*), SUM(age) FROM '{B}' WHERE height > 1.5 GROUP BY height")
    q(f"SELECT bonus, COUNT(*) FROM '{B}' GROUP BY bonus")
    q(f"SELECT gender, dept, COUNT(*), SUM(bonus) FROM '{B}' WHERE age BETWEEN 20 AND 60 AND dept != 'ops' GROUP BY gender, dept")
    q(f"SELECT COUNT(*) FROM '{B}' WHERE age > 40")
    q(f"SELECT COUNT(*) FROM '{B}' WHERE height > 1.5")
    q(f"SELECT COUNT(*), SUM(age), SUM(height), AVG(bonus) FROM '{B}' WHERE gender = 'f'")
    q("SELECT k, COUNT(*), SUM(v) FROM 'keys.csv' GROUP BY k")
    q("SELECT d, k, COUNT(*) FROM 'keys.csv' GROUP BY d, k")
    q("SELECT d, COUNT(*), MIN(k), MAX(k) FROM 'keys.csv' GROUP BY d")
    q("SELECT g, MIN(v), MAX(v), SUM(v), COUNT(v) FROM 'mixed.csv' WHERE g != 'd' GROUP BY g")
    q("SELECT g, SUM(v), AVG(v), COUNT(v), COUNT(*) FROM 'mixed.csv' GROUP BY g")
    q("SELECT MIN(v), MAX(v) FROM 'mixed.csv'")
    q("SELECT MIN(v), MAX(v) FROM 'mixed.csv' WHERE g = 'b'")
    # --- joins (Q14) ---
    J = "FROM 'orders.csv' AS o JOIN 'customers.csv' AS c ON o.customer_id = c.id"
    q(f"SELECT COUNT(*) {J}")
    q(f"SELECT o.id, c.name, o.price {J}")
    q(f"SELECT * {J}")
    q(f"SELECT c.year, COUNT(*), SUM(o.price) {J} GROUP BY c.year")
    q(f"SELECT c.name, COUNT(*), SUM(o.quantity), MIN(o.price), MAX(o.price) {J} GROUP BY c.name")
    q(f"SELECT o.id, c.name {J} WHERE o.price > 60 AND c.year = 2023")
    q(f"SELECT COUNT(*), SUM(o.price), AVG(o.tax) {J} WHERE c.name LIKE 'C%'")
    q(f"SELECT o.id, c.email {J} LIMIT 3")
    # (ON with the right table's column first indexes the other table's column position into the row:
    #  out-of-bounds read in the reference, evaluator_joins.c:49-52 - kept out of the corpus)
    q("SELECT COUNT(*) FROM 'orders.csv' AS o JOIN 'customers.csv' AS c ON o.nosuch = c.id")
    q("SELECT COUNT(*) FROM 'orders.csv' JOIN 'customers.csv' ON customer_id = id")
    q("SELECT p.name, t.label FROM 'people.csv' AS p JOIN 'tags.csv' AS t ON p.role = t.code")
    q("SELECT t.label, COUNT(*), AVG(p.age) FROM 'people.csv' AS p JOIN 'tags.csv' AS t ON p.role = t.code GROUP BY t.label")
    q("SELECT p.name, t.label FROM 'people.csv' AS p JOIN 'tags.csv' AS t ON p.role = t.code WHERE p.age > 30 ORDER BY p.name")
    # --- LEFT / RIGHT / FULL joins (evaluator_joins.c:128-171): unmatched rows of either side, NULL-extended ---
    for jt in ("LEFT", "RIGHT", "FULL"):
        JO = f"FROM 'orders.csv' AS o {jt} JOIN 'customers.csv' AS c ON o.customer_id = c.id"
        q(f"SELECT o.id, o.customer_id, c.id, c.name {JO}")
        q(f"SELECT COUNT(*), COUNT(c.name), SUM(o.price), MIN(c.name), MAX(o.id) {JO}")
        q(f"SELECT c.name, COUNT(*), SUM(o.price) {JO} GROUP BY c.name")
        q(f"SELECT o.customer_id, c.email, COUNT(*) {JO} GROUP BY o.customer_id")
        q(f"SELECT o.id, c.name {JO} WHERE o.price > 20")
        q(f"SELECT o.id, c.name {JO} WHERE c.name LIKE 'C%' ORDER BY o.id")
        q(f"SELECT p.name, t.label FROM 'people.csv' AS p {jt} JOIN 'tags.csv' AS t ON p.role = t.code")
        q(f"SELECT t.label, COUNT(*), AVG(p.age) FROM 'people.csv' AS p {jt} JOIN 'tags.csv' AS t ON p.role = t.code GROUP BY t.label")
        q(f"SELECT COUNT(*) FROM 'orders.csv' AS o {jt} JOIN 'customers.csv' AS c ON o.nosuch = c.id")
    # --- shapes the GPU planner declines at plan time: they must come out of the drop-in binary all the same ---
    q(f"SELECT COUNT(*), SUM(age) FROM '{P}' WHERE age IN ({', '.join(str(k) for k in range(18, 48))})", route="any")
    q("SELECT COUNT(*), SUM(val) FROM 'alpha.csv' WHERE id > 1", args=["-s", "x"], route="any")
    # --- shapes that keep the reference's evaluator (scalar functions, CASE, STDDEV, sub-queries, expression GROUP BY):
    # their csv_load is the GPU-backed one of cq_dispatch.c and must hand back the reference's own table ---
    G = dict(route="reference", load="gpu")
    q(f"SELECT name, UPPER(role), LENGTH(email) FROM '{P}' WHERE age > 30", **G)
    q(f"SELECT name, CASE WHEN age > 30 THEN 'old' ELSE 'young' END FROM '{P}'", **G)
    q(f"SELECT role, STDDEV(age), STDDEV(height) FROM '{P}' GROUP BY role", **G)
    q(f"SELECT id, COALESCE(city, 'nowhere'), COALESCE(age, 0), joined FROM '{P}'", **G)
    q(f"SELECT name FROM '{P}' WHERE age > (SELECT AVG(age) FROM '{P}')", route="any", load="gpu")
    q(f"SELECT YEAR(joined), MONTH(joined), DAY(joined), name FROM '{P}' WHERE YEAR(joined) >= 2021", **G)
    q(f"SELECT CONCAT(name, '-', role), ROUND(height, 1), ABS(age - 40), height * 2 FROM '{P}'", **G)
    q("SELECT id, val, UPPER(tag), COALESCE(val, 'none') FROM 'types.csv'", **G)
    q("SELECT id, LOWER(tag), val FROM 'types.csv' WHERE val > 3", **G)
    q("SELECT id, UPPER(name), role, note FROM 'quoted.csv'", **G)
    q("SELECT id, LENGTH(name), LENGTH(note) FROM 'quoted.csv' WHERE role = 'admin'", **G)
    q("SELECT id, UPPER(role), age + 1 FROM 'crlf.csv'", **G)
    q("SELECT id, UPPER(role), age FROM 'blank_lines.csv'", **G)
    q("SELECT UPPER($1), $0, $2 FROM 'noheader.csv'", args=["-n"], **G)
    q("SELECT id, UPPER(name), score FROM 'semi.csv'", args=["-s", ";"], **G)
    q("SELECT id, UPPER(name), score FROM 'tabs.csv'", args=["-s", "\t"], **G)
    q("SELECT g, STDDEV(v), COUNT(*) FROM 'mixed.csv' GROUP BY g", **G)
    q("SELECT UPPER(k), v, d FROM 'keys.csv'", **G)
    q("SELECT id, LENGTH(id) FROM 'short.csv'", **G)
    q("SELECT id, LENGTH(id) FROM 'long.csv'", route="reference", load="reference")
    q("SELECT o.id, UPPER(c.name) FROM 'orders.csv' AS o JOIN 'customers.csv' AS c ON o.customer_id = c.id", **G)
    # --- reference's own fixtures, in place ---
    q("SELECT role, COUNT(*), AVG(age) FROM 'data/users.csv' WHERE age > 25 GROUP BY role", kind="refdata")
    q("SELECT COUNT(*) FROM 'data/test_data.csv'", kind="refdata")
    q("SELECT name, age FROM 'data/test_data.csv' WHERE age > 25", kind="refdata")
    q("SELECT role, AVG(height) AS avg_height FROM 'data/test_data.csv' GROUP BY role", kind="refdata")
    q("SELECT COUNT(*) FROM 'data/users.csv' WHERE age BETWEEN 25 AND 35", kind="refdata")
    q("SELECT COUNT(*), SUM(price), AVG(tax) FROM 'data/orders.csv' WHERE quantity > 1", kind="refdata")


def run(dump, item, env=None):
    cwd = FIX if item["kind"] == "fix" else REF
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([dump] + item["args"] + [item["sql"]], cwd=cwd, capture_output=True, timeout=120, env=e)
    return p.returncode, p.stdout.decode("latin1"), p.stderr.decode("latin1")


def main():
    make_fixtures()
    make_queries()
    if not os.path.exists(REF_DUMP):
        sys.exit("oracle/_ref/ref_dump missing: run `make -C oracle ref`")
    out = []
    for i, item in enumerate(QUERIES):
        rc, text, _ = run(REF_DUMP, item)
        # which statements the planner (cq_dispatch.c) puts on the accelerated path
        _, _, err = run(ORACLE_DUMP, item, env={"CQ_GPU_TRACE": "1"})
        route = item.get("route") or ("gpu" if "route=gpu" in err else "reference")
        out.append({"id": i, "kind": item["kind"], "args": item["args"], "sql": item["sql"], "rc": rc, "route": route,
                    "expected": text})
        if item.get("load"):
            out[-1]["load"] = item["load"]
    with open(os.path.join(ROOT, "tests", "golden", "sql_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(f"{len(out)} golden cases written")


if __name__ == "__main__":
    main()
