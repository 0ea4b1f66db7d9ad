"""The seeded restatement of utils/generate_big_dataset.py: shape of the rows."""
import re

from oracle_lib import generate_bigdata


def test_generator_shape():
    data = generate_bigdata(5000, seed=1)
    lines = data.decode().split("\n")
    assert lines[0] == "name,surname,age,gender,height"
    assert lines[-1] == ""
    pat = re.compile(r"^([A-P])\1{9},([A-P])\2{7},([1-8][0-9]),[fm],(1\.\d{1,2}|2\.0)$")
    for ln in lines[1:-1]:
        m = pat.match(ln)
        assert m, ln
        assert 10 <= int(m.group(3)) <= 80
    mean_len = len(data) / 5000
    assert 29.5 < mean_len < 30.3  # the reference generator measures 29.89 B/row


def test_generator_is_row_independent():
    a = generate_bigdata(100, seed=7)
    b = generate_bigdata(200, seed=7)
    assert b.startswith(a)
    assert generate_bigdata(100, seed=8) != a


def test_generator_key_column():
    data = generate_bigdata(1000, seed=3, key_card=50).decode().split("\n")
    assert data[0].endswith(",uid")
    uids = {int(ln.rsplit(",", 1)[1]) for ln in data[1:-1]}
    assert uids <= set(range(50)) and len(uids) > 40
