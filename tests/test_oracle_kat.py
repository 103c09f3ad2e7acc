"""Pins the CPU oracle against the hand-derived known-answer vector of SURVEY.md section 8c
(the reference ships no tests of its own)."""
import numpy as np

from oracle import oracle as orc

PTS = np.array([[0.3, 0.3, 0.3], [0.9, 0.9, 0.9], [-0.9, 0.2, 0.6]], dtype=np.float32)
RGB = np.array([[200, 100, 50], [10, 20, 30], [255, 255, 255]], dtype=np.uint8)


def test_keys_octal():
    keys = orc.compute_keys(PTS, (0, 0, 0), 1.0, 2)
    assert [oct(int(k)) for k in keys] == ["0o170", "0o177", "0o164"]


def test_invalid_key_q1():
    pts = np.array([[np.inf, 0, 0], [0, np.inf, 0], [0, 0, np.nan], [0.1, np.nan, 0.1]], dtype=np.float32)
    keys = orc.compute_keys(pts, (0, 0, 0), 1.0, 2)
    assert keys[0] == 1 and keys[2] == 1
    assert keys[1] != 1 and keys[3] != 1  # Q1: y is never tested


def test_insert_once_and_twice():
    t = orc.OracleSVO((0, 0, 0), 1.0, 2)
    t.integrate_points(PTS, RGB)
    assert t.size == 24
    p = t.pool()
    w0 = p[0::2]
    w1 = p[1::2]
    assert w0[6] == 0x40000008 and w0[7] == 0x40000010
    assert w1[12] == 0x81808080
    assert w1[16] == 0x81193264
    assert w1[23] == 0x810F0A05
    assert w1[6] == 0x81101010
    assert w1[7] == 0x8105070D
    assert w1[0] == 0x81020203  # Q6 canonical
    for i in range(1, 6):
        assert w0[i] == 0 and w1[i] == 0
    untouched = [i for i in range(8, 24) if i not in (12, 16, 23)]
    for i in untouched:
        assert w0[i] == 0 and w1[i] == 0x7F000000
    c = t.counters()
    assert c.n_unique == 3 and c.n_split == 2
    # second insert of the same cloud: Q3 splits leaf 23 (key 0o177, last digit 7)
    t.integrate_points(PTS, RGB)
    assert t.size == 32
    assert t.pool()[2 * 23] == 0x40000018


def test_no_q3_without_quirks():
    t = orc.OracleSVO((0, 0, 0), 1.0, 2, quirks=False)
    t.integrate_points(PTS, RGB)
    t.integrate_points(PTS, RGB)
    assert t.size == 24
