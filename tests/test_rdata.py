"""The .rda reader (bayesianvars_b200/rdata.py): R's XDR serialisation, formats 2 and 3."""
import bz2
import struct
from pathlib import Path

import numpy as np
import pytest

from bayesianvars_b200.rdata import RMatrix, read_rda

GOLD = Path(__file__).resolve().parent / "golden" / "usmacro_growth.npz"
REF = Path("/root/reference/data/usmacro_growth.rda")


def _i(*v):
    return b"".join(struct.pack(">i", x) for x in v)


def _chr(s):
    return _i(9 | (1 << 18), len(s)) + s.encode()


def _sym(s):
    return _i(1) + _chr(s)


def hand_built_rda(version=2):
    """save(list = c("m", "k")) of m = matrix(c(1.5, -2, 3, 4.25), 2, dimnames = list(NULL, c("a", "b"))), k = 7L."""
    body = _i(version, 0x040201, 0x020300)
    if version == 3:
        body += _i(5) + b"UTF-8"
    dimnames = _i(19, 2) + _i(254) + _i(16, 2) + _chr("a") + _chr("b")
    attrs = (_i(2 | (1 << 10)) + _sym("dim") + _i(13, 2, 2, 2) +
             _i(2 | (1 << 10)) + _sym("dimnames") + dimnames + _i(254))
    m = _i(14 | (1 << 9), 4) + struct.pack(">4d", 1.5, -2.0, 3.0, 4.25) + attrs
    k = _i(13, 1, 7)
    top = _i(2 | (1 << 10)) + _sym("m") + m + _i(2 | (1 << 10)) + _sym("k") + k + _i(254)
    return b"RDX%d\nX\n" % version + body + top


@pytest.mark.parametrize("version", [2, 3])
@pytest.mark.parametrize("compress", [False, True])
def test_hand_built_stream(tmp_path, version, compress):
    raw = hand_built_rda(version)
    p = tmp_path / "x.rda"
    p.write_bytes(bz2.compress(raw) if compress else raw)
    o = read_rda(p)
    assert set(o) == {"m", "k"} and isinstance(o["m"], RMatrix)
    assert np.array_equal(o["m"].values, np.array([[1.5, 3.0], [-2.0, 4.25]]))
    assert o["m"].colnames == ["a", "b"] and o["m"].rownames is None
    assert np.array_equal(o["k"], [7])
    assert np.array_equal(o["m"].columns(["b"]), [[3.0], [4.25]])


def test_rejects_other_files(tmp_path):
    p = tmp_path / "bad.rda"
    p.write_bytes(b"not an R file")
    with pytest.raises(ValueError, match="magic"):
        read_rda(p)
    p.write_bytes(hand_built_rda()[:40])
    with pytest.raises(ValueError, match="truncated"):
        read_rda(p)


def test_usmacro_fixture_is_the_reference_dataset():
    """data/usmacro_growth.rda (R/data.R: 247 quarters x 21 series) decodes to the committed fixture."""
    g = np.load(GOLD)
    assert g["values"].shape == (247, 21) and list(g["colnames"][:3]) == ["GDPC1", "PCECC96", "GPDIC1"]
    assert g["rownames"][0] == "1959-12-01"
    if not REF.exists():
        pytest.skip("reference checkout not present on this machine")
    m = read_rda(REF)["usmacro_growth"]
    assert np.array_equal(m.values, g["values"]) and m.colnames == list(g["colnames"])
