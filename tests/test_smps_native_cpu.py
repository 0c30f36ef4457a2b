"""The C++ SMPS reader of the library (``sqlp_smps_*``; host-only, no GPU needed)
against (i) the reference's own known answers for lands (``test/smps_tests.jl``), (ii) the independent
numpy reader ``sqlp_b200.smps`` on every shipped instance, (iii) files written on the fly whose tables are
known by construction -- those travel to the GPU box, where ``/root/reference`` does not exist."""
import ctypes as C
import os

import numpy as np
import pytest

from sqlp_b200 import _lib, smps
from tests.helpers import write_smps

REF = "/root/reference/spInput"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference data not mounted")


def _paths(name):
    d = os.path.join(REF, name)
    return [os.path.join(d, f"{name}.{ext}") for ext in ("cor", "tim", "sto")]


@needs_ref
def test_native_reader_reproduces_reference_smps_tests():            # test/smps_tests.jl:4-58
    n = smps.NativeSmps(*_paths("lands"))
    cor = n.cor()
    assert cor.directions == list("NGLLLLLGGG")
    assert cor.row_names == ["OBJ", "S1C1", "S1C2", "S2C1", "S2C2", "S2C3", "S2C4", "S2C5", "S2C6", "S2C7"]
    assert cor.col_names == ["X1", "X2", "X3", "X4", "Y11", "Y21", "Y31", "Y41", "Y12", "Y22", "Y32", "Y42",
                             "Y13", "Y23", "Y33", "Y43"]
    assert sum(1 for v in cor.entries.values() if v != 0) == 52
    assert list(cor.rhs) == [0., 12, 120, 0, 0, 0, 0, 0, 3, 2]
    assert (cor.lower == 0).all() and np.isinf(cor.upper).all()
    tim = n.tim()
    assert tim.name == "LandS" and tim.periods == [("TIME1", "X1", "OBJ"), ("TIME2", "Y11", "S2C1")]
    sto = n.sto()
    assert sto.name == "LandS" and sto.positions == [("RHS", "S2C5")]
    assert sto.params[0] == ([3.0, 5.0, 7.0], [0.3, 0.4, 0.3])
    d = n.dims
    assert (d["n1"], d["n2"], d["m2"]) == (4, 12, 7)                 # sp2: 16 variables, 7 constraints
    st = n.stage2()
    assert list(st.pos_row) == [4] and list(st.pos_col) == [-1]
    assert list(st.rbar) == [0, 0, 0, 0, 0, 3, 2]


@needs_ref
@pytest.mark.parametrize("name,dims", [("lands", (4, 7, 1)), ("baa99-20", (20, 40, 20)), ("ssn", (89, 175, 86)),
                                       ("storm", (121, 528, 117)), ("transship", (7, 35, 7))])
def test_native_reader_equals_numpy_reader_on_shipped_instances(name, dims):
    n = smps.NativeSmps(*_paths(name))
    cor, tim, sto = (f(p) for f, p in zip((smps.read_cor, smps.read_tim, smps.read_sto), _paths(name)))
    st = smps.stage2_tables(cor, tim, sto)
    nc, ns, nsto = n.cor(), n.stage2(), n.sto()
    assert (nc.name, nc.directions, nc.row_names, nc.col_names) == (cor.name, cor.directions, cor.row_names, cor.col_names)
    assert nc.entries == cor.entries
    assert np.array_equal(nc.rhs, cor.rhs) and np.array_equal(nc.lower, cor.lower) and np.array_equal(nc.upper, cor.upper)
    assert n.tim() == tim
    assert (nsto.name, nsto.positions, nsto.kind, nsto.params) == (sto.name, sto.positions, sto.kind, sto.params)
    for f in ("rbar", "T_colptr", "T_rowval", "T_nzval", "W", "cost", "y_lower", "y_upper", "pos_row", "pos_col",
              "x_lower", "x_upper", "x_cost"):
        assert np.array_equal(getattr(ns, f), getattr(st, f)), f
    assert (ns.n1, ns.m2, len(ns.pos_row)) == dims


@needs_ref
def test_native_tables_equal_the_committed_fixtures():
    from tests.helpers import load_instance
    for name in ("lands", "baa99-20", "ssn", "storm"):
        st = smps.NativeSmps(*_paths(name)).stage2()
        P, z = load_instance(name)
        assert np.array_equal(st.rbar, P.rbar) and np.array_equal(st.T_nzval, P.T_nzval)
        assert np.array_equal(st.T_rowval, P.T_rowval) and np.array_equal(st.pos_row, P.pos_row)


@pytest.mark.parametrize("continuous,seed", [(False, 11), (True, 12), (False, 13)])
def test_native_reader_on_generated_files(tmp_path, continuous, seed):
    paths, ex = write_smps(str(tmp_path), seed=seed, continuous=continuous, n1=6, n2=8, m1=3, m2=9,
                           n_rhs_elems=4, n_T_elems=3)
    n = smps.NativeSmps(paths["cor"], paths["tim"], paths["sto"])
    cor = n.cor()
    assert cor.name == ex["name"] and cor.row_names == ex["rows"] and cor.col_names == ex["cols"]
    assert cor.directions == ex["dirs"]
    M = np.zeros_like(ex["M"])
    for (i, j), v in cor.entries.items():
        M[i, j] = v
    assert np.array_equal(M, ex["M"])                                 # overwritten entry, %E numbers
    assert ex["explicit_zero"] in cor.entries and cor.entries[ex["explicit_zero"]] == 0.0
    assert np.array_equal(cor.rhs, ex["rhs"])
    assert np.array_equal(cor.lower, ex["lower"]) and np.array_equal(cor.upper, ex["upper"])
    st = n.stage2()
    assert (st.n1, st.n2, st.m2) == (ex["n1"], ex["n2"], ex["m2"])
    assert np.array_equal(st.T_dense(), ex["T"]) and np.array_equal(st.W, ex["W"])
    assert len(st.T_nzval) == np.count_nonzero(ex["T"])               # the explicit zero is not stored
    for j in range(st.n1):                                            # rows ascending within a column
        assert (np.diff(st.T_rowval[st.T_colptr[j]:st.T_colptr[j + 1]]) > 0).all()
    assert np.array_equal(st.rbar, ex["rbar"]) and np.array_equal(st.cost, ex["cost"])
    assert np.array_equal(st.x_cost, ex["x_cost"])
    assert np.array_equal(st.pos_row, ex["pos_row"]) and np.array_equal(st.pos_col, ex["pos_col"])
    sto = n.sto()
    assert sto.positions == ex["positions"] and sto.kind == ex["kinds"]
    for e, k in enumerate(ex["kinds"]):
        assert sto.params[e] == (ex["tables"][e] if k == "DISCRETE" else ex["pars"][e])
    # and the numpy reader agrees
    st2 = smps.stage2_tables(smps.read_cor(paths["cor"]), smps.read_tim(paths["tim"]), smps.read_sto(paths["sto"]))
    for f in ("rbar", "T_colptr", "T_rowval", "T_nzval", "W", "cost", "pos_row", "pos_col", "y_lower", "y_upper"):
        assert np.array_equal(getattr(st, f), getattr(st2, f)), f


def test_native_reader_without_a_sto_file(tmp_path):
    paths, ex = write_smps(str(tmp_path), seed=5)
    n = smps.NativeSmps(paths["cor"], paths["tim"], None)
    assert n.dims["s"] == 0 and n.dims["max_outcomes"] == 0 and n.dims["m2"] == ex["m2"]


def _load(cor, tim, sto):
    h = C.c_void_p()
    st = _lib.lib().sqlp_smps_load(cor.encode(), tim.encode(), sto.encode() if sto else None, C.byref(h))
    if st == 0:
        _lib.lib().sqlp_smps_destroy(h)
    return st, _lib.lib().sqlp_last_error().decode()


def _edit(path, old, new, out):
    text = open(path).read()
    assert old in text, old
    with open(out, "w") as fh:
        fh.write(text.replace(old, new, 1))
    return out


def test_native_reader_error_behaviour(tmp_path):
    """Where the reference asserts or errors, the library returns a status and a message naming file:line."""
    paths, ex = write_smps(str(tmp_path), seed=7)
    cor, tim, sto = paths["cor"], paths["tim"], paths["sto"]
    bad = str(tmp_path / "bad")
    assert _load(cor, tim, sto)[0] == _lib.OK
    st, msg = _load(str(tmp_path / "nope.cor"), tim, sto)
    assert st == _lib.E_IO and "nope.cor" in msg
    # unsupported section (smps_cor.jl:44 asserts)
    st, msg = _load(_edit(cor, "RHS\n", "RANGES\n    RNG  A0  4.0\nRHS\n", bad), tim, sto)
    assert st == _lib.E_INVALID and "RANGES" in msg and "bad:" in msg
    # first row must be the objective (smps_cor.jl:178-179)
    assert "objective" in _load(_edit(cor, " N  OBJ", " G  OBJ", bad), tim, sto)[1]
    # unsupported bound type (smps_cor.jl:131-132)
    st, msg = _load(_edit(cor, " UP BND", " BV BND", bad), tim, sto)
    assert st == _lib.E_INVALID and "BV" in msg
    # a number that does not parse
    assert "not a number" in _load(_edit(cor, "123.456", "12x.456", bad), tim, sto)[1]
    # unknown row in COLUMNS (KeyError in the reference)
    assert "unknown row" in _load(_edit(cor, "123.456", "123.456   NOROW   1.0", bad), tim, sto)[1]
    # INDEP with a second keyword, unknown keyword, BLOCKS section (smps_sto.jl:69-71, 97, 104)
    assert "one keyword" in _load(cor, tim, _edit(sto, "INDEP         DISCRETE", "INDEP         DISCRETE ADD", bad))[1]
    assert "GAMMA" in _load(cor, tim, _edit(sto, "INDEP         DISCRETE", "INDEP         GAMMA", bad))[1]
    assert "BLOCKS" in _load(cor, tim, _edit(sto, "INDEP         DISCRETE", "BLOCKS        DISCRETE", bad))[1]
    # a random element outside stage 2, unknown names
    assert "not in stage 2" in _load(cor, tim, _edit(sto, f"RHS  S2R{ex['pos_row'][0]}", "RHS  A0", bad))[1]
    assert "unknown row" in _load(cor, tim, _edit(sto, f"RHS  S2R{ex['pos_row'][0]}", "RHS  ZZZ", bad))[1]
    # three periods: two-stage problems only
    st, msg = _load(cor, _edit(tim, "ENDATA", "    Y3  S2R4  TIME3\nENDATA", bad), sto)
    assert st == _lib.E_UNSUPPORTED and "two-stage" in msg
    # null path
    h = C.c_void_p()
    assert _lib.lib().sqlp_smps_load(None, tim.encode(), None, C.byref(h)) == _lib.E_INVALID


def test_name_accessor_bounds(tmp_path):
    paths, ex = write_smps(str(tmp_path), seed=3)
    n = smps.NativeSmps(paths["cor"], paths["tim"], paths["sto"])
    L = _lib.lib()
    buf = C.create_string_buffer(64)
    assert L.sqlp_smps_name(n._h, 3, len(ex["rows"]), buf, 64) == _lib.E_RANGE
    assert L.sqlp_smps_name(n._h, 3, 0, buf, 2) == _lib.E_RANGE       # "OBJ" does not fit in 2 bytes
    assert L.sqlp_smps_name(n._h, 99, 0, buf, 64) == _lib.E_INVALID
    assert L.sqlp_smps_name(n._h, 3, 0, buf, 64) == 0 and buf.value == b"OBJ"


def test_smps_reader_from_plain_c(tmp_path):
    """tests/abi_smps_demo.c: gcc against include/sqlp_b200.h, linked to the library, no Python and no GPU in
    the loop -- the tables it prints are the ones the files were generated from."""
    import json
    import subprocess
    paths, ex = write_smps(str(tmp_path), seed=17, continuous=True, n1=5, n2=7, m1=2, m2=8, n_rhs_elems=3, n_T_elems=2)
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "abi_smps_demo")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.run(["gcc", "-O1", "-o", exe, os.path.join(here, "abi_smps_demo.c"), "-L", libdir, "-lsqlp_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe, paths["cor"], paths["tim"], paths["sto"]], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout)
    assert (r["name"], r["n1"], r["n2"], r["m2"], r["s"]) == (ex["name"], ex["n1"], ex["n2"], ex["m2"], len(ex["positions"]))
    assert r["rbar"] == list(ex["rbar"])
    T = np.zeros_like(ex["T"])
    for i, j, v in r["T"]:
        T[i, j] = v
    assert np.array_equal(T, ex["T"]) and r["T_nnz"] == np.count_nonzero(ex["T"])
    for e, el in enumerate(r["elements"]):
        assert (el["col"], el["row"]) == ex["positions"][e]
        assert el["pos"] == [int(ex["pos_row"][e]), int(ex["pos_col"][e])]
        assert el["kind"] == {"DISCRETE": 0, "NORMAL": 1, "UNIFORM": 2}[ex["kinds"][e]]
        if ex["kinds"][e] == "DISCRETE":
            assert [o[0] for o in el["outcomes"]] == ex["tables"][e][0]
            assert [o[1] for o in el["outcomes"]] == ex["tables"][e][1]


@needs_ref
@pytest.mark.parametrize("name", ["lands", "baa99-20"])
def test_full_tables_from_the_native_reader_equal_the_committed_fixtures(name):
    """``smps.full_tables`` over the native reader (what ``tools/run_sd.py --smps`` runs on) reproduces the
    ``<name>_full.npz`` fixtures the SD-loop tests use."""
    from tests.helpers import load_full_instance
    n = smps.NativeSmps(*_paths(name))
    got = smps.full_tables(n.cor(), n.stage2(), n.sto())
    ref = load_full_instance(name)
    for key, val in got.items():
        assert np.array_equal(np.asarray(val), ref[key]), key


@pytest.mark.parametrize("name", ["lands", "baa99-20"])
def test_real_instances_round_trip_through_smps_text(tmp_path, name):
    """A committed full fixture written back as SMPS text and read by the native reader gives the fixture's
    tables again -- real instance data for the file-based paths where the reference checkout is absent."""
    from tests.helpers import load_full_instance, write_smps_from_full
    zf = load_full_instance(name)
    prefix = write_smps_from_full(zf, str(tmp_path), name)
    n = smps.NativeSmps(prefix + ".cor", prefix + ".tim", prefix + ".sto")
    got = smps.full_tables(n.cor(), n.stage2(), n.sto())
    for key, val in got.items():
        if key == "out_cdf":          # probabilities went through text as differences of the cdf
            assert np.allclose(np.asarray(val), zf[key], rtol=0, atol=1e-15), key
        else:
            assert np.array_equal(np.asarray(val), zf[key]), key
