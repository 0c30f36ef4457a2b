"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header
declares (no compute calls without a GPU); the product fails loudly without a device; host-side
logic of the Python mirror; the SMPS reader against the reference's test/smps_tests.jl."""
import ctypes as C
import os

import numpy as np
import pytest

from sqlp_b200 import _lib, smps


def test_library_builds_loads_and_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sqlp_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib._RESTYPE)
    assert bound == set(declared), bound ^ set(declared)
    assert b"sm_100a" in L.sqlp_version()


def test_library_is_native_sm100a_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    h = C.c_void_p()
    assert L.sqlp_ctx_create(0, C.byref(h)) == _lib.E_CUDA
    assert b"no CPU fallback" in L.sqlp_last_error()
    from sqlp_b200 import twosd as T
    with pytest.raises(T.SqlpError):
        T.sdDualVertexSet([[1.0, 2.0, 3.0]])


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "sqlp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "sqlp_oracle" not in text, f


def test_coefficients_position_table_and_scenario_flattening():
    from sqlp_b200 import twosd as T
    rows = {f"S2C{i + 1}": i for i in range(7)}
    cols = {f"X{i + 1}": i for i in range(4)}
    coef = T.sdSubprobCoefficients(np.zeros(7), np.arange(5), np.arange(4), -np.ones(4), 4, rows, cols,
                                   [("RHS", "S2C5"), ("X2", "S2C2"), ("rhs", "S2C7")])
    pr, pc = coef.positions()
    assert list(pr) == [4, 1, 6] and list(pc) == [-1, 1, -1]
    sc = [(("rhs", "S2C7"), 2.5), (("RHS", "S2C5"), 5.0), (("X2", "S2C2"), -1.5)]
    assert list(coef.scenario_values(sc)) == [5.0, -1.5, 2.5]
    assert list(coef.scenario_values([1.0, 2.0, 3.0])) == [1.0, 2.0, 3.0]
    with pytest.raises(KeyError):
        coef.scenario_values([(("RHS", "NOPE"), 1.0)])
    with pytest.raises(ValueError):
        coef.scenario_values([(("RHS", "S2C5"), 1.0)])          # two elements not realised
    bad = T.sdSubprobCoefficients(np.zeros(7), np.arange(5), np.arange(4), -np.ones(4), 4, rows, cols,
                                  [("Y11", "S2C5")])
    with pytest.raises(KeyError):
        bad.positions()


REF = "/root/reference/spInput/lands"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference data not mounted")
def test_smps_reader_matches_reference_smps_tests():      # test/smps_tests.jl:4-58
    cor = smps.read_cor(os.path.join(REF, "lands.cor"))
    assert cor.directions == list("NGLLLLLGGG")
    assert cor.row_names == ["OBJ", "S1C1", "S1C2", "S2C1", "S2C2", "S2C3", "S2C4", "S2C5", "S2C6", "S2C7"]
    assert cor.col_names == ["X1", "X2", "X3", "X4", "Y11", "Y21", "Y31", "Y41", "Y12", "Y22", "Y32",
                             "Y42", "Y13", "Y23", "Y33", "Y43"]
    assert sum(1 for v in cor.entries.values() if v != 0) == 52
    assert list(cor.rhs) == [0., 12, 120, 0, 0, 0, 0, 0, 3, 2]
    assert (cor.lower == 0).all() and np.isinf(cor.upper).all()
    tim = smps.read_tim(os.path.join(REF, "lands.tim"))
    assert tim.name == "LandS"
    assert tim.periods == [("TIME1", "X1", "OBJ"), ("TIME2", "Y11", "S2C1")]
    sto = smps.read_sto(os.path.join(REF, "lands.sto"))
    assert sto.name == "LandS" and sto.positions == [("RHS", "S2C5")]
    assert sto.params[0] == ([3.0, 5.0, 7.0], [0.3, 0.4, 0.3])
    st = smps.stage2_tables(cor, tim, sto)
    assert (st.n1, st.n2, st.m2) == (4, 12, 7)               # 4 + 12 = 16 vars, 7 constraints
    assert list(st.pos_row) == [4] and list(st.pos_col) == [-1]
    u = np.array([[0.0], [0.29], [0.3], [0.69], [0.7], [0.999]])
    assert list(smps.sample_values(sto, u)[:, 0]) == [3.0, 3.0, 5.0, 5.0, 7.0, 7.0]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference data not mounted")
def test_committed_fixtures_match_the_reference_data():
    """tests/golden/instances/*.npz were generated from spInput by tools/make_golden.py."""
    from tests.helpers import load_instance
    for name, dims in {"lands": (4, 7, 1), "baa99-20": (20, 40, 20), "ssn": (89, 175, 86),
                       "storm": (121, 528, 117)}.items():
        d = os.path.join(os.path.dirname(REF), name)
        cor = smps.read_cor(os.path.join(d, f"{name}.cor"))
        st = smps.stage2_tables(cor, smps.read_tim(os.path.join(d, f"{name}.tim")),
                                smps.read_sto(os.path.join(d, f"{name}.sto")))
        P, z = load_instance(name)
        assert (st.n1, st.m2, len(st.pos_row)) == dims == (P.n1, P.m2, P.s)
        assert np.array_equal(st.rbar, P.rbar) and np.array_equal(st.T_nzval, P.T_nzval)
        assert np.array_equal(st.pos_row, P.pos_row)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours) prints one JSON line with
    the agreed keys; a tiny workload keeps this in seconds."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(_lib.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--vertices", "128", "--scen-per-gpu", "2000", "--epigraphs", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "evals/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def test_bench_roofline_entries_from_a_recorded_profile():
    """bench.py's roofline arithmetic on a recorded device profile (no GPU): achieved = work / event time of the
    kernel class, the fraction against the sustained cuBLAS figure (and the burst one beside it), and the DRAM traffic
    of the captured shape only -- per pool, and not for the strong-scaled leg, whose shape the capture does not have."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(_lib.ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    leg = {"prof": {"contract": [0.0, 0, 0.0], "screen": [15.4, 20, 1.4752e13], "resolve": [5.9, 20, 1.0e7],
                    "fallback": [0.14, 20, 0.0]}, "ms": 24.55}
    args = types.SimpleNamespace(instance="storm", vertices=16384, scen_per_gpu=1_000_000, epigraphs=4, pool="real")
    r = b.rooflines(leg, args)
    d = r["dominant"]
    assert d["kernel"].startswith("k_screen") and d["bound"] == "tensor" and d["unit"] == "TFLOP/s"
    assert abs(d["achieved"] - 1.4752e13 / 15.4e-3 * 1e-12) < 1e-6 and abs(d["frac"] - d["achieved"] / d["peak"]) < 1e-12
    assert abs(d["share_of_step"] - 15.4 / 24.55) < 1e-12 and d["launches"] == 20
    assert d["traffic"] and "r02_step_kernels" in d["traffic_source"]
    if "frac_of_burst" in d:
        assert d["frac_of_burst"] < d["frac"]
    assert any(k["kernel"] == "k_screen_resolve" for k in r["all"])
    assert b.rooflines(leg, None, "synthetic")["dominant"]["traffic"] != d["traffic"]
    assert b.rooflines(leg, None, "-")["dominant"]["traffic"] is None            # the strong-scaled leg
    args.scen_per_gpu = 125_000
    assert b.rooflines(leg, args)["dominant"]["traffic"] is None                  # another shape
