"""The instrumented op-counter of the oracle reproduces the algorithmic flop constants of
SURVEY 8(d) (the denominators of bench.py's roofline) within +-10 %."""
import re
from pathlib import Path

import pytest

from oracle import opcount

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def counts():
    return opcount.configs()


@pytest.mark.parametrize("name", sorted(opcount.SURVEY_CONSTANTS))
def test_counts_match_survey_constants(counts, name):
    for key, ref in opcount.SURVEY_CONSTANTS[name].items():
        got = counts[name][key]
        assert abs(got / ref - 1.0) <= 0.10, (name, key, got, ref)


def test_rhs_costs(counts):
    # Lorenz c_f = 8, c_J = 11; VdP c_f = 5, c_J = 7 (+1 here: d(x^2) counted as a full
    # product); reduced-1 HH: 13 exp per compartment (hodgkin_huxley.py:12-51)
    assert (counts["lorenz"]["c_f"], counts["lorenz"]["c_J"]) == (8, 11)
    assert counts["van_der_pol"]["c_f"] == 5 and counts["van_der_pol"]["c_J"] in (7, 8)
    assert counts["hh_1comp_r1"]["exp_per_rhs"] == 13 and counts["hh_2comp_r1"]["exp_per_rhs"] == 26


def test_bench_uses_the_checked_constants():
    src = (ROOT / "bench.py").read_text()
    m = re.search(r'F_STEP = \{"Lorenz": ([0-9.]+), "VanDerPol": ([0-9.]+)\}', src)
    assert m and float(m.group(1)) == opcount.SURVEY_CONSTANTS["lorenz"]["step"]
    assert float(m.group(2)) == opcount.SURVEY_CONSTANTS["van_der_pol"]["step"]
    m = re.search(r'F_STEP_PREDICT = \{"Lorenz": ([0-9.]+), "VanDerPol": ([0-9.]+)\}', src)
    assert m and float(m.group(1)) == opcount.SURVEY_CONSTANTS["lorenz"]["predict"]
    c3 = (ROOT / "tools" / "bench_c3.py").read_text()
    assert "40.7e3, 0.99e6" in c3 and "10.0e3, 0.18e6" in c3
