"""Host helpers of the GridMapping mirror against the expectations of the reference's
tests/gridmapping/test_helpers.py: resolution rounding (round_to_fraction decides the xy_res a
grid mapping derives from coordinates, hence every tile box and window downstream), near-integer
normalisation, number pairs and the longitude conventions."""

from fractions import Fraction

import numpy as np
import pytest

from xcube_resampling_b200.gridmapping import (_normalize_number_pair, _to_int_or_float, from_lon_360,
                                               round_to_fraction, to_lon_360)

VALUES = [-1.0, 0.0, 5.247476065426347e-09, 3.427467229408875e-06, 4.501758583626108e-06, 1.1351705264714663e-05,
          0.00048171747406886744, 0.0018032657496927416, 0.0019897341919324425, 0.0041643509375105065,
          0.030607346091352187, 1.0076973439575128, 1.0, 84.54360269093455, 494.86581234602096, 987.9441243998718,
          1757.368043916636, 1143506.2928512183, 217971970.75235566]

# expected fractions per (digits, resolution), test_helpers.py:93-169
EXPECTED = {
    (2, 1): [Fraction(-1, 1), Fraction(0, 1), Fraction(13, 2500000000), Fraction(17, 5000000), Fraction(9, 2000000),
             Fraction(11, 1000000), Fraction(3, 6250), Fraction(9, 5000), Fraction(1, 500), Fraction(21, 5000),
             Fraction(31, 1000), Fraction(1, 1), Fraction(1, 1), Fraction(85, 1), Fraction(490, 1), Fraction(990, 1),
             Fraction(1800, 1), Fraction(1100000, 1), Fraction(220000000, 1)],
    (3, 0.25): [Fraction(-1, 1), Fraction(0, 1), Fraction(2099, 400000000000), Fraction(1371, 400000000),
                Fraction(1801, 400000000), Fraction(227, 20000000), Fraction(1927, 4000000), Fraction(721, 400000),
                Fraction(199, 100000), Fraction(833, 200000), Fraction(153, 5000), Fraction(403, 400), Fraction(1, 1),
                Fraction(1691, 20), Fraction(1979, 4), Fraction(988, 1), Fraction(3515, 2), Fraction(1142500, 1),
                Fraction(218000000, 1)],
    (2, 0.5): [Fraction(-1, 1), Fraction(0, 1), Fraction(21, 4000000000), Fraction(69, 20000000), Fraction(9, 2000000),
               Fraction(23, 2000000), Fraction(3, 6250), Fraction(9, 5000), Fraction(1, 500), Fraction(83, 20000),
               Fraction(61, 2000), Fraction(1, 1), Fraction(1, 1), Fraction(169, 2), Fraction(495, 1), Fraction(990, 1),
               Fraction(1750, 1), Fraction(1150000, 1), Fraction(220000000, 1)],
}


@pytest.mark.parametrize("digits,resolution", sorted(EXPECTED))
def test_round_to_fraction_tables(digits, resolution):
    for value, want in zip(VALUES, EXPECTED[(digits, resolution)]):
        got = round_to_fraction(value, digits=digits, resolution=resolution)
        assert isinstance(got, Fraction)
        assert got == want, (value, digits, resolution, got, want)


def test_round_to_fraction_defaults_are_2_digits_unit_resolution():
    for value, want in zip(VALUES, EXPECTED[(2, 1)]):
        assert round_to_fraction(value) == want


@pytest.mark.parametrize("digits,table", [
    (1, [(-1, -1.0), (0, 0.0), (1, 1.0), (1.2, 1.25), (1.3, 1.25), (1.4, 1.5), (1.45, 1.5), (1.51, 1.5), (1.7, 1.75),
         (1.9, 2.0), (1.96, 2.0), (1.98, 2.0), (2, 2.0)]),
    (2, [(-1, -1.0), (0, 0.0), (1, 1.0), (1.2, 1.2), (1.23, 1.225), (1.3, 1.3), (1.4, 1.4), (1.45, 1.45), (1.51, 1.5),
         (1.7, 1.7), (1.79, 1.8), (1.9, 1.9), (1.96, 1.95), (1.98, 1.975), (2, 2.0)]),
])
def test_round_to_fraction_quarter_steps(digits, table):
    """test_helpers.py:51-89."""
    for value, want in table:
        assert float(round_to_fraction(value, digits, 0.25)) == pytest.approx(want, abs=1e-7)


def test_round_to_fraction_rejects_bad_arguments():
    for kwargs in (dict(digits=0), dict(resolution=0), dict(resolution=0.12)):
        with pytest.raises(ValueError):
            round_to_fraction(0.29, **kwargs)


@pytest.mark.parametrize("value,want", [(90.0001, 90), (90.001, 90.001), (89.9999, 90), (89.999, 89.999),
                                        (0.99999, 1), (0.9999, 0.9999), (7, 7)])
def test_to_int_or_float(value, want):
    got = _to_int_or_float(value)
    assert got == want and type(got) is type(want)


def test_normalize_number_pair():
    assert _normalize_number_pair(5) == (5, 5)
    assert _normalize_number_pair(3.5) == (3.5, 3.5)
    assert _normalize_number_pair((2, 4)) == (2, 4)
    assert _normalize_number_pair((1.5, 2.5)) == (1.5, 2.5)
    assert _normalize_number_pair(None, default=(10, 20)) == (10, 20)
    with pytest.raises(ValueError) as e:
        _normalize_number_pair(None, name="test_var")
    assert "test_var must be a number or a sequence of two numbers" in str(e.value)


def test_longitude_conventions():
    assert np.array_equal(to_lon_360(np.array([-10, 0, 45, 190, -180])), np.array([350, 0, 45, 190, 180]))
    assert np.array_equal(from_lon_360(np.array([350, 0, 45, 190, 180])), np.array([-10, 0, 45, -170, 180]))
