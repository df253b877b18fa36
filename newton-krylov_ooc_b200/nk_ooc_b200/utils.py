"""File comparison and units-string helpers of the reference's nk_ooc/utils.py that the CI flows need
(SURVEY.md §8 f-2 / f-3), on scipy's NETCDF3 reader (the files of this path are NETCDF3_64BIT_OFFSET;
netCDF4 and pint are not dependencies here).

* `metadata_same`, `isclose_all_vars` — nk_ooc/utils.py:212-300: same dimension names and lengths, same variable
  names, dimensions and ATTRIBUTES; values close under rtol / atol with matching _FillValue patterns.
* `units_str_format`, `units_product` — nk_ooc/utils.py:186-205 canonicalises unit strings with pint's short
  format ("years m", "mmol / m^2 / s", "1 / s"); here a small parser does the same for the products and quotients
  of named units with integer powers that the hist files use.
"""

import logging

import numpy as np
from scipy.io import netcdf_file

# ---- units strings ------------------------------------------------------------------------------
_SORT_ALIAS = {"years": "a", "year": "a"}  # pint orders by symbol: the year is 'a'


def _parse_units(units_str):
    """{unit name: integer power} of strings like 'mmol / m^3 / s', 'years m', '(years) (m)', '1 / s', 'm^2 / s'"""
    powers = {}
    text = units_str.replace("(", " ").replace(")", " ").replace("**", "^").replace("*", " ")
    sign = 1
    for tok in text.replace("/", " / ").split():
        if tok == "/":
            sign = -1  # every term after a '/' up to the next '/' divides: 'a / b / c' = a b^-1 c^-1
            continue
        name, _, power = tok.partition("^")
        if name not in ("1", ""):
            powers[name] = powers.get(name, 0) + sign * (int(power) if power else 1)
    return {k: v for k, v in powers.items() if v != 0}


def _format_units(powers):
    def term(name, power):
        return name if power == 1 else f"{name}^{power}"

    names = sorted(powers, key=lambda n: _SORT_ALIAS.get(n, n))
    num = [term(n, powers[n]) for n in names if powers[n] > 0]
    den = [term(n, -powers[n]) for n in names if powers[n] < 0]
    res = " ".join(num) if num else "1"
    res = " / ".join([res] + den)
    # utils.py:201-204: 'x / s / y' and 'x / d / y' are written with the time unit last
    parts = res.split(" / ")
    if len(parts) == 3 and parts[1] in ("d", "s"):
        res = " / ".join([parts[0], parts[2], parts[1]])
    return res


def units_str_format(units_str):
    """units string in the reference's canonical format (utils.py:186-205)"""
    return _format_units(_parse_units(units_str))


def units_product(*units_strs):
    """canonical units string of a product, e.g. ('mmol / m^3', 'm') -> 'mmol / m^2'"""
    total = {}
    for units_str in units_strs:
        for name, power in _parse_units(units_str).items():
            total[name] = total.get(name, 0) + power
    return _format_units({k: v for k, v in total.items() if v != 0})


# named units of the files on this path: name -> (dimension, size in the dimension's reference unit).  The year is
# pint's Julian year (365.25 d), as the reference's `ureg` has it.
_UNIT_TABLE = {
    "m": ("length", 1.0), "cm": ("length", 1.0e-2), "mm": ("length", 1.0e-3), "km": ("length", 1.0e3),
    "s": ("time", 1.0), "min": ("time", 60.0), "h": ("time", 3600.0), "hr": ("time", 3600.0), "hour": ("time", 3600.0),
    "d": ("time", 86400.0), "day": ("time", 86400.0), "days": ("time", 86400.0),
    "a": ("time", 365.25 * 86400.0), "yr": ("time", 365.25 * 86400.0), "year": ("time", 365.25 * 86400.0),
    "years": ("time", 365.25 * 86400.0),
    "mol": ("amount", 1.0), "mmol": ("amount", 1.0e-3), "umol": ("amount", 1.0e-6), "nmol": ("amount", 1.0e-9),
    "kg": ("mass", 1.0), "g": ("mass", 1.0e-3), "mg": ("mass", 1.0e-6),
    "L": ("length3", 1.0e-3), "l": ("length3", 1.0e-3),
}


def units_conversion_factor(units_from, units_to):
    """factor f with (value in units_from) * f = (value in units_to), for products and quotients of the named units
    of _UNIT_TABLE with integer powers — what `ureg.Quantity(vals, u1).to(u2)` does for the reference
    (utils.py:306-310).  ValueError when the units are not commensurable or hold a name the table does not know."""
    def reduce(units_str):
        dims, size = {}, 1.0
        for name, power in _parse_units(units_str).items():
            if name not in _UNIT_TABLE:
                raise ValueError(f"unknown unit '{name}' in '{units_str}'")
            dim, unit_size = _UNIT_TABLE[name]
            if dim == "length3":
                dim, power_dim = "length", 3 * power
            else:
                power_dim = power
            dims[dim] = dims.get(dim, 0) + power_dim
            size *= unit_size ** power
        return {k: v for k, v in dims.items() if v != 0}, size

    dims1, size1 = reduce(units_from)
    dims2, size2 = reduce(units_to)
    if dims1 != dims2:
        raise ValueError(f"cannot convert from '{units_from}' to '{units_to}'")
    return size1 / size2


# ---- netCDF comparison ----------------------------------------------------------------------------
def _attrs(var):
    out = {}
    for key, val in var._attributes.items():  # pylint: disable=protected-access
        if isinstance(val, bytes):
            val = val.decode()
        elif isinstance(val, np.ndarray):
            val = val.tolist()
        out[key] = val
    return out


def _dimlens(nc):
    """{dimension: length}, the unlimited dimension by its current length"""
    out = {}
    for name, length in nc.dimensions.items():
        if length is None:
            length = next((v.shape[0] for v in nc.variables.values() if v.dimensions and v.dimensions[0] == name), 0)
        out[name] = length
    return out


def metadata_same(fname1, fname2):
    """True if the metadata of the two files is the same (utils.py:212-258)"""
    logger = logging.getLogger(__name__)
    res = True
    with netcdf_file(fname1, "r", mmap=False) as f1, netcdf_file(fname2, "r", mmap=False) as f2:
        d1, d2 = _dimlens(f1), _dimlens(f2)
        if d1.keys() != d2.keys():
            logger.info("    dimension name mismatch in %s and %s", fname1, fname2)
            res = False
        for dimname in d1:
            if dimname in d2 and d1[dimname] != d2[dimname]:
                logger.info("    %s length mismatch in %s and %s", dimname, fname1, fname2)
                res = False
        if f1.variables.keys() != f2.variables.keys():
            logger.info("    variable name mismatch in %s and %s: %s", fname1, fname2,
                        sorted(set(f1.variables) ^ set(f2.variables)))
            res = False
        for varname, var1 in f1.variables.items():
            if varname in f2.variables:
                var2 = f2.variables[varname]
                if var1.dimensions != var2.dimensions:
                    logger.info("    %s dimension mismatch in %s and %s", varname, fname1, fname2)
                    res = False
                if _attrs(var1) != _attrs(var2):
                    logger.info("    %s attribute mismatch in %s and %s: %s != %s", varname, fname1, fname2,
                                _attrs(var1), _attrs(var2))
                    res = False
    return res


def _native(var):
    return np.array(var.data, dtype=var.data.dtype.newbyteorder("="))


def _isclose_one_var_core(vals1, vals2, rtol, atol):
    """utils.py:303-324"""
    logger = logging.getLogger(__name__)
    close = np.isclose(vals1, vals2, rtol=rtol, atol=atol, equal_nan=True)
    if close.all():
        return True
    flat1, flat2 = np.asarray(vals1, dtype=float).reshape(-1), np.asarray(vals2, dtype=float).reshape(-1)
    for ind in np.nonzero(~close.reshape(-1))[0][:20]:
        val1, val2 = flat1[ind], flat2[ind]
        logger.info("    %.10e %.10e not close, atol_adj=%e, rtol_adj=%e", val1, val2,
                    abs(val1 - val2) - rtol * abs(val2), (abs(val1 - val2) - atol) / abs(val2) if val2 != 0 else np.inf)
    return False


def _isclose_one_var(name, var1, var2, rtol, atol, parent2=None):
    """utils.py:261-300.  parent2: for an anomaly variable (x - mean(x)), the baseline's x: the absolute tolerance
    becomes atol + rtol |x|, the tolerance x itself is compared with (see isclose_all_vars)"""
    logger = logging.getLogger(__name__)
    if var1.shape != var2.shape:
        logger.info("    var1.shape %s != var2.shape %s for %s", var1.shape, var2.shape, name)
        return False
    res = True
    vals1, vals2 = _native(var1), _native(var2)
    if vals1.dtype.kind in "SU" or vals2.dtype.kind in "SU":
        return bool((vals1 == vals2).all())
    att1, att2 = _attrs(var1), _attrs(var2)
    msv1, msv2 = att1.get("_FillValue"), att2.get("_FillValue")
    fill1 = (vals1 == msv1) if msv1 is not None else np.zeros(vals1.shape, bool)
    fill2 = (vals2 == msv2) if msv2 is not None else np.zeros(vals2.shape, bool)
    if (fill1 != fill2).any():
        logger.info("    _FillValue pattern mismatch for %s", name)
        res = False
    if (fill1 | fill2).any():
        vals1 = np.where(fill1 | fill2, np.nan, vals1)
        vals2 = np.where(fill1 | fill2, np.nan, vals2)
    units1, units2 = att1.get("units"), att2.get("units")
    if units1 is not None and units2 is not None and units1 != units2:
        if "since" in units1 or "since" in units2:
            raise ValueError(f"time-like units disagree '{units1}'!='{units2}'")
        if _parse_units(units1) != _parse_units(units2):
            # utils.py:306-310: the values of the first file are converted to the units of the second
            try:
                vals1 = vals1 * units_conversion_factor(units1, units2)
            except ValueError as err:
                logger.info("    units of %s differ and cannot be converted: %s", name, err)
                res = False
    if parent2 is not None and parent2.shape == vals2.shape:
        # |v1 - v2| <= atol + rtol (|x| + |v2|) pointwise, written as a comparison of scaled values
        scale = atol + rtol * (np.abs(parent2) + np.abs(vals2))
        close = np.abs(vals1 - vals2) <= scale
        close |= np.isnan(vals1) & np.isnan(vals2)
        if not close.all():
            worst = np.nanmax(np.where(close, 0.0, np.abs(vals1 - vals2) / scale))
            logger.info("    %s vals not close at the parent variable's tolerance (worst ratio %.3f)", name, worst)
            res = False
        return res
    if not _isclose_one_var_core(vals1, vals2, rtol=rtol, atol=atol):
        logger.info("    %s vals not close", name)
        res = False
    return res


def isclose_all_vars(fname1, fname2, rtol, atol, anom_suffix=None):
    """True if all variables common to both files are close (utils.py:261-272).

    anom_suffix (an extension, off by default): variables `<x><anom_suffix>` are anomalies x - mean(x) of a
    variable x of the same file.  They are compared with the tolerance x is compared with, atol + rtol |x|: where
    the anomaly crosses zero its plain tolerance is atol alone, below what two different time integrators of x
    can agree to (the reference's own Radau solution at rtol = atol = 1e-6 is 1.5e-6 from the converged one on
    the CI grid, profiles/r02_error_vs_steps.md)."""
    res = True
    with netcdf_file(fname1, "r", mmap=False) as f1, netcdf_file(fname2, "r", mmap=False) as f2:
        for varname, var1 in f1.variables.items():
            if varname in f2.variables:
                parent2 = None
                if anom_suffix and varname.endswith(anom_suffix) and varname[: -len(anom_suffix)] in f2.variables:
                    parent2 = _native(f2.variables[varname[: -len(anom_suffix)]])
                if not _isclose_one_var(varname, var1, f2.variables[varname], rtol=rtol, atol=atol, parent2=parent2):
                    res = False
    return res
