"""TEST INFRASTRUCTURE (build container only): runs the REFERENCE's own NewtonSolver / KrylovSolver
(/root/reference/nk_ooc/newton_solver.py, krylov_solver.py, solver_base.py, solver_state.py, stats_file.py, imported
unmodified) over the numpy stand-in state of tests/fake_state.py and stores what they did in
tests/golden/ref_solver_<problem>.json: every Newton iterate and function value, the Newton and Krylov step logs, the
Krylov solvers' saved Hessenberg matrices.  tests/test_solver_host.py runs THIS package's solvers over the same class
and compares.  netCDF4 is not installed here: the stats files go to an in-memory stand-in of netCDF4.Dataset (the
solvers never read them back, apart from the length of the iteration dimension).

    python -m oracle.gen_golden_solver
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness  # noqa: E402

SOLVERINFO = {"newton_rel_tol": "1.0e-8", "newton_max_iter": "12", "post_newton_fp_iter": "1", "krylov_rel_tol": "0.01"}
# golden name -> (problem of tests/fake_state.py, solverinfo overrides)
CASES = {
    "mild": ("mild", {}),
    "damped": ("damped", {}),
    "regions": ("regions", {}),
    # the iteration floors of solver_base.py:61-68 (input/cime_pop/newton_krylov.cfg:46 runs with krylov_min_iter = 4)
    "min_iter": ("regions", {"krylov_min_iter": "4", "newton_min_iter": "3"}),
}
_FILES = {}
PERSIST = False  # also write the stats files to disk (tests that hand a reference solve over to this package's solvers)
_NOT_ATTRS = ("name", "dimensions", "data", "_fptr", "_loose", "_datatype")


class _Section(dict):
    """a configparser section: keys are case-insensitive (the reference asks for "Newton_rel_tol")"""

    def __getitem__(self, key):
        return super().__getitem__(key.lower())

    def __setitem__(self, key, val):
        super().__setitem__(key.lower(), val)

    def __contains__(self, key):
        return super().__contains__(key.lower())


class _MemVar:
    def __init__(self, name, dims, fptr, fill, datatype="f8"):
        self.name, self.dimensions, self._fptr, self._datatype = name, tuple(dims), fptr, datatype
        if fill is not None:
            self._FillValue = fill
        self.data = np.zeros([fptr.dimensions[d].size or 0 for d in dims])

    def setncatts(self, attrs):
        self.__dict__.update(attrs)

    def __len__(self):
        return self.data.shape[0]

    def _grow(self, n):
        if self.dimensions and self._fptr.dimensions[self.dimensions[0]].size is None and n > self.data.shape[0]:
            pad = np.zeros((n - self.data.shape[0],) + self.data.shape[1:])
            self.data = np.concatenate([self.data, pad])

    def __setitem__(self, key, val):
        first = key[0] if isinstance(key, tuple) else key
        if isinstance(first, (int, np.integer)):
            if getattr(self, "_loose", False) and np.shape(val) != self.data.shape[1:]:
                self.data = np.zeros((self.data.shape[0],) + np.shape(val))  # (first write: takes the value's shape)
            self._grow(int(first) + 1)
        self.data[key] = val

    def __getitem__(self, key):
        return self.data[key]


class _MemDim:
    def __init__(self, size, fptr, name):
        self.size, self._fptr, self._name = size, fptr, name

    def __len__(self):
        if self.size is not None:
            return self.size
        lens = [len(v) for v in self._fptr.variables.values() if v.dimensions and v.dimensions[0] == self._name]
        return max(lens, default=0)


class _LooseVars(dict):
    """variables of a stats file that was NOT written through this stand-in (a solve of this package's solvers that the
    reference resumes, tests/test_solver_host.py): whatever the reference writes to is created on first use"""

    def __init__(self, fptr):
        super().__init__()
        self._fptr = fptr

    def __missing__(self, name):
        self[name] = _MemVar(name, ("iteration", "region"), self._fptr, 9.969209968386869e36)
        self[name]._loose = True  # pylint: disable=protected-access
        return self[name]


class _MemDataset:
    """netCDF4.Dataset stand-in: files written by the solvers live in memory, everything else is read from disk"""

    def __new__(cls, fname, mode="r", **kwargs):
        if mode == "r" and fname not in _FILES:
            return ref_harness._Dataset(fname, mode, **kwargs)  # pylint: disable=protected-access
        return super().__new__(cls)

    def __init__(self, fname, mode="r", **kwargs):
        if mode == "w":
            _FILES[fname] = {"dimensions": {}, "variables": {}, "attrs": {}}
        loose = fname not in _FILES
        if loose:
            _FILES[fname] = {"dimensions": {}, "variables": _LooseVars(self), "attrs": {"history": ""}}
        store = _FILES[fname]
        self.__dict__["_store"] = store
        self.__dict__["dimensions"] = store["dimensions"]
        self.__dict__["variables"] = store["variables"]
        if loose:
            store["dimensions"].update(iteration=_MemDim(None, self, "iteration"), region=_MemDim(1, self, "region"))
        if isinstance(store["variables"], _LooseVars):
            store["variables"]._fptr = self  # pylint: disable=protected-access

    def __setattr__(self, key, val):
        self._store["attrs"][key] = val

    def __getattr__(self, key):
        try:
            return self.__dict__["_store"]["attrs"][key]
        except KeyError as err:
            raise AttributeError(key) from err

    def createDimension(self, name, size):  # noqa: N802
        if name in self.dimensions:
            raise RuntimeError("NetCDF: String match to name in use")
        self.dimensions[name] = _MemDim(size, self, name)

    def createVariable(self, name, datatype, dims, fill_value=None):  # noqa: N802
        self.variables[name] = _MemVar(name, dims, self, fill_value, datatype)
        return self.variables[name]

    def _persist(self, fname):
        """the same content as a real NETCDF3_64BIT_OFFSET file (scipy's writer), so that a solve the reference was
        interrupted in can be resumed by this package's solvers, whose StatsFile reads the file back"""
        from scipy.io import netcdf_file

        with netcdf_file(fname, "w", version=2) as nc:
            for key, val in self._store["attrs"].items():
                setattr(nc, key, val)
            for name, dim in self.dimensions.items():
                nc.createDimension(name, dim.size)
            for var in self.variables.values():
                out = nc.createVariable(var.name, "i4" if var._datatype == "i4" else "f8", var.dimensions)
                for key, val in vars(var).items():
                    if key not in _NOT_ATTRS:
                        setattr(out, key, val)
                data = np.asarray(var.data)
                if var.dimensions and self.dimensions[var.dimensions[0]].size is None:
                    if data.shape[0] > 0:
                        out[: data.shape[0]] = data
                else:
                    out[:] = data

    def sync(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        fname = next(name for name, store in _FILES.items() if store is self._store)
        if PERSIST and not isinstance(self.variables, _LooseVars):
            self._persist(fname)
        return False


def reference_newton_solver():
    """the reference's NewtonSolver class, its stats files redirected to the in-memory stand-in"""
    ref_harness.install_stubs()
    import netCDF4  # the stub installed by ref_harness

    netCDF4.Dataset = _MemDataset
    for name in list(sys.modules):
        if name.startswith("nk_ooc.") and hasattr(sys.modules[name], "Dataset"):
            sys.modules[name].Dataset = _MemDataset
    from nk_ooc.newton_solver import NewtonSolver

    return NewtonSolver


def solverinfo(workdir, init_iterate_fname=None, **kw):
    return _Section(dict(SOLVERINFO, workdir=workdir, init_iterate_fname=init_iterate_fname, **kw))


def run_reference(case):
    """the reference's driver loop (nk_driver.py:58-66) over FakeState; returns the record for the golden file"""
    NewtonSolver = reference_newton_solver()  # noqa: N806

    from fake_state import FakeState

    problem, overrides = CASES[case]
    FakeState.configure(problem)
    with tempfile.TemporaryDirectory() as work:
        init = os.path.join(work, "init_iterate.nc")
        FakeState(np.ones(6)).dump(init)
        info = solverinfo(work, init, **overrides)
        solver = NewtonSolver(FakeState, info, resume=False, rewind=False)
        while not solver.converged().all():
            solver.step()
        n_iter = solver.get_iteration()

        def arr(name):
            return FakeState(os.path.join(work, name)).vals.tolist()

        def state(path):
            with open(path) as fptr:
                rec = json.load(fptr)
            rec["step_log"] = [s.replace(work, "W") for s in rec["step_log"]]
            return rec

        def stats(path):
            """what the reference wrote into a stats file, call by call (dimensions, variables in definition order,
            attributes, values)"""
            store = _FILES[path]
            return {
                "attrs": sorted(store["attrs"]),
                "dimensions": [[name, dim.size, len(dim)] for name, dim in store["dimensions"].items()],
                "variables": [
                    {"name": var.name, "dimensions": list(var.dimensions),
                     "attrs": {k: v for k, v in vars(var).items() if k not in _NOT_ATTRS},
                     "data": np.asarray(var.data).tolist()}
                    for var in store["variables"].values()],
            }

        rec = {
            "problem": problem, "solverinfo": dict(SOLVERINFO, **overrides), "iterations": n_iter, "evaluations": FakeState.calls,
            "iterate": [arr(f"iterate_{i:02}.nc") for i in range(n_iter + 1)],
            "fcn": [arr(f"fcn_{i:02}.nc") for i in range(n_iter + 1)],
            "increment": [arr(f"increment_{i:02}.nc") for i in range(n_iter)],
            "Newton_state": state(os.path.join(work, "Newton_state.json")),
            "Krylov_state": [state(os.path.join(work, f"krylov_{i:02}", "Krylov_state.json")) for i in range(n_iter)],
            "files": sorted(os.path.relpath(os.path.join(d, f), work) for d, _, fs in os.walk(work) for f in fs),
            "Newton_stats": stats(os.path.join(work, "Newton_stats.nc")),
            "Krylov_stats": [stats(os.path.join(work, f"krylov_{i:02}", "Krylov_stats.nc")) for i in range(n_iter)],
            "Armijo_factor": [np.asarray(_FILES[os.path.join(work, "Newton_stats.nc")]["variables"]
                                         [f"Armijo_factor_{tm.name}"].data)[:n_iter].tolist() for tm in solver._iterate.tracer_modules],
        }
    FakeState.configure("mild")
    return rec


def main():
    ref_harness.install_stubs()
    if ref_harness.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REF_ROOT)
    for problem in CASES:
        rec = run_reference(problem)
        path = os.path.join(ROOT, "tests", "golden", f"ref_solver_{problem}.json")
        with open(path, "w") as fptr:
            json.dump(rec, fptr, indent=1)
        print(problem, "Newton iterations", rec["iterations"], "evaluations", rec["evaluations"], "Armijo factors",
              rec["Armijo_factor"], "->", path)


if __name__ == "__main__":
    main()
