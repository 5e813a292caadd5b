"""`mpc_fatigue` — the package name the reference's scripts import (`import mpc_fatigue.pynocchio_casadi as pin`,
python/Libraries/Centauro_functions.py:2, python/Pilz_6_DOF/force_optimization_pilz_6DOF.py).  Here it holds the compiled
pybind11 module `pynocchio_casadi` built from bindings/python/pynocchio_casadi.cpp against libmpcf.so, so the import line
stays unchanged.  The implementation lives in `mpc_fatigue_b200`."""
