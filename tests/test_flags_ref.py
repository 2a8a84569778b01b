"""Pins the flag surface of row (b) to the REFERENCE'S OWN parser: Lib/Ziran/CS/Util/CommandLineFlags.h + Projects/multigrid/Configurations.h, compiled where
they lie with the flags of Projects/multigrid/main.cpp:40-84 registered (oracle/flags_ref_shim.cpp -> oracle/_ref/libflags_ref.so).  hot_b200::parseFlags
(include/hot_b200_host.hpp; driver tests/cpp/flags_ref.cpp) must accept and reject the same command lines and end at the same settings - on the command
lines of Projects/multigrid/tog.sh, on every registered flag, and on malformed input.  The expected values are committed below (they were produced by the
reference library); with the library present they are re-derived from it."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libflags_ref.so")
NAMES = ["cneps", "useAdaptiveHessian", "useCN", "matrixFree", "project", "systemBCProject", "linesearch", "boundaryType", "lsolver", "Ainv", "smoother",
         "coarseSolver", "levelCnt", "times", "levelscale", "debugMode", "omega", "topomega", "useBaselineMultigrid", "topDownMGS"]
DEFAULT = [1e-5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1.0, 0.1, 0, 0]


def _expect(**kw):
    v = list(DEFAULT)
    for k, x in kw.items():
        v[NAMES.index(k)] = x
    return v


# (command line, expected settings or None when the reference throws) - tog.sh:5-60 first
CASES = [
    ("-test 777001 --3d -cneps 1e-3 -lsolver 2 -Ainv 1 --project --linesearch -cmd0 1e9 --matfree -o twistbar_pnmf",
     _expect(cneps=1e-3, lsolver=2, Ainv=1, project=1, linesearch=1, matrixFree=1)),
    ("-test 777014 --3d --usecn -cneps 1e-7 -lsolver 2 -Ainv 1 --project --linesearch -o pn",
     _expect(useCN=1, cneps=1e-7, lsolver=2, Ainv=1, project=1, linesearch=1)),
    ("-test 777014 --3d --usecn -cneps 1e-7 -lsolver 3 -Ainv 1 --project --linesearch --bcproject -mg_level 3 -mg_times 1 -coarseSolver 2 -smoother 5 -o HOT",
     _expect(useCN=1, cneps=1e-7, lsolver=3, Ainv=1, project=1, linesearch=1, systemBCProject=1, levelCnt=3, times=1, coarseSolver=2, smoother=5)),
    ("-test 9212 --3d --usecn -cneps 1e-7 -lsolver 2 -Ainv 1 --project --linesearch --bcproject -mg_level 3 -mg_times 1 -coarseSolver 2 -smoother 5 -cmd0 1e6 -o x",
     _expect(useCN=1, cneps=1e-7, lsolver=2, Ainv=1, project=1, linesearch=1, systemBCProject=1, levelCnt=3, times=1, coarseSolver=2, smoother=5)),
    ("-test 777014 --3d --usecn -cneps 1e-7 -lsolver 3 -Ainv 1 --project --linesearch -mg_level 1 -mg_times 10000 -coarseSolver 2 -smoother 2 -o lbfgsH",
     _expect(useCN=1, cneps=1e-7, lsolver=3, Ainv=1, project=1, linesearch=1, levelCnt=1, times=10000, coarseSolver=2, smoother=2)),
    ("", list(DEFAULT)),
    ("--double --3d --adaptiveH -bc 1 -mg_scale 2 -mg_omega 0.8 -mg_jomega 0.3 -dbg 1 -t 8 -cmd1 2.5",
     _expect(useAdaptiveHessian=1, boundaryType=1, levelscale=2, omega=0.8, topomega=0.3, debugMode=1)),
    ("--help -script a.lua -i x=1 -i y=2 --run_diff_test -dtps 1e-3 -restart 4 -v_mu 0.5 --showresidual --showvcycle", list(DEFAULT)),
    ("--baseline --topDownMGS -lsolver 1", _expect(useBaselineMultigrid=1, topDownMGS=1, lsolver=1)),
    ("-lsolver 3 -lsolver 2", _expect(lsolver=2)),                 # a repeated flag: the last value stays
    ("--usecm", None),                                             # unknown flag
    ("-lsolver 3 extra", None),                                    # stray argument
    ("-mg_level", None),                                           # value missing
    ("--usecn -cneps", None),
]


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("flags_ref") / "flags_ref")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "flags_ref.cpp"),
                           "-o", exe, "-L", lib, "-lhot_b200", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def _product(exe, line):
    out = subprocess.run([exe] + line.split(), capture_output=True, text=True)
    if out.returncode != 0:
        assert out.stdout.startswith("error ")
        return None
    return [float(x) for x in out.stdout.split()]


def _reference(line):
    lib = C.CDLL(REF_LIB)
    args = ["multigrid"] + line.split()
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    out = (C.c_double * len(NAMES))()
    err = C.create_string_buffer(256)
    rc = lib.zr_flags_parse(len(args), argv, out, err, 256)
    return None if rc else list(out)


@pytest.mark.parametrize("line,expected", CASES, ids=[str(i) for i in range(len(CASES))])
def test_parse_flags_like_the_reference(driver, line, expected):
    got = _product(driver, line)
    assert (got is None) == (expected is None), (line, got)
    if expected is not None:
        assert got == [float(x) for x in expected], dict(zip(NAMES, got))


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref/libflags_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("line,expected", CASES, ids=[str(i) for i in range(len(CASES))])
def test_expected_values_are_the_reference_parsers(line, expected):
    got = _reference(line)
    assert (got is None) == (expected is None), (line, got)
    if expected is not None:
        assert got == [float(x) for x in expected], dict(zip(NAMES, got))
