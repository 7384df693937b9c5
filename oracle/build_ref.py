#!/usr/bin/env python
"""Build the reference's own CPU path into ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

This is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import what this produces.

What it does (recipe: SURVEY.md App. B, which restates /root/reference/setup.py:209-226,451-578):

* stages the reference's *unmodified* sources from where they lie under ``/root/reference`` into the
  git-ignored build directory ``oracle/_ref/assimulo`` (the same staging ``setup.py`` does into
  ``build/assimulo``) -- nothing is copied into tracked files;
* cythonizes + compiles ``ode, problem, support, explicit_ode (+ode_event_locator.c), implicit_ode,
  solvers/euler`` and ``lib/radau5ode (+radau5.c, radau5_io.c, no SuperLU)`` with
  ``-O2 -fno-strict-aliasing`` (setup.py:564);
* ``oracle/rodas_c.c`` + ``oracle/rodas_f2py_shim.py`` do the same for RODAS (thirdparty/hairer/rodas_decsol.f ->
  ``assimulo/lib/rodas.py``), so that ``rosenbrock.py::RodasODE`` runs unmodified;
* the Fortran DOPRI5 core (thirdparty/hairer/dopri5.f) cannot be built here (no Fortran compiler, no
  numpy.distutils): ``oracle/dopri5_c.c`` -- our C restatement of that file -- is compiled to
  ``oracle/_ref/libdopri5_c.so`` and ``oracle/dopri5_f2py_shim.py`` is installed as
  ``assimulo/lib/dopri5.py`` presenting the f2py signature of thirdparty/hairer/dopri5.pyf:31-61, so that the
  reference's ``runge_kutta.py::Dopri5`` class runs unmodified on top of it.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def _copy(files, src, dst):
    os.makedirs(dst, exist_ok=True)
    for f in files:
        p = os.path.join(src, f)
        if os.path.isfile(p):
            shutil.copy2(p, os.path.join(dst, f))


def stage(ref):
    pkg = os.path.join(OUT, "assimulo")
    src = os.path.join(ref, "src")
    _copy(os.listdir(src), src, pkg)
    _copy(os.listdir(os.path.join(src, "lib")), os.path.join(src, "lib"), os.path.join(pkg, "lib"))
    _copy(os.listdir(os.path.join(src, "solvers")), os.path.join(src, "solvers"), os.path.join(pkg, "solvers"))
    r5 = os.path.join(ref, "thirdparty", "radau5")
    _copy(os.listdir(r5), r5, os.path.join(pkg, "thirdparty", "radau5"))
    open(os.path.join(pkg, "thirdparty", "__init__.py"), "a").close()
    # our f2py stand-ins for the Fortran modules (NOT reference code)
    shutil.copy2(os.path.join(HERE, "dopri5_f2py_shim.py"), os.path.join(pkg, "lib", "dopri5.py"))
    shutil.copy2(os.path.join(HERE, "rodas_f2py_shim.py"), os.path.join(pkg, "lib", "rodas.py"))


def build_dopri5_c():
    so = os.path.join(OUT, "libdopri5_c.so")
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared",
           os.path.join(HERE, "dopri5_c.c"), "-o", so, "-lm"]
    subprocess.check_call(cmd)
    return so


def build_rodas_c():
    so = os.path.join(OUT, "librodas_c.so")
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared",
           os.path.join(HERE, "rodas_c.c"), "-o", so, "-lm"]
    subprocess.check_call(cmd)
    return so


def build_cython():
    import numpy as np
    from Cython.Build import cythonize
    from setuptools import Extension  # noqa: F401  (cythonize returns Extension objects)
    from setuptools.dist import Distribution
    from setuptools.command.build_ext import build_ext

    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        directives = {"language_level": "3str"}
        exts = cythonize([os.path.join("assimulo", "explicit_ode.pyx")], include_path=[".", "assimulo"],
                         force=True, compiler_directives=directives, quiet=True)
        exts[-1].include_dirs += ["assimulo"]
        exts[-1].sources += [os.path.join("assimulo", "ode_event_locator.c")]
        exts += cythonize([os.path.join("assimulo", "%s.pyx" % x) for x in
                           ("implicit_ode", "ode", "problem", "support")],
                          include_path=[".", "assimulo"], force=True, compiler_directives=directives, quiet=True)
        exts += cythonize([os.path.join("assimulo", "solvers", "euler.pyx")],
                          include_path=[".", "assimulo", os.path.join("assimulo", "solvers")],
                          force=True, compiler_directives=directives, quiet=True)
        r5dir = os.path.join("assimulo", "thirdparty", "radau5")
        r5 = cythonize([os.path.join(r5dir, "radau5ode.pyx")],
                       include_path=[".", "assimulo", os.path.join("assimulo", "lib")],
                       force=True, compiler_directives=directives, quiet=True)
        r5[-1].sources += [os.path.join(r5dir, "radau5.c"), os.path.join(r5dir, "radau5_io.c")]
        r5[-1].include_dirs += ["assimulo", os.path.join("assimulo", "lib"), r5dir]
        r5[-1].name = "assimulo.lib.radau5ode"
        r5[-1].libraries = ["m"]
        exts += r5
        for e in exts:
            e.include_dirs += [np.get_include()]
            e.extra_compile_args += ["-O2", "-fno-strict-aliasing", "-w"]
        dist = Distribution({"name": "assimulo_ref", "ext_modules": exts})
        cmd = build_ext(dist)
        cmd.inplace = 1
        cmd.build_temp = os.path.join(OUT, "_build_tmp")
        cmd.ensure_finalized()
        cmd.run()
    finally:
        os.chdir(cwd)


def is_built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    need = ["assimulo/ode", "assimulo/problem", "assimulo/support", "assimulo/explicit_ode",
            "assimulo/implicit_ode", "assimulo/solvers/euler", "assimulo/lib/radau5ode"]
    ok = all(os.path.exists(os.path.join(OUT, n + suffix)) for n in need)
    return ok and os.path.exists(os.path.join(OUT, "libdopri5_c.so")) and \
        os.path.exists(os.path.join(OUT, "assimulo", "lib", "dopri5.py"))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args(argv)
    if is_built() and not a.force:
        # refresh only our own pieces (cheap)
        build_dopri5_c()
        build_rodas_c()
        shutil.copy2(os.path.join(HERE, "dopri5_f2py_shim.py"), os.path.join(OUT, "assimulo", "lib", "dopri5.py"))
        shutil.copy2(os.path.join(HERE, "rodas_f2py_shim.py"), os.path.join(OUT, "assimulo", "lib", "rodas.py"))
        print("oracle/_ref already built")
        return 0
    if not os.path.isdir(a.reference):
        print("reference tree %s not present; cannot build oracle/_ref" % a.reference, file=sys.stderr)
        return 1
    os.makedirs(OUT, exist_ok=True)
    stage(a.reference)
    build_dopri5_c()
    build_rodas_c()
    build_cython()
    shutil.rmtree(os.path.join(OUT, "_build_tmp"), ignore_errors=True)
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    for c in glob.glob(os.path.join(OUT, "assimulo", "**", "*.c"), recursive=True):
        if "thirdparty" not in c and not c.endswith("ode_event_locator.c"):
            os.remove(c)  # cython-generated C, large
    print("built oracle/_ref")
    return 0


if __name__ == "__main__":
    sys.exit(main())
