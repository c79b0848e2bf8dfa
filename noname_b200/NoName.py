"""Driver surface of the reference (src/NoName.jl:19-100): `run_noname()` -> gp, gd, p.

conf["Initializer"] / conf["Evolve"] / conf["ProblemDefinition"] name functions, as in the
reference (function-name dispatch, NoName.jl:82-83); here they are looked up in registries
instead of eval'ed.  Restart (src/restart.jl) uses an .npz file with the JLD file's keys
(conf, grids, gridData, particles): JLD/HDF5 are not available to this host.
"""
from __future__ import annotations

import argparse
import pickle

import numpy as np

from . import ParticleMesh as PM
from .conf import load_conf

conf = {}


def write_all(fname, conf_, gp, gd, p):
    """src/IOtools.jl:118-122"""
    blob = {"conf": {k: v for k, v in conf_.items() if not k.startswith("_")}, "grids": gp, "gridData": gd,
            "particles": {k: v for k, v in p.items() if not k.startswith("_")}}
    with open(fname, "wb") as f:
        pickle.dump(blob, f)


def load_all(fname, conf_, gp, gd, p):
    """src/IOtools.jl:124-138"""
    with open(fname, "rb") as f:
        blob = pickle.load(f)
    conf_.update(blob["conf"])
    gp.update(blob["grids"])
    gd.update(blob["gridData"])
    p.update(blob["particles"])


def restart(gp, gd, p, conf_):
    """src/restart.jl (with the arity bug of SURVEY.md F10 fixed)."""
    load_all(conf_["_restart_file"], conf_, gp, gd, p)
    if conf_.get("_rconf"):
        from .conf import parseconf
        conf_.update(parseconf(conf_["_rconf"]))
    conf_["StartTime"] = conf_["CurrentTime"]      # resume where the dump was taken
    conf_["StartCycle"] = conf_["CurrentCycle"]
    if "Cosmology" in conf_:
        from .cosmology import parse_cosmology
        conf_["_cosmo"] = parse_cosmology(conf_["Cosmology"])
    conf_["_rank"] = p["x"].shape[1]


INITIALIZERS = {"InitializeParticleSimulation": PM.InitializeParticleSimulation, "restart": restart}


def noname(conf_):
    """NoName.jl:19-39"""
    gp, gd, p = {}, {}, {}
    INITIALIZERS[conf_["Initializer"]](gp, gd, p, conf_)
    PM.EVOLVERS[conf_["Evolve"]](gp, gd, p, conf_)
    if conf_.get("RestartFileName") and not conf_.get("OutputDisabled"):
        write_all(conf_["RestartFileName"], conf_, gp, gd, p)
    if conf_.get("ProfilingOn", False) and conf_.get("_profile") and conf_.get("ProfilingResultsFile"):
        write_profiling_results(conf_["ProfilingResultsFile"], conf_["_profile"])
    return gp, gd, p


def write_profiling_results(fname, prof):
    """src/IOtools.jl:142-147 writes Julia's sampling-profiler dump; here: total and per-step CUDA-event milliseconds of
    every phase of the device step (deposit, solve, push, sort, exchange, comm, fft_*) and the launch count, as JSON."""
    import json
    n = max(1, prof["steps"])
    out = {"steps": prof["steps"], "kernel_launches": prof["kernel_launches"],
           "ms_total": prof["ms"], "ms_per_step": {k: v / n for k, v in prof["ms"].items()}}
    with open(fname, "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1)


def run_noname(configuration_file=None, restart_=False, rconf=None, overrides=None, argv=None):
    """NoName.jl:51-100.  With no arguments, parses the command line like IOtools.parse_commandline
    (positional configuration_file, --restart/-r, --rconf)."""
    global conf
    if configuration_file is None:
        ap = argparse.ArgumentParser()
        ap.add_argument("configuration_file")
        ap.add_argument("--restart", "-r", action="store_true")
        ap.add_argument("--rconf", default="")
        a = ap.parse_args(argv)
        configuration_file, restart_, rconf = a.configuration_file, a.restart, a.rconf
    if restart_:
        conf = load_conf(None, overrides)
        conf["Initializer"] = "restart"
        conf["_restart_file"] = configuration_file
        conf["_rconf"] = rconf
    else:
        conf = load_conf(configuration_file, overrides)
    return noname(conf)
