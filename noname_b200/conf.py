"""AppConf-style `key = value` configuration files, layered exactly as the reference does it
(src/NoName.jl:58-72): src/default.conf -> ~/.noname/default.conf -> the file on the command line;
later layers win.  AppConf evals each value as Julia; here values are parsed into Python:
numbers, true/false, "strings", [a, b, c] / (a, b, c) lists, bare words (function names such as
`evolvePM`, or `cic`) as strings, and constructor calls (`cosmology(h=..)`, `scaleFreePS(-1, .3)`)
kept as strings for the consumer to parse.  `#` starts a comment.

Aliases named by the north star but absent from the reference's code are mapped
onto the real keys: dims -> TopgridDimensions (when the latter is not set by the user file),
cosm_Ωm / cosm_h -> Cosmology = cosmology(h=.., OmegaM=.., OmegaR=0.0).
"""
from __future__ import annotations

import os
import re

# verbatim values of the reference's src/default.conf (data, not code)
DEFAULTS = {
    "dims": [30, 1, 1], "Nghosts": [3, 3, 3],
    "DomainLeftEdge": [0.0, 0.0, 0.0], "DomainRightEdge": [1.0, 1.0, 1.0],
    "CourantFactor": 0.3, "CurrentOutputNumber": 0, "OutputDirectory": "/tmp/noname",
    "OutputSeparateDirectories": True, "LastOutputCycle": -1, "LastOutputTime": 0.0,
    "OutputEveryNCycle": 0, "OutputDtDataDump": 0.0, "verbose": False,
    "RestartFileName": "/tmp/restart.jld", "ProfilingResultsFile": "profiling_results.bin",
    "InitialDt": 0.0000001, "DtFractionOfTime": 10000.0, "MaximumDt": 1e30, "StartTime": 0,
    "CurrentTime": 0, "StopTime": 0, "StartCycle": 0, "CurrentCycle": 0, "StopCycle": 1000000,
    "ParticleDimensions": [1, 1, 1], "ParticleDepositInterpolation": "cic",
    "ParticleBackInterpolation": "cic",
}

_num = re.compile(r"^[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)$")


def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            if ch == q:
                q = None
        elif ch in "\"'":
            q = ch
        elif ch == "#":
            break
        out.append(ch)
    return "".join(out).strip()


def _split_top(s, sep=","):
    parts, depth, cur, q = [], 0, [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "\"'":
            q = ch
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return parts


def parse_value(s):
    s = s.strip()
    if s == "":
        return ""
    if s in ("true", "false"):
        return s == "true"
    if len(s) >= 2 and s[0] == s[-1] and s[0] in "\"'":
        return s[1:-1]
    if _num.match(s.replace("_", "")):
        t = s.replace("_", "")
        if re.match(r"^[+-]?\d+$", t):
            return int(t)
        return float(t)
    if (s[0] == "[" and s[-1] == "]") or (s[0] == "(" and s[-1] == ")"):
        return [parse_value(p) for p in _split_top(s[1:-1])]
    return s           # bare word or call expression


def parseconf(path):
    """One file -> dict (AppConf.parseconf)."""
    out = {}
    with open(path, encoding="utf-8") as f:
        for raw in f:
            line = _strip_comment(raw)
            if not line or "=" not in line:
                continue
            k, _, v = line.partition("=")
            out[k.strip()] = parse_value(v)
    return out


def load_conf(path=None, overrides=None, user_default=True):
    """Layered configuration (src/NoName.jl:58-72) + alias handling."""
    conf = dict(DEFAULTS)
    user = {}
    home = os.path.join(os.path.expanduser("~"), ".noname", "default.conf")
    if user_default and os.path.isfile(home):
        conf.update(parseconf(home))
    if path is not None:
        user = parseconf(path)
        conf.update(user)
    if overrides:
        user = {**user, **overrides}
        conf.update(overrides)
    if "TopgridDimensions" not in conf and "dims" in user:
        conf["TopgridDimensions"] = list(conf["dims"])
    if "Cosmology" not in conf and ("cosm_Ωm" in conf or "cosm_h" in conf):
        conf["Cosmology"] = "cosmology(h=%r,OmegaM=%r,OmegaR=0.0)" % (
            float(conf.get("cosm_h", 0.69)), float(conf.get("cosm_Ωm", 0.29)))
    return conf
