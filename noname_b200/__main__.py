"""`python -m noname_b200 particle_mesh.conf [--restart] [--rconf FILE]` -- the reference's command line
(`julia -e "using NoName; gp, gd, p = run_noname();" particle_mesh.conf`, src/NoName.jl:15-16, IOtools.parse_commandline).
With `NGPUs = P` in the conf file, launch it once per GPU: `torchrun --nproc-per-node P -m noname_b200 particle_mesh.conf`."""
from .NoName import run_noname

if __name__ == "__main__":
    run_noname()
