"""Experimental build of the library with extra -D flags, for A/B timing on the GPU box:
    python scripts/build_exp.py <name> [-DFLAG ...]   ->  dial_mpc_b200/csrc/exp/libdial_b200_<name>.so
    DIAL_B200_LIB=dial_mpc_b200/csrc/exp/libdial_b200_<name>.so python scripts/prof_cfg.py 1 3 --time"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
d = os.path.join(g.CSRC, "exp")
os.makedirs(d, exist_ok=True)
lib = os.path.join(d, f"libdial_b200_{name}.so")
g.build_library(extra=flags, lib=lib, objdir=os.path.join(g.OBJDIR, "exp_" + name))
print(lib)
