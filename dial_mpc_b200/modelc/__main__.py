"""``python -m dial_mpc_b200.modelc scene.xml -o model.json`` — compile an MJCF file."""
import argparse

from .mjcf import compile_mjcf


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("xml")
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--name", default=None)
    a = ap.parse_args()
    m = compile_mjcf(a.xml, a.name)
    m.save(a.out)
    print(f"{a.out}: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} ngeom={m.ngeom} "
          f"ncon={m.ncon} meaninertia={m.meaninertia:.6g}")


if __name__ == "__main__":
    main()
