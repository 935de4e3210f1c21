"""Packaging of the B200 DIAL-MPC sampling core.  The CUDA library is built in-tree by
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a); this file only installs the
Python surface and the console scripts the reference's setup.py defines (dial-mpc,
dial-mpc-plan; setup.py:24-31 of the reference)."""
from setuptools import find_packages, setup

setup(
    name="dial_mpc_b200",
    version="0.1.0",
    packages=find_packages(include=["dial_mpc_b200", "dial_mpc_b200.*"]),
    package_data={"dial_mpc_b200": ["models/*.json", "examples/*.yaml", "examples/custom_env/*",
                                    "csrc/*.so", "csrc/*.cu", "csrc/*.cuh", "csrc/*.h"]},
    install_requires=["numpy", "torch", "pyyaml"],
    entry_points={"console_scripts": [
        "dial-mpc=dial_mpc_b200.core.dial_core:main",
        "dial-mpc-plan=dial_mpc_b200.deploy.dial_plan:main",
    ]},
)
