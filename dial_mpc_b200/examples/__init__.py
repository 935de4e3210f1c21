"""Example configs (same names as dial_mpc/examples/__init__.py for the in-scope envs)."""
examples = [
    "unitree_h1_jog",
    "unitree_go2_trot",
    "unitree_go2_seq_jump",
    "allegro_reorient",
    "unitree_h1_loco",
]
