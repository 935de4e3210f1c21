"""Example configs (same names as dial_mpc/examples/__init__.py for the in-scope envs)."""
examples = [
    "unitree_h1_jog",
    "unitree_go2_trot",
    "unitree_go2_seq_jump",
    "allegro_reorient",
    "unitree_h1_loco",
]
# deploy (dial-mpc-plan) configurations of the reference's examples directory
deploy_examples = [
    "unitree_go2_trot_deploy",
    "unitree_go2_seq_jump_deploy",
    "unitree_h1_loco_deploy",
]
