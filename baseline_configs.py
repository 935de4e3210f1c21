"""The five workloads of BASELINE.json (`configs`), shared by bench.py and the tests.

Planner fields that BASELINE.json does not name come from the reference's example YAML of the
same env; the env configuration is the example's too.  `N` is the sample
count of ONE GPU (configs[4]: 65536 samples sharded over 8 GPUs = 8192 per GPU)."""

ENV_CFG = {
    "unitree_go2_walk": dict(default_vx=0.8, ramp_up_time=1.0),                       # unitree_go2_trot.yaml
    "unitree_go2_seq_jump": dict(
        pose_target_sequence=[[0, 0, 0.27], [0.4, 0, 0.27], [0.8, 0, 0.27], [1.2, 0, 0.27], [1.6, 0, 0.27]],
        yaw_target_sequence=[0.0] * 5),                                              # unitree_go2_seq_jump.yaml
    "unitree_h1_walk": dict(default_vx=2.0, ramp_up_time=3.0),                        # unitree_h1_jog.yaml (gait jog)
    "allegro_reorient": dict(dt=0.02, timestep=0.005, leg_control="position"),        # allegro_reorient.yaml
    "unitree_h1_loco": dict(default_vx=0.6, ramp_up_time=3.0, gait="walk"),           # unitree_h1_loco.yaml
}

BASELINE = {
    0: dict(name="unitree_go2_trot", env="unitree_go2_walk", N=128, Hs=16, Hn=4, Ndiffuse=2, temp=0.05, hdf=0.9, tdf=0.5,
            gpus=1, note="configs[0]: the reference's CPU-runnable case"),
    1: dict(name="unitree_go2_seq_jump", env="unitree_go2_seq_jump", N=2048, Hs=25, Hn=5, Ndiffuse=4, temp=0.05, hdf=0.9,
            tdf=0.5, gpus=1, note="configs[1]: the headline"),
    2: dict(name="unitree_h1_jog", env="unitree_h1_walk", N=2048, Hs=30, Hn=5, Ndiffuse=6, temp=0.05, hdf=1.0, tdf=0.5,
            gpus=1, note="configs[2]"),
    3: dict(name="allegro_reorient", env="allegro_reorient", N=4096, Hs=20, Hn=4, Ndiffuse=6, temp=0.05, hdf=1.0, tdf=0.5,
            gpus=1, note="configs[3]: 4 physics substeps per env step"),
    4: dict(name="unitree_go2_trot_65536", env="unitree_go2_walk", N=8192, Hs=25, Hn=4, Ndiffuse=4, temp=0.05, hdf=0.9,
            tdf=0.5, gpus=8, note="configs[4]: 65536 samples over 8 GPUs = 8192 per GPU"),
}


def dial_config(i: int, world: int = 1):
    from dial_mpc_b200.core.dial_config import DialConfig
    b = BASELINE[i]
    return DialConfig(env_name=b["env"], Nsample=b["N"] * world, Hsample=b["Hs"], Hnode=b["Hn"], Ndiffuse=b["Ndiffuse"],
                      temp_sample=b["temp"], horizon_diffuse_factor=b["hdf"], traj_diffuse_factor=b["tdf"])


def product_env(env_name: str):
    import numpy as np
    import dial_mpc_b200.envs as E
    cfg_t = E.get_config(env_name)
    ecfg = cfg_t(**{k: (np.array(v) if isinstance(v, list) else v) for k, v in ENV_CFG[env_name].items()})
    return E.get_environment(env_name, config=ecfg)
