"""Cost model behind profiles/r02_allegro_lockstep_model.md: what different ways of synchronising the 14 warps
of a CTA would cost on the dense (Allegro) path, evaluated on the iteration traces written by
scripts/allegro_iteration_trace.py (scratch/its_<rows>_<i>.npy, lss_<rows>_<i>.pkl).
    python scripts/allegro_iteration_trace.py 14 0 && python scripts/allegro_schedule_model.py 14 0
Block costs in warp instructions (ncu source table): prepare TP (kinematics ... warm start), Newton iteration
body NB, line-search iteration LS.  A lock-step round lasts as long as its slowest warp."""
import pickle
import sys

import numpy as np

nrows, level = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (14, 0)
its = np.load(f"scratch/its_{nrows}_{level}.npy")
lss = pickle.load(open(f"scratch/lss_{nrows}_{level}.pkl", "rb"))
W, nsub = min(14, nrows), its.shape[1]
rows = range(W)


def today(TP, NB, LS):                       # barrier per Newton iteration
    return sum(TP + sum(max(NB + LS * lss[r][s][j] for r in rows if its[r, s] > j) for j in range(its[:W, s].max()))
               for s in range(nsub))


def per_substep(TP, NB, LS):                 # barrier per substep, Newton loop free
    return sum(TP + max(sum(NB + LS * l for l in lss[r][s]) for r in rows) for s in range(nsub))


def queues(TP, NB, LS, chunk=None):
    out = []
    for r in rows:
        q = []
        for s in range(nsub):
            q.append(TP)
            for j in range(its[r, s]):
                if chunk is None:
                    q.append(NB + LS * lss[r][s][j])
                else:
                    n, first = lss[r][s][j], True
                    while True:
                        k = min(n, chunk)
                        n -= k
                        q.append((NB if first else 0) + LS * k)
                        first = False
                        if n == 0:
                            break
        out.append(q)
    return out


def rounds(qs):                              # one block per warp and round, whatever block it is
    L = max(len(q) for q in qs)
    return sum(max(q[i] for q in qs if i < len(q)) for i in range(L))


def sliced(TP, NB, LS, K, BAR=300):          # line search in slices of K iterations, prepare stays per substep
    tot = 0
    for s in range(nsub):
        tot += TP
        qs = []
        for r in rows:
            q = []
            for j in range(its[r, s]):
                n, first = lss[r][s][j], True
                while True:
                    k = min(n, K)
                    n -= k
                    q.append((NB if first else 0) + LS * k)
                    first = False
                    if n == 0:
                        break
            qs.append(q)
        L = max((len(q) for q in qs), default=0)
        tot += sum(max(q[i] for q in qs if i < len(q)) + BAR for i in range(L))
    return tot


print(f"Newton iterations per substep: mean {its[:W].mean():.2f}, slowest of {W} warps {its[:W].max(0).mean():.2f}")
allls = [v for r in rows for s_ in lss[r] for v in s_]
mean_ls = np.mean([sum(sum(x) for x in lss[r]) for r in rows])
wait_ls = sum(max(lss[r][s][j] for r in rows if its[r, s] > j) for s in range(nsub) for j in range(its[:W, s].max()))
print(f"line-search iterations: {np.mean(allls):.1f} per Newton iteration, {np.mean(np.array(allls) >= 50):.1%} use all 50; "
      f"per warp {mean_ls:.0f}, waited for by the CTA {wait_ls} ({wait_ls / mean_ls:.1f}x)")
for TP, NB, LS in ((9000, 3500, 250), (6000, 3500, 250), (9000, 5000, 150), (12000, 3000, 300)):
    t = today(TP, NB, LS)
    qs = queues(TP, NB, LS)
    print(f"TP {TP} NB {NB} LS {LS}: per-substep barrier {per_substep(TP, NB, LS) / t:.2f}, mixed rounds {rounds(qs) / t:.2f}, "
          f"LS slices of 8/16 {sliced(TP, NB, LS, 8) / t:.2f}/{sliced(TP, NB, LS, 16) / t:.2f}, "
          f"independent warps {max(sum(q) for q in qs) / t:.2f} (of today's time)")
