"""Offline sweep for the compute-share controller: random watcher trajectories produced by the REFERENCE's own code
(oracle/_ref/ref_cosim, through golden/generate.py's watcher_traj) and replayed through the oracle restatement
(oracle/vgpu_oracle.c orc_fold_utilization + orc_watcher_step) - share, bucket, up_limit, valid, user and sys
utilisation must agree at every control step.  `python tests/fuzz_watcher.py SEED TRAJECTORIES [STEPS]`.
Not collected by pytest (test_oracle_parity.py runs a small seeded slice of it where oracle/_ref exists)."""
import ctypes as C
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import helpers as H  # noqa: E402
import generate as G  # noqa: E402

def random_trajectory(rng, steps, name="fuzz"):
    return G.random_trajectory(rng, steps, name)


def replay(o, tr):
    """First step at which the oracle leaves the reference's trajectory, or None."""
    import test_oracle_parity as T
    g = T._gpu(o, tr["sm"], tr["thr"])
    cfg = H.CfgDev(hard_core=tr["hard"], soft_core=tr["soft"], core_limit=tr["core_limit"], hard_limit=tr["hard_limit"])
    w = H.OrcWatcher()
    o.orc_watcher_init(C.byref(w), C.byref(cfg))

    def step(util, bucket_in):
        b = C.c_int64(bucket_in)
        o.orc_watcher_step(C.byref(g), C.byref(cfg), C.byref(w), C.byref(util), C.byref(b))
        return w.share, b.value, w.up_limit

    try:
        T.replay_watcher(o, tr, step)
    except AssertionError as e:
        return str(e)[:600]
    return None


def main():
    seed, count = int(sys.argv[1], 0), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    H.build_all()
    if not H.have_reference():
        sys.exit("oracle/_ref is missing: run this where /root/reference exists")
    rng = random.Random(seed)
    o = H.oracle()
    bad = 0
    for i in range(count):
        tr = random_trajectory(rng, steps, "fuzz%d" % i)
        why = replay(o, tr)
        if why:
            bad += 1
            print("trajectory %d (mode %d hard %d soft %d geom %dx%d) diverges: %s" % (i, tr["mode"], tr["hard"], tr["soft"], tr["sm"], tr["thr"], why), flush=True)
        if i % 20 == 19:
            print("trajectories", i + 1, "divergent", bad, flush=True)
    print("done: %d trajectories x %d steps, %d divergent" % (count, steps, bad))


if __name__ == "__main__":
    main()
