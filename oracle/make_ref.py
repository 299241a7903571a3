"""Stage the UNMODIFIED reference env modules where the GPU box can run them: oracle/_ref/ (git-ignored,
NOT gpurun-ignored — it travels with the snapshot like a built .so; nothing of it enters the history).

The reference is pure Python + numba (no build step): "building" it for the CPU baseline means copying
beacon/<env>/<env>.py and its init_field.dat next to the import shims.  `bench.py --impl reference` and the
cpu_baseline leg then time the real reference (`kind: "reference"`) beside the C port (`kind: "port"`).

    python oracle/make_ref.py            (called by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("BEACON_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
ENVS = ("shkadov", "rayleigh", "mixing", "burgers", "sloshing", "lorenz", "vortex")


def stage():
    if not os.path.isdir(os.path.join(SRC, "beacon")):
        return None
    for e in ENVS:
        d = os.path.join(DST, "beacon", e)
        os.makedirs(d, exist_ok=True)
        for f in (e + ".py", "init_field.dat"):
            s = os.path.join(SRC, "beacon", e, f)
            if os.path.exists(s):
                shutil.copyfile(s, os.path.join(d, f))
    for f in ("LICENSE",):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return DST


if __name__ == "__main__":
    print(stage() or "reference checkout not present: nothing staged", file=sys.stderr)
