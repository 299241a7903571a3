"""Convert the reference's developed-flow initial states (init_field.dat, text '%.5e')
into the binary fixture beacon_b200/data/init_fields.npz.

The three files are DATA inputs of the path (shkadov 2900x3 (x,h,q),
rayleigh 208x52 (u,v,p,T stacked), sloshing 200x3 (x,h,q)); they are parsed with
np.loadtxt exactly as the reference's load() does (shkadov.py:364-368, rayleigh.py:356-362,
sloshing.py:310-314) and stored as float64 so every consumer sees identical bits.

Run in the build container:  python tools/import_init_fields.py [/root/reference]
"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "beacon_b200", "data", "init_fields.npz")

s = np.loadtxt(os.path.join(ref, "beacon/shkadov/init_field.dat"))
r = np.loadtxt(os.path.join(ref, "beacon/rayleigh/init_field.dat"))
l = np.loadtxt(os.path.join(ref, "beacon/sloshing/init_field.dat"))
n = r.shape[1]
np.savez_compressed(
    out,
    shkadov_h=s[:, 1], shkadov_q=s[:, 2],
    rayleigh_u=r[0 * n:1 * n], rayleigh_v=r[1 * n:2 * n], rayleigh_p=r[2 * n:3 * n], rayleigh_T=r[3 * n:4 * n],
    sloshing_h=l[:, 1], sloshing_q=l[:, 2],
)
print("wrote", os.path.normpath(out), {k: v.shape for k, v in np.load(out).items()})
