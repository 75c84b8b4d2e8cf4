"""Builds tests/golden/bundled_frames.npz from the reference's bundled dataset
(/root/reference/example_data/data, Appendix C of SURVEY.md for the format).

The reference ships no golden vectors; the only fixed answers it holds for the NN path are the
landmark ids embedded in every meas-*.dat line: two measurements in consecutive frames are a true
match iff their landmark ids are equal (the logic of src/tests/compute_corr.cpp:68-82).  This
script extracts, for a handful of frames, the landmark id, pixel and 10-D appearance of every
measurement, plus world.dat, so the tests can check NN association against id-based truth without
/root/reference being present (it does not exist on the GPU box).

Run in the authoring container:  python tests/golden/make_bundled_fixture.py
"""
import os

import numpy as np

DATA = "/root/reference/example_data/data"
FRAMES = [0, 1, 2, 3, 59, 60, 61, 118, 119, 120]


def read_meas(path):
    ids, uv, app = [], [], []
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w or w[0] != "point":
                continue
            ids.append(int(w[2]))
            uv.append([float(w[3]), float(w[4])])
            app.append([float(x) for x in w[5:15]])
    return np.array(ids, np.int32), np.array(uv, np.float32), np.array(app, np.float32)


def main():
    out = {}
    for k in FRAMES:
        ids, uv, app = read_meas(os.path.join(DATA, f"meas-{k:05d}.dat"))
        out[f"ids_{k}"], out[f"uv_{k}"], out[f"app_{k}"] = ids, uv, app
    world = np.loadtxt(os.path.join(DATA, "world.dat"), dtype=np.float64)
    out["world_ids"] = world[:, 0].astype(np.int32)
    out["world_xyz"] = world[:, 1:4].astype(np.float32)
    out["world_app"] = world[:, 4:14].astype(np.float32)
    out["frames"] = np.array(FRAMES, np.int32)
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "bundled_frames.npz"), **out)
    print("wrote", os.path.join(here, "bundled_frames.npz"), {k: v.shape for k, v in out.items() if k.startswith("ids")})


if __name__ == "__main__":
    main()
