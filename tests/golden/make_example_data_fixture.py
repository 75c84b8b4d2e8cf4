"""Packs the reference's bundled dataset (config 1 of BASELINE.json) and the UNMODIFIED
reference's output on it into fixtures, so the GPU box (which has no /root/reference) can run
vo_complete on the same input and compare.

  tests/golden/example_data.tar.gz          = /root/reference/example_data/data  (121 frames)
  tests/golden/ref_vo_complete_outputs.npz  = trajectory_est_complete / trajectory_est_data /
                                              map produced by oracle/_ref/bin/vo_complete (the
                                              reference's own main compiled against mini_eigen,
                                              oracle/build_ref.sh), plus evaluation's summary, plus
                                              the trajectory of the SAME reference sources built
                                              with FMA contraction (-march=native
                                              -ffp-contract=fast): the pipeline amplifies rounding
                                              differences exponentially along the 121 frames, and
                                              this second CPU build is the yardstick for how far
                                              two correct FP32 implementations drift apart.
Run in the authoring container:  python tests/golden/make_example_data_fixture.py
"""
import os
import subprocess
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATA = "/root/reference/example_data/data"
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def main():
    with tarfile.open(os.path.join(HERE, "example_data.tar.gz"), "w:gz") as tf:
        tf.add(DATA, arcname="data")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([os.path.join(BIN, "vo_complete"), DATA], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        ev = subprocess.run([os.path.join(BIN, "evaluation"), DATA], cwd=tmp, check=True,
                            capture_output=True, text=True).stdout
        out = {
            "trajectory_est_complete": np.loadtxt(os.path.join(tmp, "trajectory_est_complete.txt")),
            "trajectory_est_data": np.loadtxt(os.path.join(tmp, "trajectory_est_data.txt")),
            "map": np.loadtxt(os.path.join(tmp, "map.txt")),
            "evaluation_stdout": np.array(ev),
        }
    with tempfile.TemporaryDirectory() as tmp:
        ref = "/root/reference"
        srcs = [f"{ref}/src/apps/vo_complete.cpp"] + [f"{ref}/src/{n}.cpp" for n in
                ("picp_solver", "camera", "utils", "epipolar_utils", "files_utils")]
        subprocess.run(["g++", "-std=c++17", "-O3", "-DNDEBUG", "-march=native", "-ffp-contract=fast",
                        "-w", "-I", os.path.join(ROOT, "third_party", "mini_eigen"),
                        "-I", f"{ref}/include", *srcs, "-o", os.path.join(tmp, "vo_fma")], check=True)
        subprocess.run([os.path.join(tmp, "vo_fma"), DATA], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        out["trajectory_est_complete_fma_build"] = np.loadtxt(
            os.path.join(tmp, "trajectory_est_complete.txt"))
    dev = np.abs(out["trajectory_est_complete_fma_build"] - out["trajectory_est_complete"]).max(1)
    print("CPU SSE2 build vs CPU FMA build, max |dt| at frames 1,10,30,60,120:",
          [float(dev[i]) for i in (1, 10, 30, 60, 120)])
    np.savez_compressed(os.path.join(HERE, "ref_vo_complete_outputs.npz"), **out)
    print(ev)
    print({k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
