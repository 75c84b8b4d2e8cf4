"""CPU: host-side pieces of the synthetic-sequence pipeline (BASELINE config 5).
 * the drop-in PointCloudVector::update (hash index) against the reference's first-match /
   append semantics (PointCloud.h:52-66);
 * the sequence driver compiled against the REFERENCE's headers and sources
   (oracle/_ref/bin/vo_sequence): the reference pipeline must track the generator's ground truth."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_pointcloud_update_hash_index_matches_reference_semantics(tmp_path):
    exe = str(tmp_path / "pointcloud_update_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17",
                           "-I", os.path.join(ROOT, "visual-odometry_b200", "host", "include"),
                           "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "third_party", "mini_eigen"),
                           os.path.join(HERE, "pointcloud_update_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout


def test_reference_pipeline_tracks_ground_truth_on_a_synthetic_sequence(tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "vo_sequence")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bin/vo_sequence not built (oracle/build_ref.sh needs the reference checkout)")
    out = subprocess.run([exe, "1500", "24", "1000", "100", str(tmp_path / "poses.txt")],
                         capture_output=True, text=True, check=True).stdout
    r = json.loads(out.strip().splitlines()[-1])
    assert r["impl"] == "reference-cpu" and r["frames"] == 20
    assert r["mean_correspondences"] > 80
    assert r["rot_err_mean_rad"] < 1e-3
    # monocular scale is arbitrary but must stay the one fixed by the first pair
    assert abs(r["scale_median"] / r["scale_first_pair"] - 1.0) < 1e-3
