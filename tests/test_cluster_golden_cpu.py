"""CPU: the start-pose grid (foundationpose_b200/hypotheses.py: sample_views_icosphere, make_rotation_grid, cluster_poses)
against golden vectors produced by the reference's own code (tools/make_golden_cluster.py): the C++ `cluster_poses` +
`rotationGeodesicDistance` compiled from the reference tree (oracle/build_ref.py, Eigen replaced by oracle/eigen_shim.h),
called from the reference's unmodified `FoundationPose.make_rotation_grid` (estimater.py:106-124) and
`sample_views_icosphere` (Utils.py:483-507).  Unpinned by construction: trimesh's icosphere vertex ORDER (trimesh is absent;
both sides use this repository's icosphere)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden_cluster as gen  # symmetry_sets() only

    return gen, dict(np.load(os.path.join(ROOT, "tests", "golden", "cluster_golden.npz")))


def test_views_and_rotation_grid_match_the_reference_methods(golden):
    from foundationpose_b200 import hypotheses as hy

    gen, g = golden
    assert np.abs(hy.sample_views_icosphere(40) - g["views_40"]).max() < 1e-12
    assert np.abs(hy.sample_views_icosphere(1, subdivisions=2, radius=0.5) - g["views_sub2"]).max() < 1e-12
    for name, syms in gen.symmetry_sets().items():
        got = hy.make_rotation_grid(40, 60, syms)
        want = g[f"rot_grid.{name}"]
        assert got.shape == want.shape, (name, got.shape, want.shape)      # 252 / 126 / 20 / 63 start poses
        assert np.abs(got - want).max() < 1e-6, name
    assert np.abs(hy.make_rotation_grid(10, 90, None) - g["rot_grid.identity_10_90"]).max() < 1e-6


def test_cluster_poses_matches_the_reference_cpp(golden):
    from foundationpose_b200 import hypotheses as hy

    gen, g = golden
    grid = g["rot_grid.identity"]
    for name, syms in gen.symmetry_sets().items():
        for ang in (10, 61):
            assert np.array_equal(hy.cluster_poses(ang, 99999, grid, syms), g[f"cluster.{name}.{ang}"]), (name, ang)
    assert np.array_equal(hy.cluster_poses(30, 0.01, g["moved_poses"], gen.symmetry_sets()["half_z"]), g["cluster.moved.half_z"])


def test_cluster_poses_against_the_compiled_reference_function(golden):
    """The same, live against oracle/_ref/libcluster_ref.so (built by __graft_entry__.build() where /root/reference
    exists; it travels to the GPU box with the snapshot) on random pose sets the fixture does not hold."""
    from foundationpose_b200 import hypotheses as hy
    from oracle import build_ref

    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref/libcluster_ref.so not built (no reference tree)")
    gen, g = golden
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(3)
    for trial in range(4):
        n = 150
        poses = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
        poses[:, :3, :3] = Rotation.random(n, random_state=trial).as_matrix()
        poses[:, :3, 3] = rng.normal(0, 0.02, (n, 3))
        for name, syms in gen.symmetry_sets().items():
            a = hy.cluster_poses(25 + 5 * trial, 0.03, poses, syms)
            b = ref(25 + 5 * trial, 0.03, poses, syms)
            assert a.shape == b.shape and np.array_equal(a, b), (trial, name, a.shape, b.shape)
