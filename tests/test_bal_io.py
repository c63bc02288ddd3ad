"""BAL loader / problem construction (host only, no device): the reference's own loader tests
(crates/apex-io/src/bal.rs:396-800) ported onto apex_bal_load, an independent pure-Python parse of the same text,
write -> load round trips, and the CLI's problem construction (bin/bundle_adjustment.rs:212-441) against the generator's
problem."""
import os
import subprocess

import numpy as np
import pytest

from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.bal import BalDataset, dataset_from_problem, load_bal

MINIMAL = "1 1 1\n0 0 -123.456 456.789\n" + "\n".join(str(v) for v in [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 500.0, -0.1, 0.05]) + "\n1\n2\n3\n"


def write(tmp_path, text, name="p.txt"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def with_focal(f):
    return "1 1 1\n0 0 0.0 0.0\n" + "\n".join(str(v) for v in [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, f, 0.0, 0.0]) + "\n0\n0\n0\n"


def test_load_minimal_dataset_and_values(tmp_path):  # bal.rs:432-481
    ds = load_bal(write(tmp_path, MINIMAL))
    assert ds.cameras.shape == (1, 9) and ds.points.shape == (1, 3) and ds.obs_cam.shape == (1,)
    assert np.abs(ds.cameras[0] - [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 500.0, -0.1, 0.05]).max() < 1e-12
    assert (ds.obs_cam[0], ds.obs_pt[0]) == (0, 0)
    assert abs(ds.obs_uv[0, 0] + 123.456) < 1e-10 and abs(ds.obs_uv[0, 1] - 456.789) < 1e-10
    assert np.abs(ds.points[0] - [1.0, 2.0, 3.0]).max() < 1e-12


@pytest.mark.parametrize("focal,expect", [(-100.0, 500.0), (0.0, 500.0), (300.0, 300.0), (float("inf"), 500.0), (float("nan"), 500.0), (1e-300, 1e-300)])
def test_normalize_focal_length(tmp_path, focal, expect):  # bal.rs:483-514, :107-113
    ds = load_bal(write(tmp_path, with_focal(focal)))
    assert ds.cameras[0, 6] == expect


ERROR_CASES = [  # (text, status, fragment of the IoError Display text)  bal.rs:516-800
    ("", F.ERR_PARSE, "Parse error at line 1: Missing header line"),
    ("\n\n   \n", F.ERR_PARSE, "Missing header line"),
    ("1 1\n", F.ERR_MISSING_FIELDS, "Missing required fields at line 1"),
    ("1 abc 1\n", F.ERR_INVALID_NUMBER, "Invalid number format at line 1: abc"),
    ("bad 1 1\n", F.ERR_INVALID_NUMBER, "line 1: bad"),
    ("1 1 bad\n", F.ERR_INVALID_NUMBER, "line 1: bad"),
    ("1 1 -1\n", F.ERR_INVALID_NUMBER, "line 1: -1"),
    ("1 1 2\n0 0 1.0 1.0\n", F.ERR_PARSE, "Unexpected end of file in observations section"),
    ("1 1 1\n0 1.0\n", F.ERR_MISSING_FIELDS, "line 2"),
    ("1 1 1\n0 0 1.0 2.0 3.0\n", F.ERR_MISSING_FIELDS, "line 2"),
    ("1 1 1\nbad 0 1.0 2.0\n", F.ERR_INVALID_NUMBER, "line 2: bad"),
    ("1 1 1\n0 bad 1.0 2.0\n", F.ERR_INVALID_NUMBER, "line 2: bad"),
    ("1 1 1\n0 0 bad_x 1.0\n", F.ERR_INVALID_NUMBER, "line 2: bad_x"),
    ("1 1 1\n0 0 1.0 bad\n", F.ERR_INVALID_NUMBER, "line 2: bad"),
    ("1 1 1\n0 0 1.0 1.0\n0.1\n0.2\n0.3\n", F.ERR_PARSE, "Unexpected end of file in camera 0 parameter 3"),
    ("1 1 1\n0 0 1.0 1.0\nbad\n", F.ERR_INVALID_NUMBER, "line 3: bad"),
    ("1 1 1\n\n0 0 1.0 1.0\n\n0\n0\n0\n0\n0\n0\n1 2\n", F.ERR_INVALID_NUMBER, "line 11: 1 2"),   # line numbers count blank lines
    ("1 1 1\n0 0 1.0 1.0\n" + "0\n" * 9 + "1\n2\n", F.ERR_PARSE, "Unexpected end of file in point 0 coordinate 2"),
    ("1 1 1\n0 0 1.0 1.0\n" + "0\n" * 9 + "1\n2\nzz\n", F.ERR_INVALID_NUMBER, "line 14: zz"),
    ("1 1 1\n0 0 0x10 1.0\n", F.ERR_INVALID_NUMBER, "0x10"),                      # Rust's f64 parser has no hex floats
    ("1 1 1\n5 0 1.0 1.0\n", F.ERR_PARSE, "observation index out of range"),       # would panic in the reference CLI (:405)
]


@pytest.mark.parametrize("text,status,frag", ERROR_CASES)
def test_loader_errors(tmp_path, text, status, frag):
    with pytest.raises(F.ApexError) as e:
        load_bal(write(tmp_path, text))
    assert e.value.status == status and frag in str(e.value)


def test_load_nonexistent_file():  # bal.rs:516-520
    with pytest.raises(F.ApexError) as e:
        load_bal("/nonexistent/path/file.bal")
    assert e.value.status == F.ERR_IO


def test_load_multiple_cameras_points_blank_lines_and_whitespace(tmp_path):  # bal.rs:576-606, :144-149
    text = "  2 2 3  \n\n0 0 1.0 1.0\n0 1\t2.0   2.0\n\n1 0 3.0 3.0\n" + "\n".join(["0.0"] * 6 + ["100.0", "0.0", "0.0"]) + "\n\n" \
           + "\n".join(["0.0"] * 6 + [" 200.0 ", "1e-3", "-2.5E-4"]) + "\n1\n2\n3\n\n4\n5\n6"
    ds = load_bal(write(tmp_path, text))
    assert ds.cameras.shape == (2, 9) and ds.points.shape == (2, 3) and ds.obs_cam.tolist() == [0, 0, 1] and ds.obs_pt.tolist() == [0, 1, 0]
    assert ds.cameras[0, 6] == 100.0 and ds.cameras[1, 6:].tolist() == [200.0, 1e-3, -2.5e-4]
    assert ds.points.tolist() == [[1, 2, 3], [4, 5, 6]] and ds.obs_uv.tolist() == [[1, 1], [2, 2], [3, 3]]


def python_parse(text):
    """Independent restatement of the format (bal.rs:138-400) in plain Python."""
    lines = [l.strip() for l in text.splitlines() if l.strip()]
    ncam, npts, nobs = (int(x) for x in lines[0].split())
    obs = [l.split() for l in lines[1:1 + nobs]]
    cams = np.array([float(x) for x in lines[1 + nobs:1 + nobs + 9 * ncam]]).reshape(ncam, 9)
    pts = np.array([float(x) for x in lines[1 + nobs + 9 * ncam:1 + nobs + 9 * ncam + 3 * npts]]).reshape(npts, 3)
    return cams, pts, np.array([int(o[0]) for o in obs]), np.array([int(o[1]) for o in obs]), np.array([[float(o[2]), float(o[3])] for o in obs])


def test_write_load_round_trip_is_bit_exact_and_matches_python_parse(tmp_path):
    prob = synth.make_problem(9, 300, 4.0, seed=17)
    ds = dataset_from_problem(prob)
    path = str(tmp_path / "problem-9-300-pre.txt")
    ds.write(path)
    back = load_bal(path)
    for a, b in ((ds.cameras, back.cameras), (ds.points, back.points), (ds.obs_uv, back.obs_uv)):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert np.array_equal(ds.obs_cam, back.obs_cam) and np.array_equal(ds.obs_pt, back.obs_pt)
    cams, pts, oc, op, uv = python_parse(open(path).read())
    assert np.array_equal(cams, back.cameras) and np.array_equal(pts, back.points) and np.array_equal(uv, back.obs_uv)
    assert np.array_equal(oc, back.obs_cam) and np.array_equal(op, back.obs_pt)


@pytest.mark.parametrize("opt_type", ["self-calibration", "bundle-adjustment"])
def test_problem_construction_matches_cli_semantics(opt_type):  # bin/bundle_adjustment.rs:212-441
    prob = synth.make_problem(7, 200, 4.0, seed=23)
    ds = dataset_from_problem(prob)
    p = ds.problem(optimization_type=opt_type)
    assert p.camera_model == F.CAM_BAL and p.loss_id == F.LOSS_HUBER and p.loss_params[0] == 1.0 and p.intr_vars_present
    assert p.opt_flags == (F.OPT_POSE | F.OPT_LANDMARK | (F.OPT_INTRINSIC if opt_type == "self-calibration" else 0))
    assert p.pose_fixed[0] == 0x3F and not p.pose_fixed[1:].any()                      # pose_0000, all 6 DOF (:294-298)
    assert np.array_equal(p.pose[:, :3], prob.pose[:, :3]) and np.array_equal(p.intr, prob.intr) and np.array_equal(p.pt, prob.pt)
    # axis-angle -> quaternion (:200-208): same rotation as the generator's pose, unit norm, identity below 1e-10
    assert np.abs(np.abs((p.pose[:, 3:] * prob.pose[:, 3:]).sum(1)) - 1.0).max() < 1e-14
    assert np.abs(np.linalg.norm(p.pose[:, 3:], axis=1) - 1.0).max() < 1e-15
    tiny = BalDataset(cameras=[[1e-11, 0, 0, 1, 2, 3, 400, 0, 0]], points=[[0, 0, -5]], obs_cam=[0], obs_pt=[0], obs_uv=[[0, 0]]).problem()
    assert tiny.pose[0].tolist() == [1, 2, 3, 1, 0, 0, 0]
    # -n: landmarks 0..n-1, observations with point_index < n in file order (:170-171, :258-262)
    q = ds.problem(num_points=50, optimization_type=opt_type)
    keep = prob.obs_pt < 50
    assert q.npts == 50 and np.array_equal(q.obs_pt, prob.obs_pt[keep]) and np.array_equal(q.obs_cam, prob.obs_cam[keep]) and np.array_equal(q.obs_uv, prob.obs_uv[keep])
    assert ds.problem(num_points=10**9).npts == prob.npts
    for dead in ("only-pose", "only-landmarks", "only-intrinsics"):
        with pytest.raises(F.ApexError) as e:
            ds.problem(optimization_type=dead)
        assert e.value.status == F.ERR_UNSUPPORTED


def test_cli_without_a_device_fails_loudly(tmp_path):
    """The CLI binary parses the reference's flags and loads the file on the host; without a CUDA device the run ends
    with an error (no CPU fallback). On a GPU box the same command optimises (tests/test_gpu_parity.py)."""
    exe = os.path.join(os.path.dirname(F.LIB_PATH), "bundle_adjustment")
    assert os.path.exists(exe), "built by make -C apex_solver_b200/csrc"
    path = str(tmp_path / "problem-6-150-pre.txt")
    dataset_from_problem(synth.make_problem(6, 150, 4.0, seed=29)).write(path)
    r = subprocess.run([exe, path, "-n", "100", "-s", "explicit", "-t", "bundle-adjustment"], capture_output=True, text=True, timeout=120)
    assert "Cameras: 6" in r.stdout and "Points to use: 100" in r.stdout and "Solver variant: explicit" in r.stdout
    if F.load_library().apex_device_count() <= 0:
        assert r.returncode == 1 and "cannot create a GPU context" in r.stderr
    else:
        assert r.returncode == 0 and "Optimization completed!" in r.stdout
    assert subprocess.run([exe, path, "-s", "nope"], capture_output=True, text=True).returncode == 2
    assert subprocess.run([exe, str(tmp_path / "missing.txt")], capture_output=True, text=True).returncode == 1
    assert subprocess.run([exe], capture_output=True, text=True).returncode == 2


def test_header_counts_larger_than_the_file_fail_cleanly(tmp_path):
    """An untrusted header (4e9 observations / cameras / points in a 30-byte file) must end in the loader's end-of-file
    errors, not in a 100 GB allocation or a C++ exception crossing the C ABI."""
    from apex_solver_b200 import _ffi as F
    from apex_solver_b200.bal import load_bal
    for text, what in (("3 3 4000000000\n0 0 1.0 2.0\n", "observations section"), ("4000000000 1 1\n0 0 1.0 2.0\n1.0\n", "camera 0 parameter 1"),
                       ("1 4000000000 0\n" + "1.0\n" * 9 + "2.0\n", "point 0 coordinate 1")):
        path = tmp_path / "hostile.txt"
        path.write_text(text)
        with pytest.raises(F.ApexError) as e:
            load_bal(str(path))
        assert e.value.status == F.ERR_PARSE and "Unexpected end of file in " + what in str(e.value), str(e.value)
