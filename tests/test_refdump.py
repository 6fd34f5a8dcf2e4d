"""Pins the CPU oracle to TRUE outputs of the Go reference when a dump made by tools/refdump/main.go is present
(oracle/_ref/refdump.json, or a committed copy tests/golden/refdump.json); skipped otherwise (no Go toolchain in this
image or on the GPU box -- see oracle/_ref/README.md).  The consumer itself is exercised on every run by
test_refdump_consumer_on_a_synthetic_dump, which builds a dump in the tool's format from the oracle and checks that the
comparison accepts it and rejects a perturbed copy."""
import copy
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import pigo_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "oracle", "_ref", "refdump.json"), os.path.join(ROOT, "tests", "golden", "refdump.json")]
INPUTS = os.path.join(ROOT, "oracle", "_ref", "inputs")


def _dets(lst):
    a = np.zeros(len(lst), dtype=O.DET_DTYPE)
    for i, d in enumerate(lst):
        a[i] = (d["row"], d["col"], d["scale"], np.array([d["q_bits"]], dtype=np.uint32).view(np.float32)[0])
    return a


def _f32(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


def compare_dump(dump, load_input):
    """Returns a list of mismatch descriptions (empty = the oracle reproduces the dump bit for bit)."""
    bad = []
    face = O.OracleFace(pigo_b200.load_cascade("facefinder"))
    plcs = {}

    def plc(name):
        if name not in plcs:
            plcs[name] = O.OraclePuploc(pigo_b200.load_cascade(name))
        return plcs[name]

    manifest = {m["file"]: m for m in json.load(open(os.path.join(INPUTS, "manifest.json")))} if os.path.exists(os.path.join(INPUTS, "manifest.json")) else {}
    for fo in dump["files"]:
        geo = fo.get("geometry") or manifest[fo["file"]]
        rows, cols, dim = geo["rows"], geo["cols"], geo["dim"]
        img = load_input(fo["file"], rows, dim)
        for run in fo["runs"]:
            p = run["params"]
            want = _dets(run["detections"] or [])
            got = face.run_cascade(img, rows, cols, dim, p["min_size"], p["max_size"], p["shift_factor"], p["scale_factor"], p["angle"])
            if got.tobytes() != want.tobytes():
                bad.append(f"{fo['file']} RunCascade {p}: {len(got)} vs {len(want)} detections or different values")
                continue
            for key, cl in (run.get("clusters_by_iou") or {}).items():
                srt_o, cl_o = O.cluster(got, float(key))
                want_cl, want_srt = _dets(cl or []), _dets(run["sorted_by_iou"][key] or [])
                ties = len(np.unique(got["q"])) != len(got)     # sort.Slice is unstable: tie order is the Go version's business
                if srt_o.tobytes() != want_srt.tobytes() and not ties:
                    bad.append(f"{fo['file']} ClusterDetections({key}) in-place sort differs")
                if cl_o.tobytes() != want_cl.tobytes() and not ties:
                    bad.append(f"{fo['file']} ClusterDetections({key}) clusters differ")
        for pu in fo.get("pupils") or []:
            p = pu["params"]
            rnd = np.zeros(189, dtype=np.float32)
            rnd[:len(pu["randoms_bits"])] = _f32(pu["randoms_bits"])
            e = plc(p["cascade"]).run_detector(p["row"], p["col"], float(np.float32(p["scale"])), p["perturbs"], rnd, img, rows, cols, dim, p["angle"], p["flipv"])
            if (e[0], e[1]) != (pu["row"], pu["col"]) or np.float32(e[2]).view(np.uint32) != np.uint32(pu["scale_bits"]):
                bad.append(f"{fo['file']} RunDetector {p}: oracle {e} vs reference ({pu['row']}, {pu['col']}, bits {pu['scale_bits']})")
        for lm in fo.get("landmarks") or []:
            p = lm["params"]
            rnd = np.zeros(189, dtype=np.float32)
            rnd[:len(lm["randoms_bits"])] = _f32(lm["randoms_bits"])
            r0, c0, s0 = O.landmark_seed(p["left_row"], p["left_col"], p["right_row"], p["right_col"])
            e = plc(p["cascade"]).run_detector(r0, c0, float(s0), p["perturbs"], rnd, img, rows, cols, dim, 0.0, p["flipv"])
            if (e[0], e[1]) != (lm["row"], lm["col"]) or np.float32(e[2]).view(np.uint32) != np.uint32(lm["scale_bits"]):
                bad.append(f"{fo['file']} GetLandmarkPoint {p}: oracle {e} vs reference ({lm['row']}, {lm['col']})")
    return bad


def test_oracle_reproduces_the_go_reference_dump():
    path = next((p for p in CANDIDATES if os.path.exists(p)), None)
    if path is None:
        pytest.skip("no refdump.json: needs a Go toolchain once (oracle/_ref/README.md); the oracle stays 'parity unpinned'")
    dump = json.load(open(path))

    def load(name, rows, dim):
        return np.fromfile(os.path.join(INPUTS, name), dtype=np.uint8)

    if not os.path.exists(os.path.join(INPUTS, "manifest.json")):
        import subprocess
        import sys
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "refdump", "make_inputs.py")])
    assert compare_dump(dump, load) == []


def test_refdump_consumer_on_a_synthetic_dump(sample_gray):
    """The dump format and the comparison logic, exercised without Go: a dump built from the oracle itself must be
    accepted, a copy with one flipped score bit / one moved pupil must be rejected."""
    face = O.OracleFace(pigo_b200.load_cascade("facefinder"))
    prm = {"min_size": 20, "max_size": 1000, "shift_factor": 0.2, "scale_factor": 1.1, "angle": 0.0, "iou": [0.1]}
    d = face.run_cascade(sample_gray, 400, 320, 320, 20, 1000, 0.2, 1.1, 0.0)
    srt, cl = O.cluster(d, 0.1)

    def ser(a):
        return [{"row": int(x["row"]), "col": int(x["col"]), "scale": int(x["scale"]), "q_bits": int(np.float32(x["q"]).view(np.uint32))} for x in a]
    rnd = np.random.default_rng(1).random(189, dtype=np.float32)
    pp = {"cascade": "puploc", "row": 186, "col": 118, "scale": 60.0, "perturbs": 63, "angle": 0.0, "flipv": False, "seed": 1}
    e = O.OraclePuploc(pigo_b200.load_cascade("puploc")).run_detector(186, 118, 60.0, 63, rnd, sample_gray, 400, 320, 320)
    dump = {"go_version": "synthetic", "files": [{
        "file": "sample", "geometry": {"rows": 400, "cols": 320, "dim": 320},
        "runs": [{"params": prm, "detections": ser(d), "sorted_by_iou": {"0.1": ser(srt)}, "clusters_by_iou": {"0.1": ser(cl)}}],
        "pupils": [{"params": pp, "randoms_bits": [int(x) for x in rnd.view(np.uint32)], "row": int(e[0]), "col": int(e[1]),
                    "scale_bits": int(np.float32(e[2]).view(np.uint32))}], "landmarks": []}]}

    def load(name, rows, dim):
        return sample_gray.reshape(-1)
    assert compare_dump(dump, load) == []
    broken = copy.deepcopy(dump)
    broken["files"][0]["runs"][0]["detections"][0]["q_bits"] ^= 1
    broken["files"][0]["pupils"][0]["row"] += 1
    assert len(compare_dump(broken, load)) == 2
