"""Face -> cluster -> pupils -> 15 landmark points for a batch of frames (BASELINE.json configs[4]).

The sequencing (which seeds, which cascades, which flips) is the CALLER's logic in the reference --
core/flploc_test.go:75-154 and cmd/pigo/main.go:416-563 -- so it lives here on the host side above the C-ABI and uses
only mirrored API calls: RunCascadeBatch, ClusterDetections, RunDetector (batched over seeds), GetLandmarkPoint.
`randoms_for(frame, call_index)` (optional) injects the [63][3] perturbation randoms of each RunDetector call so that a
CPU oracle can replay the exact same pipeline; without it the library's counter-based generator is used."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import CascadeParams, ImageParams, Pigo, Puploc, PuplocCascade

EYE_CASCADES = ["lp46", "lp44", "lp42", "lp38", "lp312"]      # core/flploc_test.go:77
MOUTH_CASCADES = ["lp93", "lp84", "lp82", "lp81"]             # core/flploc_test.go:78


@dataclass
class Face:
    det: tuple                   # clustered (row, col, scale, q)
    left_eye: Optional[Puploc] = None
    right_eye: Optional[Puploc] = None
    landmarks: List[Puploc] = field(default_factory=list)   # 15 points in the reference test's call order


def eye_seeds(row: int, col: int, scale: int, perturbs: int):
    """core/flploc_test.go:103-118 (float32 arithmetic, int() truncation)."""
    f32 = np.float32
    r = row - int(f32(0.075) * f32(scale))
    sc = float(f32(scale) * f32(0.25))
    left = Puploc(r, col - int(f32(0.175) * f32(scale)), sc, perturbs)
    right = Puploc(r, col + int(f32(0.185) * f32(scale)), sc, perturbs)
    return left, right


def landmark_calls():
    calls = [(e, f) for e in EYE_CASCADES for f in (False, True)]
    calls += [(m, False) for m in MOUTH_CASCADES] + [("lp84", True)]
    return calls


def detect_batch(clf: Pigo, plc: PuplocCascade, flpcs: Dict[str, PuplocCascade], frames: np.ndarray, cp: CascadeParams,
                 iou: float = 0.1, min_face: int = 50, eye_perturbs: int = 50, flp_perturbs: int = 63,
                 randoms_for: Optional[Callable[[int, int], np.ndarray]] = None) -> List[List[Face]]:
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    dets, cnt = clf.RunCascadeBatch(frames, cp, 0.0)
    out: List[List[Face]] = []
    for f in range(frames.shape[0]):
        img = ImageParams(frames[f], cp.ImageParams.Rows, cp.ImageParams.Cols, cp.ImageParams.Dim)
        _, clusters = clf.cluster_array(dets[f, :cnt[f]].copy(), iou)
        faces, call = [], 0
        for c in clusters:
            face = Face((int(c["row"]), int(c["col"]), int(c["scale"]), float(c["q"])))
            if c["scale"] > min_face:
                ls, rs = eye_seeds(int(c["row"]), int(c["col"]), int(c["scale"]), eye_perturbs)
                rl = randoms_for(f, call) if randoms_for else None
                rr = randoms_for(f, call + 1) if randoms_for else None
                rnd = np.stack([rl, rr]) if randoms_for else None
                face.left_eye, face.right_eye = plc.run_detector_batch([ls, rs], img, 0.0, [False, False], rnd, rng_seed=1000 * f + call)
                call += 2
                for name, flip in landmark_calls():
                    rnd = randoms_for(f, call) if randoms_for else None
                    face.landmarks.append(flpcs[name].GetLandmarkPoint(face.left_eye, face.right_eye, img, flp_perturbs, flip,
                                                                      randoms=rnd, rng_seed=1000 * f + call))
                    call += 1
            faces.append(face)
        out.append(faces)
    return out
