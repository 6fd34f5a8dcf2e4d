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
    """Per frame: one batched RunCascade for all frames, ClusterDetections, then ONE RunDetector launch for the eye seeds
    of all faces of the frame and ONE launch per (landmark cascade) for all faces and both flips -- the per-call results
    are identical to calling GetLandmarkPoint face by face (each seed carries its own randoms / RNG key)."""
    from . import landmark_seed_host
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    dets, cnt = clf.RunCascadeBatch(frames, cp, 0.0)
    out: List[List[Face]] = []
    calls = landmark_calls()
    for f in range(frames.shape[0]):
        img = ImageParams(frames[f], cp.ImageParams.Rows, cp.ImageParams.Cols, cp.ImageParams.Dim)
        _, clusters = clf.cluster_array(dets[f, :cnt[f]].copy(), iou)
        faces = [Face((int(c["row"]), int(c["col"]), int(c["scale"]), float(c["q"]))) for c in clusters]
        big = [k for k, c in enumerate(clusters) if c["scale"] > min_face]
        # call indices follow the reference's sequential order: per face 2 eye calls, then the 15 landmark calls
        base = {k: i * (2 + len(calls)) for i, k in enumerate(big)}
        if big:
            seeds, rnds = [], []
            for k in big:
                c = clusters[k]
                ls, rs = eye_seeds(int(c["row"]), int(c["col"]), int(c["scale"]), eye_perturbs)
                seeds += [ls, rs]
                if randoms_for:
                    rnds += [randoms_for(f, base[k]), randoms_for(f, base[k] + 1)]
            eyes = plc.run_detector_batch(seeds, img, 0.0, [False] * len(seeds), np.stack(rnds) if randoms_for else None, rng_seed=1000 * f)
            for i, k in enumerate(big):
                faces[k].left_eye, faces[k].right_eye = eyes[2 * i], eyes[2 * i + 1]
                faces[k].landmarks = [None] * len(calls)
            for name in sorted(set(n for n, _ in calls)):
                seeds, flips, rnds, where = [], [], [], []
                for k in big:
                    for ci, (n, flip) in enumerate(calls):
                        if n != name:
                            continue
                        seeds.append(landmark_seed_host(faces[k].left_eye, faces[k].right_eye, flp_perturbs))
                        flips.append(flip)
                        where.append((k, ci))
                        if randoms_for:
                            rnds.append(randoms_for(f, base[k] + 2 + ci))
                pts = flpcs[name].run_detector_batch(seeds, img, 0.0, flips, np.stack(rnds) if randoms_for else None,
                                                     rng_seed=1000 * f + 7)
                for (k, ci), p in zip(where, pts):
                    faces[k].landmarks[ci] = p
        out.append(faces)
    return out
