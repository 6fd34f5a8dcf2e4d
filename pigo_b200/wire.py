"""The JSON wire format of the reference CLI (`pigo -json`; cmd/pigo/main.go:88-100, :394-398, :452-456, :566-570), so that
switching the backend is invisible to consumers of that output.  Mirrors go/pigo/wire.go.

Quirks reproduced on purpose:
  * `coord{Row "x", Col "y", Scale "size"}` with `omitempty`: zero values are dropped from the object;
  * the CLI stores the COLUMN-derived value in Row and the ROW-derived value in Col: face box
    `Col: face.Row - face.Scale/2, Row: face.Col - face.Scale/2` (main.go:394-398), points `Col: p.Row, Row: p.Col`;
  * eyes / landmark points accumulate over the faces of an image (the slices live outside the loop, main.go:364-366), so
    detection i lists the points of faces 0..i;
  * only faces with Q > 5.0 are reported (main.go:360,:370); points only when Row > 0 && Col > 0;
  * Go's encoding/json: compact separators, struct field order, an all-zero `face` still appears as {} (omitempty never
    drops a struct), empty slices are dropped."""
from __future__ import annotations

import json
from typing import List

Q_THRESH = 5.0


def _coord(row: int, col: int, scale: int) -> dict:
    d = {}
    if row != 0:
        d["x"] = int(row)
    if col != 0:
        d["y"] = int(col)
    if scale != 0:
        d["size"] = int(scale)
    return d


def _trunc_div2(v: int) -> int:
    return int(v / 2)   # Go integer division truncates toward zero


def wire_detections(faces) -> List[dict]:
    """faces: the List[Face] of one image from pipeline.detect_batch / detect_batch_device (Face.det = (row, col, scale, q))."""
    import numpy as np
    dets, eyes, lms = [], [], []
    for f in faces:
        row, col, scale, q = f.det
        if not (np.float32(q) > np.float32(Q_THRESH)):
            continue
        face = _coord(col - _trunc_div2(scale), row - _trunc_div2(scale), scale)     # Row <- Col-derived, Col <- Row-derived
        if f.left_eye is not None:
            for e in (f.left_eye, f.right_eye):
                if e.Row > 0 and e.Col > 0:
                    eyes.append(_coord(e.Col, e.Row, int(e.Scale)))
            for p in f.landmarks:
                if p.Row > 0 and p.Col > 0:
                    lms.append(_coord(p.Col, p.Row, int(p.Scale)))
        d = {}
        if eyes:
            d["eyes"] = list(eyes)
        if lms:
            d["landmark_points"] = list(lms)
        d["face"] = face
        dets.append(d)
    return dets


def marshal_wire(faces) -> str:
    return json.dumps(wire_detections(faces), separators=(",", ":"))
