// wire.go -- the JSON wire format of the reference CLI (`pigo -json`, cmd/pigo/main.go:88-100, :394-398, :452-456, :566-570),
// so that a backend switch is invisible to consumers of that output (SURVEY.md section 8f row N4).  In the reference these
// types are unexported members of package main; they are restated here with the same field tags and the same quirks:
//   - Row is serialised as "x" and Col as "y", and the CLI stores Col in Row and Row in Col for the face box
//     (faceCoord.Col = face.Row - face.Scale/2, faceCoord.Row = face.Col - face.Scale/2) and for every point
//     (Col: p.Row, Row: p.Col);
//   - `omitempty` drops zero coordinates;
//   - eyes and landmark points ACCUMULATE over the faces of an image (the slices are declared outside the loop), so
//     detection i carries the points of faces 0..i;
//   - only faces with Q > 5.0 are reported.
package pigo

import "encoding/json"

// Coord mirrors `coord`, cmd/pigo/main.go:88-92.
type Coord struct {
	Row   int `json:"x,omitempty"`
	Col   int `json:"y,omitempty"`
	Scale int `json:"size,omitempty"`
}

// DetectionJSON mirrors `detection`, cmd/pigo/main.go:95-99.
type DetectionJSON struct {
	EyePoints      []Coord `json:"eyes,omitempty"`
	LandmarkPoints []Coord `json:"landmark_points,omitempty"`
	FacePoints     Coord   `json:"face,omitempty"`
}

// WireDetections builds what drawFaces returns (cmd/pigo/main.go:369-574) from DetectBatch results of one image.
func WireDetections(faces []FaceResult) []DetectionJSON {
	const qThresh = 5.0 // main.go:360
	dets := make([]DetectionJSON, 0, len(faces))
	eyes := make([]Coord, 0, len(faces))
	lms := make([]Coord, 0, len(faces))
	for _, f := range faces {
		if !(f.Face.Q > qThresh) {
			continue
		}
		fc := Coord{Col: f.Face.Row - f.Face.Scale/2, Row: f.Face.Col - f.Face.Scale/2, Scale: f.Face.Scale} // main.go:394-398
		if f.LeftEye != nil {
			for _, e := range []*Puploc{f.LeftEye, f.RightEye} {
				if e.Row > 0 && e.Col > 0 { // main.go:423, :463
					eyes = append(eyes, Coord{Col: e.Row, Row: e.Col, Scale: int(e.Scale)})
				}
			}
			for _, p := range f.Landmarks {
				if p.Row > 0 && p.Col > 0 { // main.go:497
					lms = append(lms, Coord{Col: p.Row, Row: p.Col, Scale: int(p.Scale)})
				}
			}
		}
		dets = append(dets, DetectionJSON{FacePoints: fc, EyePoints: eyes, LandmarkPoints: lms})
	}
	return dets
}

// MarshalWire is json.Marshal of WireDetections (what the CLI writes with -json).
func MarshalWire(faces []FaceResult) ([]byte, error) { return json.Marshal(WireDetections(faces)) }
