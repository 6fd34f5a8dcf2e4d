// pipeline.go -- additive: the face -> ClusterDetections -> pupils -> landmarks sequence of cmd/pigo/main.go:369-565 /
// core/flploc_test.go:75-154 for a whole frame batch in ONE library call (pigo_detect_batch: sequenced on the device).
package pigo

/*
#include <stdlib.h>
#include "pigo_b200.h"
*/
import "C"

import (
	"sync/atomic"
	"unsafe"
)

// LandmarkCall is one GetLandmarkPoint call of the per-face sequence.
type LandmarkCall struct {
	Cascade *PuplocCascade
	FlipV   bool
}

// ReferenceLandmarkCalls returns the reference's 15-call sequence (core/flploc_test.go:122-146, cmd/pigo/main.go:493-563).
func ReferenceLandmarkCalls(flpcs map[string][]*FlpCascade) []LandmarkCall {
	var calls []LandmarkCall
	for _, eye := range []string{"lp46", "lp44", "lp42", "lp38", "lp312"} {
		for _, flpc := range flpcs[eye] {
			calls = append(calls, LandmarkCall{flpc.PuplocCascade, false}, LandmarkCall{flpc.PuplocCascade, true})
		}
	}
	for _, mouth := range []string{"lp93", "lp84", "lp82", "lp81"} {
		for _, flpc := range flpcs[mouth] {
			calls = append(calls, LandmarkCall{flpc.PuplocCascade, false})
		}
	}
	return append(calls, LandmarkCall{flpcs["lp84"][0].PuplocCascade, true})
}

// FaceResult is one clustered face with its refinements (nil / empty when Scale <= MinFaceScale).
type FaceResult struct {
	Face      Detection
	LeftEye   *Puploc
	RightEye  *Puploc
	Landmarks []Puploc
}

// PipelineParams are the knobs the reference's callers hard-code (main.go:404,420; flploc_test.go:102,107).
type PipelineParams struct {
	IoU          float64
	Angle        float64
	MinFaceScale int
	EyePerturbs  int
	FlpPerturbs  int
	Sharded      bool // spread the frames over the GPUs of InitDevices
}

// DetectBatch runs the whole sequence for every frame; result[f] lists frame f's clusters in the reference's order.
func (pg *Pigo) DetectBatch(plc *PuplocCascade, calls []LandmarkCall, frames [][]uint8, cp CascadeParams, pp PipelineParams) [][]FaceResult {
	nf := len(frames)
	if nf == 0 {
		return nil
	}
	stride := cp.Rows * cp.Dim
	var pinned unsafe.Pointer
	if _, err := call(func() C.int { return C.pigo_alloc_pinned(&pinned, C.size_t(stride*nf)) }); err != nil {
		panic(err)
	}
	defer C.pigo_free_pinned(pinned)
	host := unsafe.Slice((*uint8)(pinned), stride*nf)
	for i, f := range frames {
		copy(host[i*stride:(i+1)*stride], f)
	}
	ncalls := len(calls)
	// C arrays of the call list (cgo: no Go pointers to Go pointers)
	hs := (*[1 << 10]*C.pigo_puploc)(C.malloc(C.size_t(ncalls+1) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(hs))
	fl := make([]C.uint8_t, ncalls+1)
	for i, c := range calls {
		hs[i] = c.Cascade.h
		if c.FlipV {
			fl[i] = 1
		}
	}
	prm := C.pigo_pipeline_params{min_size: C.int32_t(cp.MinSize), max_size: C.int32_t(cp.MaxSize), shift_factor: C.double(cp.ShiftFactor),
		scale_factor: C.double(cp.ScaleFactor), angle: C.double(pp.Angle), iou_threshold: C.double(pp.IoU),
		min_face_scale: C.int32_t(pp.MinFaceScale), eye_perturbs: C.int32_t(pp.EyePerturbs), flp_perturbs: C.int32_t(pp.FlpPerturbs)}
	key := atomic.AddUint64(&plc.seed, 1)
	faceCap := 32
	for {
		faces := make([]C.pigo_det, nf*faceCap)
		nfaces := make([]C.int, nf)
		points := make([]C.pigo_point, nf*faceCap*(2+ncalls))
		rc, err := call(func() C.int {
			if pp.Sharded {
				return C.pigo_detect_batch_sharded(pg.h, plc.h, (**C.pigo_puploc)(unsafe.Pointer(hs)), &fl[0], C.int(ncalls), (*C.uint8_t)(pinned), C.int(nf),
					C.size_t(stride), C.int(cp.Rows), C.int(cp.Cols), C.int(cp.Dim), &prm, nil, C.uint64_t(key), &faces[0], C.int(faceCap), &nfaces[0], &points[0])
			}
			return C.pigo_detect_batch(pg.h, plc.h, (**C.pigo_puploc)(unsafe.Pointer(hs)), &fl[0], C.int(ncalls), (*C.uint8_t)(pinned), C.int(nf),
				C.size_t(stride), C.int(cp.Rows), C.int(cp.Cols), C.int(cp.Dim), &prm, nil, C.uint64_t(key), &faces[0], C.int(faceCap), &nfaces[0], &points[0],
				C.PIGO_MEM_HOST, nil)
		})
		if rc == C.PIGO_E_CAP {
			grown := false
			for _, n := range nfaces {
				if int(n) > faceCap {
					faceCap, grown = int(n), true
				}
			}
			if !grown {
				prm.det_cap = 4 * (prm.det_cap + 512) // raw detections per frame exceeded det_cap
			}
			continue
		}
		if err != nil {
			panic(err)
		}
		out := make([][]FaceResult, nf)
		for f := 0; f < nf; f++ {
			for k := 0; k < int(nfaces[f]); k++ {
				d := faces[f*faceCap+k]
				fr := FaceResult{Face: Detection{int(d.row), int(d.col), int(d.scale), float32(d.q)}}
				if int(d.scale) > pp.MinFaceScale {
					p := points[(f*faceCap+k)*(2+ncalls):]
					fr.LeftEye = &Puploc{Row: int(p[0].row), Col: int(p[0].col), Scale: float32(p[0].scale)}
					fr.RightEye = &Puploc{Row: int(p[1].row), Col: int(p[1].col), Scale: float32(p[1].scale)}
					for c := 0; c < ncalls; c++ {
						fr.Landmarks = append(fr.Landmarks, Puploc{Row: int(p[2+c].row), Col: int(p[2+c].col), Scale: float32(p[2+c].scale)})
					}
				}
				out[f] = append(out[f], fr)
			}
		}
		return out
	}
}
