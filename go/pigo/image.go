// image.go -- the reference's ingest helpers (core/image.go, core/grayscale.go) so that callers such as
// cmd/pigo/main.go:288-300 (GetImage -> RgbToGrayscale -> ImageParams) build against this package unchanged.
// File I/O and container decoding stay in Go (image.Decode, like the reference); the per-pixel arithmetic runs on the GPU:
//   - *image.YCbCr (what a JPEG decodes to): pigo_ycbcr_to_nrgba, bit-exact with color.YCbCrToRGB (core/image.go:60-76);
//   - *image.NRGBA -> []uint8 luma: pigo_rgba_to_gray (core/grayscale.go:8-23).
package pigo

/*
#include "pigo_b200.h"
*/
import "C"

import (
	"image"
	"image/color"
	_ "image/jpeg"
	_ "image/png"
	"io"
	"os"
	"unsafe"
)

// GetImage mirrors core/image.go:13-21.
func GetImage(input string) (*image.NRGBA, error) {
	file, err := os.Open(input)
	if err != nil {
		return nil, err
	}
	defer file.Close()
	return DecodeImage(file)
}

// DecodeImage mirrors core/image.go:24-32.
func DecodeImage(f io.Reader) (*image.NRGBA, error) {
	src, _, err := image.Decode(f)
	if err != nil {
		return nil, err
	}
	return ImgToNRGBA(src), nil
}

// ImgToNRGBA mirrors core/image.go:35-90.
func ImgToNRGBA(img image.Image) *image.NRGBA {
	srcBounds := img.Bounds()
	if srcBounds.Min.X == 0 && srcBounds.Min.Y == 0 {
		if src0, ok := img.(*image.NRGBA); ok {
			return src0
		}
	}
	dstBounds := srcBounds.Sub(srcBounds.Min)
	dstW, dstH := dstBounds.Dx(), dstBounds.Dy()
	dst := image.NewNRGBA(dstBounds)
	if dstW == 0 || dstH == 0 {
		return dst
	}
	// *image.YCbCr (what a JPEG decodes to): the per-pixel conversion runs on the GPU.  Negative rectangle origins keep the
	// reference's own loop below (Go's x/2 truncates toward zero, the library's chroma indexing assumes origins >= 0).
	if src, ok := img.(*image.YCbCr); ok && srcBounds.Min.X >= 0 && srcBounds.Min.Y >= 0 {
		yo, co := src.YOffset(srcBounds.Min.X, srcBounds.Min.Y), src.COffset(srcBounds.Min.X, srcBounds.Min.Y)
		if _, err := call(func() C.int {
			return C.pigo_ycbcr_to_nrgba((*C.uint8_t)(unsafe.Pointer(&src.Y[yo])), (*C.uint8_t)(unsafe.Pointer(&src.Cb[co])),
				(*C.uint8_t)(unsafe.Pointer(&src.Cr[co])), C.int(src.YStride), C.int(src.CStride), C.int(src.SubsampleRatio),
				C.int(srcBounds.Min.X), C.int(srcBounds.Min.Y), C.int(dstW), C.int(dstH), (*C.uint8_t)(unsafe.Pointer(&dst.Pix[0])), nil,
				C.PIGO_MEM_HOST, nil)
		}); err != nil {
			panic(err)
		}
		return dst
	}
	switch src := img.(type) {
	case *image.NRGBA:
		for y := 0; y < dstH; y++ { // row copy, core/image.go:52-59
			si := src.PixOffset(srcBounds.Min.X, srcBounds.Min.Y+y)
			copy(dst.Pix[y*dst.Stride:y*dst.Stride+dstW*4], src.Pix[si:si+dstW*4])
		}
	default:
		for y := 0; y < dstH; y++ { // core/image.go:77-88 (and :60-76 for YCbCr images with a negative origin: color.NRGBAModel goes through YCbCrToRGB too)
			di := dst.PixOffset(0, y)
			for x := 0; x < dstW; x++ {
				c := color.NRGBAModel.Convert(img.At(srcBounds.Min.X+x, srcBounds.Min.Y+y)).(color.NRGBA)
				dst.Pix[di+0], dst.Pix[di+1], dst.Pix[di+2], dst.Pix[di+3] = c.R, c.G, c.B, c.A
				di += 4
			}
		}
	}
	return dst
}

// RgbToGrayscale mirrors core/grayscale.go:8-23.
func RgbToGrayscale(src image.Image) []uint8 {
	nrgba := ImgToNRGBA(src) // *image.NRGBA at origin (0,0) is returned as is, like the reference's callers pass it
	w, h := nrgba.Bounds().Dx(), nrgba.Bounds().Dy()
	gray := make([]uint8, w*h)
	if w*h == 0 {
		return gray
	}
	pix := nrgba.Pix
	if nrgba.Stride != 4*w { // compact the rows: the library takes [npixels][4]
		pix = make([]uint8, 4*w*h)
		for y := 0; y < h; y++ {
			copy(pix[y*4*w:(y+1)*4*w], nrgba.Pix[y*nrgba.Stride:y*nrgba.Stride+4*w])
		}
	}
	if _, err := call(func() C.int {
		return C.pigo_rgba_to_gray((*C.uint8_t)(unsafe.Pointer(&pix[0])), C.size_t(w*h), (*C.uint8_t)(unsafe.Pointer(&gray[0])), C.PIGO_MEM_HOST, nil)
	}); err != nil {
		panic(err)
	}
	return gray
}
