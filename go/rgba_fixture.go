//go:build ignore

// rgba_fixture.go -- run ONCE on a machine that has Go:   go run go/rgba_fixture.go > tests/golden/go_rgba_fixture.bin
//
// Frame.RGBA() in the reference is Go standard-library arithmetic (video.go:31-36: draw.Draw of an image.YCbCr with
// 4:2:0 subsampling into an image.RGBA with draw.Src), which is not part of /root/reference and could not be executed
// in the build image.  This program writes what that call produces for deterministic planes of a macroblock-padded frame
// with odd display sizes; tests/test_oracle_golden.py::test_rgba_against_go_fixture compares the oracle's restatement
// (oracle/orc_pixel.c:orc_rgba) with it when the file is present.  Until someone runs it the RGBA row of SURVEY 8a
// stays "parity unpinned" (DESIGN.md section 1).
//
// Layout of the output (little endian): magic "RGBAFIX1", n_cases uint32, then per case: width, height, luma_w, luma_h
// uint32; Y (luma_w*luma_h), Cb, Cr ((luma_w/2)*(luma_h/2) each), RGBA (width*height*4).
package main

import (
	"bufio"
	"encoding/binary"
	"image"
	"image/draw"
	"os"
)

func main() {
	w := bufio.NewWriter(os.Stdout)
	defer w.Flush()
	cases := [][2]int{{160, 120}, {33, 17}, {1, 1}, {352, 288}, {47, 31}}
	w.WriteString("RGBAFIX1")
	binary.Write(w, binary.LittleEndian, uint32(len(cases)))
	seed := uint32(20260925)
	next := func() byte { // the LCG of Numerical Recipes; the high byte
		seed = seed*1664525 + 1013904223
		return byte(seed >> 24)
	}
	for _, c := range cases {
		width, height := c[0], c[1]
		lw, lh := ((width+15)>>4)<<4, ((height+15)>>4)<<4 // video.go:314-322
		cw, ch := lw/2, lh/2
		y, cb, cr := make([]byte, lw*lh), make([]byte, cw*ch), make([]byte, cw*ch)
		for i := range y {
			y[i] = next()
		}
		for i := range cb {
			cb[i] = next()
			cr[i] = next()
		}
		// extremes in the first pixels: saturation on both sides of every channel
		copy(y, []byte{0, 255, 0, 255, 16, 235})
		copy(cb, []byte{0, 255, 255})
		copy(cr, []byte{255, 0, 255})
		src := image.YCbCr{Y: y, Cb: cb, Cr: cr, SubsampleRatio: image.YCbCrSubsampleRatio420, YStride: lw, CStride: cw,
			Rect: image.Rect(0, 0, width, height)} // video.go:357-365
		dst := image.RGBA{Pix: make([]byte, width*height*4), Stride: 4 * width, Rect: image.Rect(0, 0, width, height)}
		b := src.Bounds()
		draw.Draw(&dst, b.Bounds(), &src, b.Min, draw.Src) // video.go:32-33
		for _, v := range []int{width, height, lw, lh} {
			binary.Write(w, binary.LittleEndian, uint32(v))
		}
		w.Write(y)
		w.Write(cb)
		w.Write(cr)
		w.Write(dst.Pix)
	}
}
