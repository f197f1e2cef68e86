//go:build b200

// video_b200.go -- the hooks that bind gen2brain/mpeg's OWN Go parser (video.go @ 27c6f084) to libmpegb200.so.
// Drop this file into the reference's package directory next to video.go and build with `-tags b200`; the five
// edits to video.go listed at the bottom replace the per-macroblock / per-block kernel calls by record packing and
// one cgo call per picture.  Every helper INTEGRATION.md names is defined here: newRecord, packPrediction,
// copySparse, resolveRewrites, launchPicture.
//
// STATUS: NOT COMPILED (no Go toolchain in the build image).  The logic is the one the C++ host parser
// (mpeg_b200/csrc/host_parser.cpp: new_record, pack_prediction, decode_block hand-over, emit_picture) and the Python
// packer (mpeg_b200/packing.py: resolve_rewrites) implement and the tests pin against the reference's golden hashes.
package mpeg

/*
#cgo CFLAGS:  -I${SRCDIR}/include
#cgo LDFLAGS: -L${SRCDIR}/lib -lmpegb200
#include "mpegb200.h"
*/
import "C"

import (
	"math/bits"
	"unsafe"
)

// b200Video is the state a Video carries with the b200 tag (embed it in type Video: `b200 b200Video`).
type b200Video struct {
	ctx    *C.mpegb200_ctx
	stream C.int32_t

	cur, fwd, bwd C.uint8_t // physical buffers playing frameCurrent / frameForward / frameBackward

	mbs     []C.mpegb200_mb // records of the picture under construction
	levels  []int16         // 64 per coded block: the value of video.go:729-741 before the premultiply (:744)
	nBlocks int
	rec     int // index of the current macroblock's record in mbs

	lastWriter []int32 // per macroblock address: index into mbs of the record that wrote it last, -1 = none
	rewrites   bool
}

// initB200 goes where decodeSequenceHeader calls initFrame three times (video.go:324-326).
func (v *Video) initB200(ctx *C.mpegb200_ctx, stream int) bool {
	b := &v.b200
	b.ctx, b.stream = ctx, C.int32_t(stream)
	b.cur, b.fwd, b.bwd = 0, 1, 2
	b.lastWriter = make([]int32, v.mbSize)
	C.mpegb200_set_validate(ctx, 1) // bitstream-derived records: malformed launches fail instead of decoding
	return C.mpegb200_video_open(ctx, C.int(stream), C.int(v.width), C.int(v.height)) == 0
}

// beginPicture goes at the top of decodePicture's macroblock loop (video.go:411): the rotation of
// video.go:406-409 happens on indices (`temp := b.fwd; if I or P { b.fwd = b.bwd }`).
func (v *Video) beginPicture() {
	b := &v.b200
	b.mbs = b.mbs[:0]
	b.levels = b.levels[:0]
	b.nBlocks = 0
	b.rec = -1
	b.rewrites = false
	for i := range b.lastWriter {
		b.lastWriter[i] = -1
	}
}

// newRecord appends the record of the macroblock at (v.mbRow, v.mbCol): called where decodeMacroblock has the
// address of a coded macroblock (video.go:514-515) and once per skipped macroblock (video.go:503-510).
func (v *Video) newRecord() *C.mpegb200_mb {
	b := &v.b200
	b.mbs = append(b.mbs, C.mpegb200_mb{
		mb_row:      C.uint16_t(v.mbRow),
		mb_col:      C.uint16_t(v.mbCol),
		coeff_block: C.uint32_t(b.nBlocks),
	})
	b.rec = len(b.mbs) - 1
	addr := v.mbRow*v.mbWidth + v.mbCol
	if b.lastWriter[addr] >= 0 {
		b.rewrites = true // a slice revisits a macroblock: serial semantics are restored in resolveRewrites
	}
	b.lastWriter[addr] = int32(b.rec)
	return &b.mbs[b.rec]
}

// packPrediction replaces predictMacroblock (video.go:608-637): the decision only, no pixels.  In a B picture with
// both vectors set the backward copy overwrites the forward one (:626-630), so only the backward prediction is kept.
func (v *Video) packPrediction(rec *C.mpegb200_mb) {
	h, w := v.motionForward.H, v.motionForward.V
	if v.motionForward.FullPx != 0 {
		h, w = h*2, w*2
	}
	useBwd := false
	if v.pictureType == pictureTypeB && (!v.motionForward.IsSet || v.motionBackward.IsSet) {
		useBwd = true
		h, w = v.motionBackward.H, v.motionBackward.V
		if v.motionBackward.FullPx != 0 {
			h, w = h*2, w*2
		}
	}
	rec.flags |= C.MPEGB200_MB_PREDICT
	if useBwd {
		rec.flags |= C.MPEGB200_MB_REF_BWD
	}
	rec.mv_h, rec.mv_v = C.int16_t(h), C.int16_t(w)
}

func clampInt16(x int) int16 {
	if x > 32767 {
		return 32767
	}
	if x < -32768 {
		return -32768
	}
	return int16(x)
}

// copySparse copies rows / columns 0..3, all the sparse branch of idct reads (video.go:807-866, n < 10).
func copySparse(dst []int16, src *[64]int) {
	for r := 0; r < 4; r++ {
		for c := 0; c < 4; c++ {
			dst[r*8+c] = clampInt16(src[r*8+c])
		}
	}
}

// packBlock replaces the tail of decodeBlock (video.go:747-798).  v.blockData must hold the LEVEL, i.e. the value of
// video.go:729-741; the multiplication by videoPremultiplierMatrix (:744) moves to the GPU, and the intra DC is stored
// as dc*8 instead of dc<<8 (:672; 8 * premultiplier[0] == 256).  n is the reference's coefficient count.
func (v *Video) packBlock(block, n int) {
	b := &v.b200
	b.levels = append(b.levels, make([]int16, 64)...)
	dst := b.levels[b.nBlocks*64 : b.nBlocks*64+64]
	switch {
	case n == 1: // DC only: the rest of blockData is ignored by the reference and survives (video.go:774-777)
		dst[0] = clampInt16(v.blockData[0])
		v.blockData[0] = 0
	case n < 10:
		copySparse(dst, &v.blockData)
		v.blockData = [64]int{}
	default:
		for i := range dst {
			dst[i] = clampInt16(v.blockData[i])
		}
		v.blockData = [64]int{}
	}
	b.mbs[b.rec].cbp |= C.uint8_t(0x20 >> block)
	b.nBlocks++
}

// wave is one duplicate-free launch of a picture.
type wave struct {
	mbs     []C.mpegb200_mb
	levels  []int16
	nBlocks int
}

// resolveRewrites splits the picture's records into launches without double writes (INTEGRATION.md section 5): a
// later record that defines all six blocks of its macroblock (predicted, or intra with cbp 63) replaces the earlier
// one; a later record that defines only some blocks must see the earlier result, so the picture is cut there.
func (v *Video) resolveRewrites() []wave {
	b := &v.b200
	n := len(b.mbs)
	dead := make([]bool, n)
	cuts := []int{}
	if b.rewrites {
		seen := make([]int32, v.mbSize)
		reset := func() {
			for i := range seen {
				seen[i] = -1
			}
		}
		reset()
		for i := 0; i < n; i++ {
			m := &b.mbs[i]
			addr := int(m.mb_row)*v.mbWidth + int(m.mb_col)
			if seen[addr] >= 0 {
				complete := m.flags&C.MPEGB200_MB_PREDICT != 0 || (m.flags&C.MPEGB200_MB_INTRA != 0 && m.cbp == 0x3f)
				if complete {
					dead[seen[addr]] = true
				} else {
					cuts = append(cuts, i)
					reset()
				}
			}
			seen[addr] = int32(i)
		}
	}
	cuts = append(cuts, n)
	waves := make([]wave, 0, len(cuts))
	begin := 0
	for _, cut := range cuts {
		var w wave
		for i := begin; i < cut; i++ {
			if dead[i] {
				continue
			}
			m := b.mbs[i]
			nc := bits.OnesCount8(uint8(m.cbp))
			src := b.levels[int(m.coeff_block)*64 : (int(m.coeff_block)+nc)*64]
			m.coeff_block = C.uint32_t(w.nBlocks)
			m.pic = 0
			w.mbs = append(w.mbs, m)
			w.levels = append(w.levels, src...)
			w.nBlocks += nc
		}
		if len(w.mbs) > 0 {
			waves = append(waves, w)
		}
		begin = cut
	}
	return waves
}

// launchPicture replaces the end of decodePicture (video.go:429-433): one cgo call per wave, then the rotation of
// the buffer indices.  Returns false on any failure (Decode then returns nil, video.go:211,231,241).
func (v *Video) launchPicture(temp C.uint8_t) bool {
	b := &v.b200
	for _, w := range v.resolveRewrites() {
		pic := C.mpegb200_picture{stream: b.stream, _type: C.uint8_t(v.pictureType), dst_buf: b.cur, fwd_buf: b.fwd,
			bwd_buf: b.bwd, first_mb: 0, n_mb: C.uint32_t(len(w.mbs))}
		var co *C.int16_t
		if w.nBlocks > 0 {
			co = (*C.int16_t)(unsafe.Pointer(&w.levels[0]))
		}
		if C.mpegb200_video_decode_pictures(b.ctx, 1, &pic, C.size_t(len(w.mbs)), &w.mbs[0], C.size_t(w.nBlocks), co) != 0 {
			return false
		}
	}
	// the arrays are re-used by the next picture: the uploads issued from them must have completed
	if C.mpegb200_sync_uploads(b.ctx) != 0 {
		return false
	}
	if v.pictureType == pictureTypeIntra || v.pictureType == pictureTypePredictive { // video.go:430-433
		b.bwd = b.cur
		b.cur = temp
	}
	return true
}

// fetchFrame fills Plane.Data of the frame Decode is about to return (buffer index as in video.go:244-258).
func (v *Video) fetchFrame(f *Frame, buf C.uint8_t) bool {
	b := &v.b200
	return C.mpegb200_video_read_planes(b.ctx, C.int(b.stream), C.int(buf), (*C.uint8_t)(unsafe.Pointer(&f.Y.Data[0])),
		(*C.uint8_t)(unsafe.Pointer(&f.Cb.Data[0])), (*C.uint8_t)(unsafe.Pointer(&f.Cr.Data[0]))) == 0
}

/*
The edits to video.go (all under the b200 tag; the pure-Go file keeps building without it):

 1. type Video: add the field `b200 b200Video`; decodeSequenceHeader: after the three initFrame calls (video.go:324-326)
    call v.initB200(ctx, stream).
 2. decodePicture (video.go:406-411): `temp := v.b200.fwd; if I or P { v.b200.fwd = v.b200.bwd }`; v.beginPicture();
    at the end (video.go:429-433) `if !v.launchPicture(temp) { return }` instead of the struct rotation.
 3. decodeMacroblock: `rec := v.newRecord()` where the address is known (video.go:514-515) and in the skipped-macroblock
    loop (video.go:503-510, followed by v.packPrediction(rec)); intra: `rec.flags = C.MPEGB200_MB_INTRA`
    (video.go:525-531); else `v.decodeMotionVectors(); v.packPrediction(rec)` instead of v.predictMacroblock() (:544).
 4. decodeBlock: keep video.go:639-743; drop the premultiply of :744 (`v.blockData[deZigZagged] = level`), store the intra
    DC as `v.dcPredictor[planeIndex] * 8` at :672; replace :747-798 by `v.packBlock(block, n)`.
 5. Decode (video.go:244-258): the frame to return is a buffer index (noDelay: b.bwd; B picture: b.cur; else b.fwd when a
    reference exists); call v.fetchFrame(frame, index) before returning it.  Frame.RGBA (video.go:31-36) becomes
    `C.mpegb200_video_rgba(ctx, stream, index, &f.imRGBA.Pix[0])`.
*/
