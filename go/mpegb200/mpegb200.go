//go:build b200

// Package mpegb200 is the cgo shim over libmpegb200.so: the public surface of gen2brain/mpeg
// (mpeg.New, Video.Decode() *Frame, Audio.Decode() *Samples, Frame.RGBA(); mpeg.go:85, video.go:209,31,
// audio.go:163 @ 27c6f084) with every pixel and sample produced by the sm_100a kernels behind the C-ABI of
// include/mpegb200.h.
//
// STATUS: written against cgo and the reference's identifiers, NOT COMPILED -- no Go toolchain exists in the
// build image (SURVEY section 0).  The same call sequence is exercised end to end by the Python mirror
// (mpeg_b200/mpeg.py) and the plain-C driver tests/c_abi/driver.c, both of which run on the GPU box.
//
// This file drives the library's own host half (include/mpegb200_host.h: bit reader, VLC parse,
// dequantisation, PS demux -- the C++ restatement of the reference's serial half) and the kernels.  A
// maintainer who prefers to keep the reference's Go parser links go/inpackage/video_b200.go into package
// mpeg instead; both end in the same two calls per picture.
//
// Conventions kept from the reference: constructors return errors (ErrInvalidMPEG, mpeg.go:55); Decode
// calls return nil at the end of the stream or on any failure (video.go:211,231,241; audio.go:165-170);
// a *Frame / *Samples is valid until the next Decode of its decoder (mpeg.go:413-415).
package mpegb200

/*
#cgo CFLAGS:  -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../mpeg_b200 -lmpegb200
#include <stdlib.h>
#include "mpegb200.h"
#include "mpegb200_host.h"
*/
import "C"

import (
	"bytes"
	"errors"
	"fmt"
	"image"
	"image/color"
	"io"
	"runtime"
	"sync"
	"unsafe"
)

// ErrInvalidMPEG mirrors mpeg.ErrInvalidMPEG (mpeg.go:55).
var ErrInvalidMPEG = errors.New("invalid MPEG-PS")

// ErrNoDevice is returned when no sm_100 (B200) device is present: the library has no CPU fallback.
var ErrNoDevice = errors.New("mpegb200: an sm_100 CUDA device is required")

// SamplesPerFrame mirrors mpeg.SamplesPerFrame (audio.go:9).
const SamplesPerFrame = C.MPEGB200_SAMPLES_PER_FRAME

// AudioFormat mirrors mpeg.AudioFormat (audio.go:12-23); the values are the C-ABI's.
type AudioFormat int

const (
	AudioF32N   AudioFormat = C.MPEGB200_AUDIO_F32N
	AudioF32NLR AudioFormat = C.MPEGB200_AUDIO_F32NLR
	AudioF32    AudioFormat = C.MPEGB200_AUDIO_F32
	AudioS16    AudioFormat = C.MPEGB200_AUDIO_S16
)

// ---------------------------------------------------------------------------------------------------
// GPU context: one per device, shared by the decoders that live on it (INTEGRATION.md section 2)
// ---------------------------------------------------------------------------------------------------

// GPU owns a mpegb200_ctx.  Decoders on one GPU share it and each owns one stream id.
type GPU struct {
	ctx     *C.mpegb200_ctx
	mu      sync.Mutex
	free    []int // stream ids not in use
	streams int
}

// OpenGPU creates a context for up to maxStreams video and maxStreams audio streams on `device`.
func OpenGPU(device, maxStreams int) (*GPU, error) {
	var rc C.int
	ctx := C.mpegb200_create(C.int(device), C.int(maxStreams), &rc)
	if ctx == nil {
		if rc == C.MPEGB200_ECUDA {
			return nil, ErrNoDevice
		}
		return nil, fmt.Errorf("mpegb200_create: error %d", int(rc))
	}
	g := &GPU{ctx: ctx, streams: maxStreams}
	for s := maxStreams - 1; s >= 0; s-- {
		g.free = append(g.free, s)
	}
	runtime.SetFinalizer(g, (*GPU).Close)
	return g, nil
}

// Close releases the context and all device memory.
func (g *GPU) Close() {
	if g.ctx != nil {
		C.mpegb200_destroy(g.ctx)
		g.ctx = nil
	}
}

func (g *GPU) lastError() string { return C.GoString(C.mpegb200_last_error(g.ctx)) }

func (g *GPU) takeStream() (int, error) {
	g.mu.Lock()
	defer g.mu.Unlock()
	if len(g.free) == 0 {
		return -1, fmt.Errorf("mpegb200: all %d stream ids in use", g.streams)
	}
	s := g.free[len(g.free)-1]
	g.free = g.free[:len(g.free)-1]
	return s, nil
}

func (g *GPU) giveStream(s int) {
	g.mu.Lock()
	g.free = append(g.free, s)
	g.mu.Unlock()
}

var (
	defaultGPU     *GPU
	defaultGPUErr  error
	defaultGPUOnce sync.Once
)

// DefaultGPU is device 0 with room for 64 streams, created on first use (what mpeg.New uses).
func DefaultGPU() (*GPU, error) {
	defaultGPUOnce.Do(func() { defaultGPU, defaultGPUErr = OpenGPU(0, 64) })
	return defaultGPU, defaultGPUErr
}

// ---------------------------------------------------------------------------------------------------
// Frame / Plane (video.go:10-54)
// ---------------------------------------------------------------------------------------------------

// Plane mirrors mpeg.Plane: macroblock-padded plane data (video.go:45-54).
type Plane struct {
	Width  int
	Height int
	Data   []byte
}

// Frame mirrors mpeg.Frame (video.go:10-24).  The planes are fetched from the device when the frame is
// returned; RGBA is converted on the device on demand.
type Frame struct {
	Time float64

	Width  int
	Height int

	Y  Plane
	Cb Plane
	Cr Plane

	imYCbCr image.YCbCr
	imRGBA  image.RGBA

	v   *Video
	buf int // physical buffer 0..2 on the device
}

// YCbCr returns the frame as image.YCbCr (video.go:26-28).
func (f *Frame) YCbCr() *image.YCbCr { return &f.imYCbCr }

// RGBA returns the frame as image.RGBA (video.go:31-36): converted by rgba_kernel, copied into Pix.
func (f *Frame) RGBA() *image.RGBA {
	rc := C.mpegb200_video_rgba(f.v.gpu.ctx, C.int(f.v.stream), C.int(f.buf), (*C.uint8_t)(unsafe.Pointer(&f.imRGBA.Pix[0])))
	if rc != 0 {
		return nil
	}
	return &f.imRGBA
}

// Pixels returns the frame as a slice of color.RGBA (video.go:39-43).
func (f *Frame) Pixels() []color.RGBA {
	img := f.RGBA()
	if img == nil {
		return nil
	}
	return unsafe.Slice((*color.RGBA)(unsafe.Pointer(&img.Pix[0])), len(img.Pix)/4)
}

// ---------------------------------------------------------------------------------------------------
// Video (video.go:57-268)
// ---------------------------------------------------------------------------------------------------

// Video decodes an MPEG-1 video elementary stream.
type Video struct {
	gpu    *GPU
	stream int
	parser *C.mpegb200_video_parser
	data   unsafe.Pointer // C copy of the elementary stream (the parser keeps its own copy; freed at once)
	opened bool
	deviceVLC bool // SetDeviceVLC: slices parsed on the GPU from the stream's resident copy
	frames [3]Frame // one per physical buffer, like frameCurrent / frameForward / frameBackward (video.go:97-99)
	time   float64
}

// NewVideo mirrors mpeg.NewVideo (video.go:110-122) for a fully resident elementary stream.
func NewVideo(gpu *GPU, es []byte) (*Video, error) {
	stream, err := gpu.takeStream()
	if err != nil {
		return nil, err
	}
	var p *C.uint8_t
	if len(es) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&es[0]))
	}
	parser := C.mpegb200_video_parser_new(p, C.size_t(len(es))) // copies the bytes
	if parser == nil {
		gpu.giveStream(stream)
		return nil, errors.New("mpegb200: out of memory")
	}
	// the parser writes the variable-width coefficient form straight from its zig-zag walk (no int16[64] blocks)
	C.mpegb200_video_parser_set_vlen(parser, 1)
	v := &Video{gpu: gpu, stream: stream, parser: parser}
	runtime.SetFinalizer(v, (*Video).Close)
	return v, nil
}

// Close releases the parser and the stream's frame buffers.
func (v *Video) Close() {
	if v.parser != nil {
		C.mpegb200_video_parser_free(v.parser)
		v.parser = nil
		if v.opened {
			C.mpegb200_video_close(v.gpu.ctx, C.int(v.stream))
		}
		v.gpu.giveStream(v.stream)
	}
}

// HasHeader mirrors Video.HasHeader (video.go:130-147).
func (v *Video) HasHeader() bool { return C.mpegb200_video_parser_has_header(v.parser) != 0 }

// Framerate mirrors Video.Framerate (video.go:150-156).
func (v *Video) Framerate() float64 { return float64(C.mpegb200_video_parser_framerate(v.parser)) }

// Width mirrors Video.Width (video.go:159-165).
func (v *Video) Width() int { return int(C.mpegb200_video_parser_width(v.parser)) }

// Height mirrors Video.Height (video.go:168-174).
func (v *Video) Height() int { return int(C.mpegb200_video_parser_height(v.parser)) }

// SetNoDelay mirrors Video.SetNoDelay (video.go:176-180).
func (v *Video) SetNoDelay(noDelay bool) {
	on := C.int(0)
	if noDelay {
		on = 1
	}
	C.mpegb200_video_parser_set_no_delay(v.parser, on)
}

// Time mirrors Video.Time (video.go:183-186).
func (v *Video) Time() float64 { return v.time }

// Rewind mirrors Video.Rewind (video.go:195-201).
func (v *Video) Rewind() {
	C.mpegb200_video_parser_rewind(v.parser)
	v.time = 0
}

// HasEnded mirrors Video.HasEnded (video.go:204-206).
func (v *Video) HasEnded() bool { return C.mpegb200_video_parser_has_ended(v.parser) != 0 }

// open allocates the device frame buffers and the host planes once the sequence header is known
// (initFrame, video.go:333-372).
func (v *Video) open() bool {
	if v.opened {
		return true
	}
	w, h := v.Width(), v.Height()
	if w <= 0 || h <= 0 {
		return false
	}
	if C.mpegb200_video_open(v.gpu.ctx, C.int(v.stream), C.int(w), C.int(h)) != 0 {
		return false
	}
	// records parsed from a bitstream are untrusted: a malformed launch fails loudly (the reference panics)
	C.mpegb200_set_validate(v.gpu.ctx, 1)
	var lw, lh, cw, ch C.int
	var fb C.size_t
	C.mpegb200_video_geometry(v.gpu.ctx, C.int(v.stream), &lw, &lh, &cw, &ch, &fb)
	for i := range v.frames {
		f := &v.frames[i]
		lumaSize, chromaSize := int(lw)*int(lh), int(cw)*int(ch)
		base := make([]byte, lumaSize+2*chromaSize)
		f.Width, f.Height = w, h
		f.Y = Plane{int(lw), int(lh), base[0:lumaSize]}
		f.Cb = Plane{int(cw), int(ch), base[lumaSize : lumaSize+chromaSize]}
		f.Cr = Plane{int(cw), int(ch), base[lumaSize+chromaSize:]}
		f.imYCbCr = image.YCbCr{Y: f.Y.Data, Cb: f.Cb.Data, Cr: f.Cr.Data, SubsampleRatio: image.YCbCrSubsampleRatio420,
			YStride: int(lw), CStride: int(cw), Rect: image.Rect(0, 0, w, h)}
		f.imRGBA = image.RGBA{Pix: make([]byte, w*h*4), Stride: 4 * w, Rect: image.Rect(0, 0, w, h)}
		f.v, f.buf = v, i
	}
	v.opened = true
	return true
}

// SetDeviceVLC switches the stream to the slice-parallel VLC stage (SURVEY 8f1): the elementary stream is uploaded to device
// memory once, its start codes are indexed there, and from then on Decode only reads headers on the host; the slices are
// parsed by one GPU thread each (mpegb200_video_decode_bitstream) and pictures whose slices are not independent (damaged
// streams) finish on the host parser with the serial semantics of the reference.  Call before the first Decode.
func (v *Video) SetDeviceVLC(es []byte) error {
	if len(es) == 0 {
		return errors.New("mpegb200: empty stream")
	}
	ctx := v.gpu.ctx
	if C.mpegb200_video_stream_upload(ctx, C.int(v.stream), (*C.uint8_t)(unsafe.Pointer(&es[0])), C.size_t(len(es))) != 0 {
		return errors.New("mpegb200: " + v.gpu.lastError())
	}
	at := make([]C.uint64_t, len(es)/64+4096)
	var n C.size_t
	rc := C.mpegb200_video_stream_index(ctx, C.int(v.stream), &at[0], C.size_t(len(at)), &n)
	if rc != 0 && int(n) > len(at) { // a stream of (almost) nothing but start codes
		at = make([]C.uint64_t, int(n))
		rc = C.mpegb200_video_stream_index(ctx, C.int(v.stream), &at[0], C.size_t(len(at)), &n)
	}
	if rc != 0 || C.mpegb200_video_parser_set_start_codes(v.parser, &at[0], n) != 0 {
		return errors.New("mpegb200: start-code index: " + v.gpu.lastError())
	}
	v.deviceVLC = true
	return nil
}

// runLaunches executes the launches of a host-parsed step (variable-width coefficient form).
func (v *Video) runLaunches(step *C.mpegb200_video_step) bool {
	launches := unsafe.Slice(step.launches, int(step.n_launches))
	vlen := unsafe.Slice(step.vlen_launches, int(step.n_launches))
	for i := range launches {
		ln := &launches[i]
		if ln.n_mb == 0 {
			continue
		}
		pic := ln.picture
		pic.stream = C.int32_t(v.stream)
		mbs := (*C.mpegb200_mb)(unsafe.Add(unsafe.Pointer(step.mbs), uintptr(ln.first_mb)*C.sizeof_mpegb200_mb))
		lv := &vlen[i]
		if C.mpegb200_video_decode_pictures_vlen(v.gpu.ctx, 1, &pic, C.size_t(ln.n_mb), mbs, C.size_t(ln.n_blocks),
			(*C.uint32_t)(unsafe.Add(unsafe.Pointer(step.vlen_headers), uintptr(ln.first_block)*4)),
			(*C.uint64_t)(unsafe.Add(unsafe.Pointer(step.vlen_chunk_offsets), uintptr(lv.first_chunk)*8)),
			(*C.uint8_t)(unsafe.Add(unsafe.Pointer(step.vlen_payload), uintptr(lv.payload_offset))), C.size_t(lv.payload_bytes)) != 0 {
			return false
		}
	}
	return true
}

// decodeOnDevice is Decode with the slices parsed on the GPU: one wave per picture of the step (a step decodes more than one
// picture only at the start of a stream), the tables built like mpegb200_video_batch_next_scan builds them for many streams.
func (v *Video) decodeOnDevice() (frameBuf int, time float64, ok bool) {
	ctx := v.gpu.ctx
	var ss C.mpegb200_video_scan_step
	if C.mpegb200_video_parser_next_scan(v.parser, &ss) != 0 || ss.has_frame == 0 {
		return 0, 0, false
	}
	frameBuf, time = int(ss.frame_buf), float64(ss.time)
	if ss.host_step != nil { // stale coefficients of a dropped block pending: the host parsed this step itself
		return frameBuf, time, v.runLaunches(ss.host_step)
	}
	pictures := unsafe.Slice(ss.pictures, int(ss.n_pictures))
	mbW, mbSize := int(ss.mb_w), int(ss.mb_w)*int(ss.mb_h)
	for k := range pictures {
		p := &pictures[k]
		scan := unsafe.Slice((*C.mpegb200_scan_slice)(unsafe.Add(unsafe.Pointer(ss.slices), uintptr(p.first_slice)*C.sizeof_mpegb200_scan_slice)), int(p.n_slices))
		slices := make([]C.mpegb200_vlc_slice, len(scan))
		slot := 0
		for i := range scan {
			from := (int(scan[i].vpos) - 1) * mbW
			to := mbSize
			if i+1 < len(scan) {
				to = min((int(scan[i+1].vpos)-1)*mbW, mbSize)
			}
			room := (max(0, to-from) + 15) &^ 15 // record slots: the addresses from this slice's row to the next slice's
			slices[i] = C.mpegb200_vlc_slice{data_offset: scan[i].offset, // resident stream: offsets count from its first byte
				next_code:   C.uint32_t(scan[i].next_code - scan[i].offset),
				stream_left: C.uint32_t(min(uint64(ss.stream_len)-uint64(scan[i].offset), 1<<32-1)),
				pic:         0, vpos: scan[i].vpos, mb_slot: C.uint32_t(slot), mb_cap: C.uint32_t(room)}
			slot += room
		}
		pic := C.mpegb200_vlc_picture{stream: C.int32_t(v.stream), _type: p._type, dst_buf: p.dst_buf, fwd_buf: p.fwd_buf, bwd_buf: p.bwd_buf,
			fwd_full_px: p.fwd_full_px, fwd_r_size: p.fwd_r_size, bwd_full_px: p.bwd_full_px, bwd_r_size: p.bwd_r_size,
			first_slice: 0, n_slices: C.uint32_t(len(slices)), mb_slot: 0, n_mb_slots: C.uint32_t(slot), quant: 0}
		var first *C.mpegb200_vlc_slice
		if len(slices) > 0 {
			first = &slices[0]
		}
		if C.mpegb200_video_decode_bitstream(ctx, 1, &pic, C.size_t(len(slices)), first, nil, 0, ss.quant, 1, C.size_t(slot)) != 0 {
			return 0, 0, false
		}
		var flag C.int
		if C.mpegb200_video_bitstream_flags(ctx, &flag, 1) < 0 {
			return 0, 0, false
		}
		if flag != 0 { // not a set of independent slices: the host parser finishes the step from this picture on
			var tail C.mpegb200_video_step
			if C.mpegb200_video_parser_redo(v.parser, C.int(k), &tail) != 0 || !v.runLaunches(&tail) {
				return 0, 0, false
			}
			return int(tail.frame_buf), float64(tail.time), tail.has_frame != 0
		}
	}
	return frameBuf, time, true
}

// Decode mirrors Video.Decode (video.go:209-268): parse up to the picture that makes a frame due, run its
// launches on the GPU, return the frame (nil at the end of the stream or on any failure).
func (v *Video) Decode() *Frame {
	if !v.HasHeader() || !v.open() {
		return nil
	}
	if v.deviceVLC {
		buf, t, ok := v.decodeOnDevice()
		if !ok {
			return nil
		}
		f := &v.frames[buf]
		f.Time, v.time = t, t
		if C.mpegb200_video_read_planes(v.gpu.ctx, C.int(v.stream), C.int(buf), (*C.uint8_t)(unsafe.Pointer(&f.Y.Data[0])),
			(*C.uint8_t)(unsafe.Pointer(&f.Cb.Data[0])), (*C.uint8_t)(unsafe.Pointer(&f.Cr.Data[0]))) != 0 {
			return nil
		}
		return f
	}
	var step C.mpegb200_video_step
	if C.mpegb200_video_parser_next(v.parser, &step) != 0 || step.has_frame == 0 {
		return nil
	}
	launches := unsafe.Slice(step.launches, int(step.n_launches))
	vlen := unsafe.Slice(step.vlen_launches, int(step.n_launches))
	for i := range launches {
		ln := &launches[i]
		if ln.n_mb == 0 {
			continue
		}
		pic := ln.picture
		pic.stream = C.int32_t(v.stream)
		mbs := (*C.mpegb200_mb)(unsafe.Add(unsafe.Pointer(step.mbs), uintptr(ln.first_mb)*C.sizeof_mpegb200_mb))
		lv := &vlen[i]
		rc := C.mpegb200_video_decode_pictures_vlen(v.gpu.ctx, 1, &pic, C.size_t(ln.n_mb), mbs, C.size_t(ln.n_blocks),
			(*C.uint32_t)(unsafe.Add(unsafe.Pointer(step.vlen_headers), uintptr(ln.first_block)*4)),
			(*C.uint64_t)(unsafe.Add(unsafe.Pointer(step.vlen_chunk_offsets), uintptr(lv.first_chunk)*8)),
			(*C.uint8_t)(unsafe.Add(unsafe.Pointer(step.vlen_payload), uintptr(lv.payload_offset))), C.size_t(lv.payload_bytes))
		if rc != 0 {
			return nil
		}
	}
	f := &v.frames[int(step.frame_buf)]
	f.Time = float64(step.time)
	v.time = f.Time
	// Plane.Data of the returned frame (synchronises the context's stream: the parser may re-use its arrays afterwards)
	if C.mpegb200_video_read_planes(v.gpu.ctx, C.int(v.stream), C.int(step.frame_buf), (*C.uint8_t)(unsafe.Pointer(&f.Y.Data[0])),
		(*C.uint8_t)(unsafe.Pointer(&f.Cb.Data[0])), (*C.uint8_t)(unsafe.Pointer(&f.Cr.Data[0]))) != 0 {
		return nil
	}
	return f
}

// ---------------------------------------------------------------------------------------------------
// Samples / Audio (audio.go:27-182)
// ---------------------------------------------------------------------------------------------------

// Samples mirrors mpeg.Samples (audio.go:27-36).
type Samples struct {
	Time        float64
	S16         []int16
	F32         []float32
	Left        []float32
	Right       []float32
	Interleaved []float32

	format AudioFormat
}

// Bytes mirrors Samples.Bytes (audio.go:39-50).
func (s *Samples) Bytes() []byte {
	switch s.format {
	case AudioF32N:
		return unsafe.Slice((*byte)(unsafe.Pointer(&s.Interleaved[0])), len(s.Interleaved)*4)
	case AudioF32:
		return unsafe.Slice((*byte)(unsafe.Pointer(&s.F32[0])), len(s.F32)*4)
	case AudioS16:
		return unsafe.Slice((*byte)(unsafe.Pointer(&s.S16[0])), len(s.S16)*2)
	default:
		return nil
	}
}

// Audio decodes an MP2 elementary stream.
type Audio struct {
	gpu     *GPU
	stream  int
	parser  *C.mpegb200_audio_parser
	format  AudioFormat
	fused   bool
	samples Samples
	pinned  []int32 // requantised samples [2][36][32] of one frame
	time    float64
}

// NewAudio mirrors mpeg.NewAudio (audio.go:83-105) for a fully resident elementary stream.
func NewAudio(gpu *GPU, es []byte) (*Audio, error) {
	stream, err := gpu.takeStream()
	if err != nil {
		return nil, err
	}
	var p *C.uint8_t
	if len(es) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&es[0]))
	}
	parser := C.mpegb200_audio_parser_new(p, C.size_t(len(es)))
	if parser == nil || C.mpegb200_audio_open(gpu.ctx, C.int(stream)) != 0 {
		if parser != nil {
			C.mpegb200_audio_parser_free(parser)
		}
		gpu.giveStream(stream)
		return nil, errors.New("mpegb200: audio stream could not be opened")
	}
	a := &Audio{gpu: gpu, stream: stream, parser: parser, pinned: make([]int32, 2*36*32)}
	a.samples.S16 = make([]int16, SamplesPerFrame*2)
	a.samples.F32 = make([]float32, SamplesPerFrame*2)
	a.samples.Interleaved = make([]float32, SamplesPerFrame*2)
	planar := make([]float32, SamplesPerFrame*2)
	a.samples.Left, a.samples.Right = planar[:SamplesPerFrame], planar[SamplesPerFrame:]
	runtime.SetFinalizer(a, (*Audio).Close)
	return a, nil
}

// Close releases the parser and the stream's synthesis state.
func (a *Audio) Close() {
	if a.parser != nil {
		C.mpegb200_audio_parser_free(a.parser)
		a.parser = nil
		C.mpegb200_audio_close(a.gpu.ctx, C.int(a.stream))
		a.gpu.giveStream(a.stream)
	}
}

// HasHeader mirrors Audio.HasHeader (audio.go:112-120).
func (a *Audio) HasHeader() bool { return C.mpegb200_audio_parser_has_header(a.parser) != 0 }

// Samplerate mirrors Audio.Samplerate (audio.go:123-129).
func (a *Audio) Samplerate() int { return int(C.mpegb200_audio_parser_samplerate(a.parser)) }

// Channels mirrors Audio.Channels (audio.go:132-134).
func (a *Audio) Channels() int { return int(C.mpegb200_audio_parser_channels(a.parser)) }

// Time mirrors Audio.Time (audio.go:137-140).
func (a *Audio) Time() float64 { return a.time }

// SetFormat selects the output format (MPEG.SetAudioFormat, mpeg.go:234-238).
func (a *Audio) SetFormat(format AudioFormat) { a.format = format }

// SetFusedWindow selects the synthesis window arithmetic: false (default) rounds the product and the sum of every
// tap separately like the reference's Go / SSE back-end (golden 0xf1b76cdf8e6cdea5, mpeg_test.go:194); true uses
// one fused multiply-add per tap like its AVX2 / NEON back-ends (golden 0x50f3ab75f5fb0fb5, mpeg_test.go:195).
func (a *Audio) SetFusedWindow(on bool) { a.fused = on }

// Rewind mirrors Audio.Rewind (audio.go:149-154): the V history on the device survives, like in the reference.
func (a *Audio) Rewind() {
	C.mpegb200_audio_parser_rewind(a.parser)
	a.time = 0
}

// Decode mirrors Audio.Decode (audio.go:163-182): one frame of 1152 samples per channel, nil at the end.
func (a *Audio) Decode() *Samples {
	var t C.double
	if C.mpegb200_audio_parser_next(a.parser, (*C.int32_t)(unsafe.Pointer(&a.pinned[0])), &t) == 0 {
		return nil
	}
	var out unsafe.Pointer
	switch a.format {
	case AudioF32N:
		out = unsafe.Pointer(&a.samples.Interleaved[0])
	case AudioF32NLR:
		out = unsafe.Pointer(&a.samples.Left[0]) // Left and Right share one allocation: 1152 left, then 1152 right
	case AudioF32:
		out = unsafe.Pointer(&a.samples.F32[0])
	default:
		out = unsafe.Pointer(&a.samples.S16[0])
	}
	format := C.int(a.format)
	if a.fused {
		format |= C.MPEGB200_AUDIO_WINDOW_FMA
	}
	id := C.int32_t(a.stream)
	if C.mpegb200_audio_synth(a.gpu.ctx, 1, &id, 1, (*C.int32_t)(unsafe.Pointer(&a.pinned[0])), format, out) != 0 {
		return nil
	}
	a.samples.Time = float64(t)
	a.samples.format = a.format
	a.time = float64(t) + float64(SamplesPerFrame)/float64(a.Samplerate())
	return &a.samples
}

// ---------------------------------------------------------------------------------------------------
// MPEG (mpeg.go:58-456): a program stream with one video and one audio stream
// ---------------------------------------------------------------------------------------------------

// MPEG mirrors the decode half of mpeg.MPEG.  The player logic of mpeg.go (Decode(tick), callbacks, Seek) is
// host-side control flow over DecodeVideo / DecodeAudio and is out of scope of the hot path (SURVEY section 8).
type MPEG struct {
	video *Video
	audio *Audio

	NumVideoPackets int
	NumAudioPackets int
}

// New mirrors mpeg.New (mpeg.go:85-116): reads the whole program stream, checks the pack start code,
// demultiplexes it and sets up the decoders on the default GPU.
func New(r io.Reader) (*MPEG, error) {
	gpu, err := DefaultGPU()
	if err != nil {
		return nil, err
	}
	return NewOnGPU(gpu, r)
}

// NewOnGPU is New with an explicit context (multi-GPU: pick the GPU the stream lives on).
func NewOnGPU(gpu *GPU, r io.Reader) (*MPEG, error) {
	data, err := io.ReadAll(r)
	if err != nil {
		return nil, err
	}
	if len(data) < 4 || !bytes.Equal([]byte{0x00, 0x00, 0x01, 0xBA}, data[0:4]) { // mpeg.go:95-100
		return nil, ErrInvalidMPEG
	}
	var vp, ap *C.uint8_t
	var vl, al C.size_t
	var nv, na C.int
	if C.mpegb200_demux_split((*C.uint8_t)(unsafe.Pointer(&data[0])), C.size_t(len(data)), &vp, &vl, &ap, &al, &nv, &na) != 0 {
		return nil, ErrInvalidMPEG // demux.ErrInvalidHeader, demux.go:32
	}
	defer C.mpegb200_buffer_free(unsafe.Pointer(vp))
	defer C.mpegb200_buffer_free(unsafe.Pointer(ap))
	m := &MPEG{NumVideoPackets: int(nv), NumAudioPackets: int(na)}
	if vl > 0 {
		if m.video, err = NewVideo(gpu, unsafe.Slice((*byte)(unsafe.Pointer(vp)), int(vl))); err != nil {
			return nil, err
		}
	}
	if al > 0 {
		if m.audio, err = NewAudio(gpu, unsafe.Slice((*byte)(unsafe.Pointer(ap)), int(al))); err != nil {
			return nil, err
		}
	}
	return m, nil
}

// Video mirrors MPEG.Video (mpeg.go:160-162).
func (m *MPEG) Video() *Video { return m.video }

// Audio mirrors MPEG.Audio (mpeg.go:224-226).
func (m *MPEG) Audio() *Audio { return m.audio }

// Width mirrors MPEG.Width (mpeg.go:197-203).
func (m *MPEG) Width() int {
	if m.video == nil {
		return 0
	}
	return m.video.Width()
}

// Height mirrors MPEG.Height (mpeg.go:206-212).
func (m *MPEG) Height() int {
	if m.video == nil {
		return 0
	}
	return m.video.Height()
}

// Framerate mirrors MPEG.Framerate (mpeg.go:215-221).
func (m *MPEG) Framerate() float64 {
	if m.video == nil {
		return 0
	}
	return m.video.Framerate()
}

// Samplerate mirrors MPEG.Samplerate (mpeg.go:282-288).
func (m *MPEG) Samplerate() int {
	if m.audio == nil {
		return 0
	}
	return m.audio.Samplerate()
}

// SetAudioFormat mirrors MPEG.SetAudioFormat (mpeg.go:234-238).
func (m *MPEG) SetAudioFormat(format AudioFormat) {
	if m.audio != nil {
		m.audio.SetFormat(format)
	}
}

// DecodeVideo mirrors MPEG.DecodeVideo (mpeg.go:416-435).
func (m *MPEG) DecodeVideo() *Frame {
	if m.video == nil {
		return nil
	}
	return m.video.Decode()
}

// DecodeAudio mirrors MPEG.DecodeAudio (mpeg.go:438-457).
func (m *MPEG) DecodeAudio() *Samples {
	if m.audio == nil {
		return nil
	}
	return m.audio.Decode()
}

// Close releases both decoders.
func (m *MPEG) Close() {
	if m.video != nil {
		m.video.Close()
	}
	if m.audio != nil {
		m.audio.Close()
	}
}
