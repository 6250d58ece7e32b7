// Package b200 binds libauditory_b200.so (C-ABI: include/auditory_b200.h) for
// github.com/emer/auditory.  It is the only cgo in the module: sound.SndEnv,
// dft.Params, mel.Params and agabor.FilterSet keep their exported surface and
// call into this package instead of running the Go loops (see INTEGRATION.md).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain): kept mechanical
// and thin on purpose.  Needs Go 1.21+ (runtime.Pinner).  Build where Go exists with
//
//	CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/auditory_b200/lib -lauditory_b200" go build ./...
package b200

/*
#include <stdlib.h>
#include "auditory_b200.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// Params mirrors aud_params: SndEnv.Params after Init (samples), dft.Params,
// mel.Params / mel.FilterBank and the gabor FilterSet geometry.
type Params struct {
	SampleRate                                                                            int
	WinSamples, StepSamples, SegmentSamples, StrideSamples, SegmentSteps, BorderSteps     int
	CompLogPow                                                                            bool
	LogMin, LogOffSet, PrevSmooth, CurSmooth                                              float64
	NMel                                                                                  int
	MelLogOff, MelLogMin                                                                  float64
	Renorm                                                                                bool
	RenormMin, RenormScale                                                                float64
	MFCC                                                                                  bool
	NCoefs                                                                                int
	Deltas, C0Energy                                                                      bool
	GaborNF, GaborSizeX, GaborSizeY, GaborStrideX, GaborStrideY                           int
	GaborGain                                                                             float64
	GaborShape                                                                            []int // 2 or 4 dims
	GaborByTime                                                                           bool
}

func b2i(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

// lastErr reads the calling THREAD's error text: callers hold runtime.LockOSThread across the failing call and this
// one, so that the goroutine cannot migrate to another OS thread in between.
func lastErr(rc C.int32_t) error {
	return errors.New("auditory_b200: " + C.GoString(C.aud_last_error()))
}

// Pipeline owns one aud_handle (one GPU, one goroutine at a time).
type Pipeline struct {
	h    *C.aud_handle
	dims C.aud_dims
}

// New uploads the tables the Go Init code already computes: mel.Params.BinPts,
// MelFilters.Values (float64 [NFilters, NFilters+2]) and FilterSet.Filters.Values.
func New(p *Params, binPts []int32, melFilters []float64, gaborFilters []float64, device int) (*Pipeline, error) {
	var cp C.aud_params
	cp.sample_rate = C.int32_t(p.SampleRate)
	cp.win_samples, cp.step_samples = C.int32_t(p.WinSamples), C.int32_t(p.StepSamples)
	cp.segment_samples, cp.stride_samples = C.int32_t(p.SegmentSamples), C.int32_t(p.StrideSamples)
	cp.segment_steps, cp.border_steps = C.int32_t(p.SegmentSteps), C.int32_t(p.BorderSteps)
	cp.comp_log_pow = b2i(p.CompLogPow)
	cp.log_min, cp.log_offset = C.double(p.LogMin), C.double(p.LogOffSet)
	cp.prev_smooth, cp.cur_smooth = C.double(p.PrevSmooth), C.double(p.CurSmooth)
	cp.n_mel = C.int32_t(p.NMel)
	cp.mel_log_off, cp.mel_log_min = C.double(p.MelLogOff), C.double(p.MelLogMin)
	cp.renorm = b2i(p.Renorm)
	cp.renorm_min, cp.renorm_scale = C.double(p.RenormMin), C.double(p.RenormScale)
	cp.mfcc, cp.n_coefs, cp.deltas, cp.mfcc_c0_energy = b2i(p.MFCC), C.int32_t(p.NCoefs), b2i(p.Deltas), b2i(p.C0Energy)
	cp.gabor_nf = C.int32_t(p.GaborNF)
	cp.gabor_size_x, cp.gabor_size_y = C.int32_t(p.GaborSizeX), C.int32_t(p.GaborSizeY)
	cp.gabor_stride_x, cp.gabor_stride_y = C.int32_t(p.GaborStrideX), C.int32_t(p.GaborStrideY)
	cp.gabor_gain = C.double(p.GaborGain)
	cp.gabor_out_dims = C.int32_t(len(p.GaborShape))
	for i, d := range p.GaborShape {
		cp.gabor_shape[i] = C.int32_t(d)
	}
	cp.gabor_by_time = b2i(p.GaborByTime)
	if len(binPts) != p.NMel+2 || len(melFilters) != p.NMel*(p.NMel+2) || (p.GaborNF > 0 && len(gaborFilters) == 0) {
		return nil, errors.New("auditory_b200: mel / gabor tables do not match the parameters")
	}
	var gab *C.double
	if p.GaborNF > 0 {
		gab = (*C.double)(unsafe.Pointer(&gaborFilters[0]))
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	pl := &Pipeline{}
	// cp lives on the Go stack and holds no pointers; the table slices are passed as direct arguments (plain Go
	// pointers to pointer-free memory), which the cgo rules allow without pinning
	var h *C.aud_handle
	rc := C.aud_create(&cp, (*C.int32_t)(unsafe.Pointer(&binPts[0])), (*C.double)(unsafe.Pointer(&melFilters[0])),
		gab, nil, C.int32_t(device), &h)
	if rc != 0 {
		return nil, lastErr(rc)
	}
	pl.h = h
	C.aud_get_dims(pl.h, &pl.dims)
	runtime.SetFinalizer(pl, func(p *Pipeline) { p.Close() })
	return pl, nil
}

func (pl *Pipeline) Close() {
	if pl.h != nil {
		C.aud_destroy(pl.h)
		pl.h = nil
	}
}

// SegCnt is SndEnv.Init's segment count for a signal of n samples (sndenv.go:263-265).
func (pl *Pipeline) SegCnt(n int) int { return int(C.aud_seg_count(pl.h, C.int32_t(n))) }

// Outputs holds the per-segment tensors of a batch, [segments][...] row-major
// in the layouts of MelFBankSegment, MFCCSegment, ..., GborOutput.
type Outputs struct {
	Mel, MFCC, Deltas, DeltaDeltas, Energy, Gabor []float32
	Segments                                      int
}

func ptr(s []float32) *C.float {
	if len(s) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&s[0]))
}

// pinAll pins the backing arrays of the slices a call hands to C inside C structs.  The cgo rules allow a Go pointer
// to memory that holds Go pointers only if those pointers are pinned (runtime.Pinner, Go 1.21+): aud_batch and
// aud_outputs are such structs, so every slice they point at is pinned for the duration of the call.
func pinAll(pin *runtime.Pinner, f32 [][]float32, i64 []int64, i32 []int32, i16 []int16) {
	for _, s := range f32 {
		if len(s) > 0 {
			pin.Pin(&s[0])
		}
	}
	if len(i64) > 0 {
		pin.Pin(&i64[0])
	}
	if len(i32) > 0 {
		pin.Pin(&i32[0])
	}
	if len(i16) > 0 {
		pin.Pin(&i16[0])
	}
}

func (o *Outputs) all() [][]float32 {
	return [][]float32{o.Mel, o.MFCC, o.Deltas, o.DeltaDeltas, o.Energy, o.Gabor}
}

// Process runs every segment of every utterance through the fused kernel.  wave, uttOffset, uttLen and the output
// slices are ordinary Go memory, pinned for the call; the library stages them through its own page-locked bounce
// buffers and never keeps a pointer after the call returns.  AllocPinned gives buffers that skip that staging.
func (pl *Pipeline) Process(wave []float32, uttOffset []int64, uttLen []int32, addSamples int, want Outputs) (Outputs, error) {
	if len(uttLen) == 0 || len(uttOffset) != len(uttLen) {
		return Outputs{}, checkLens(uttOffset, uttLen)
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	out, o := pl.outputs(uttLen, want)
	var pin runtime.Pinner
	defer pin.Unpin()
	pinAll(&pin, append(out.all(), wave), uttOffset, uttLen, nil)
	b := C.aud_batch{wave: ptr(wave), utt_offset: (*C.int64_t)(unsafe.Pointer(&uttOffset[0])),
		utt_len: (*C.int32_t)(unsafe.Pointer(&uttLen[0])), n_utt: C.int32_t(len(uttLen)), add_samples: C.int32_t(addSamples)}
	if rc := C.aud_process_host(pl.h, &b, &o); rc != 0 {
		return out, lastErr(rc)
	}
	return out, nil
}

// ProcessMulti is Process over several GPUs of the box: pls[g] was built with the same Params on device g.  The
// library cuts the utterances into contiguous blocks, runs one host thread per GPU and lets every GPU write its own
// range of the output slices; there is no collective (aud_process_host_multi).
func ProcessMulti(pls []*Pipeline, wave []float32, uttOffset []int64, uttLen []int32, addSamples int, want Outputs) (Outputs, error) {
	if len(pls) == 0 {
		return Outputs{}, errors.New("auditory_b200: no pipelines")
	}
	if len(uttLen) == 0 || len(uttOffset) != len(uttLen) {
		return Outputs{}, checkLens(uttOffset, uttLen)
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	out, o := pls[0].outputs(uttLen, want)
	hs := (**C.aud_handle)(C.malloc(C.size_t(len(pls)) * C.size_t(unsafe.Sizeof(pls[0].h)))) // C memory: holds C pointers only
	defer C.free(unsafe.Pointer(hs))
	for g, p := range pls {
		unsafe.Slice(hs, len(pls))[g] = p.h
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	pinAll(&pin, append(out.all(), wave), uttOffset, uttLen, nil)
	b := C.aud_batch{wave: ptr(wave), utt_offset: (*C.int64_t)(unsafe.Pointer(&uttOffset[0])),
		utt_len: (*C.int32_t)(unsafe.Pointer(&uttLen[0])), n_utt: C.int32_t(len(uttLen)), add_samples: C.int32_t(addSamples)}
	if rc := C.aud_process_host_multi(hs, C.int32_t(len(pls)), &b, &o); rc != 0 {
		return out, lastErr(rc)
	}
	return out, nil
}

// ProcessPCM16 is Process for 16-bit PCM as decoded from a WAV file, before Wave.GetFloatAtIdx normalises it
// (sound/sound.go:130-141): the samples are divided by 0x7FFF on the GPU and only half the bytes cross PCIe.
// This is the recommended ingest for files.
func (pl *Pipeline) ProcessPCM16(wave []int16, uttOffset []int64, uttLen []int32, addSamples int, want Outputs) (Outputs, error) {
	if len(uttLen) == 0 || len(uttOffset) != len(uttLen) || len(wave) == 0 {
		return Outputs{}, checkLens(uttOffset, uttLen)
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	out, o := pl.outputs(uttLen, want)
	var pin runtime.Pinner
	defer pin.Unpin()
	pinAll(&pin, out.all(), uttOffset, uttLen, wave)
	rc := C.aud_process_host_i16(pl.h, (*C.int16_t)(unsafe.Pointer(&wave[0])), (*C.int64_t)(unsafe.Pointer(&uttOffset[0])),
		(*C.int32_t)(unsafe.Pointer(&uttLen[0])), C.int32_t(len(uttLen)), C.int32_t(addSamples), &o)
	if rc != 0 {
		return out, lastErr(rc)
	}
	return out, nil
}

// checkLens: an empty batch is not an error (no utterances, no segments); mismatched slices are.
func checkLens(uttOffset []int64, uttLen []int32) error {
	if len(uttOffset) != len(uttLen) {
		return errors.New("auditory_b200: uttOffset and uttLen differ in length")
	}
	return nil
}

// outputs sizes (or reuses) the result slices for a batch and points an aud_outputs at them.  uttLen is not empty.
func (pl *Pipeline) outputs(uttLen []int32, want Outputs) (Outputs, C.aud_outputs) {
	var pin runtime.Pinner
	pin.Pin(&uttLen[0])
	nseg := int(C.aud_total_segments(pl.h, (*C.int32_t)(unsafe.Pointer(&uttLen[0])), C.int32_t(len(uttLen)), nil))
	pin.Unpin()
	S := int(pl.dims.segment_steps)
	grow := func(s []float32, per int, on bool) []float32 {
		if !on {
			return nil
		}
		if cap(s) < nseg*per {
			return make([]float32, nseg*per)
		}
		return s[:nseg*per]
	}
	out := Outputs{Segments: nseg}
	out.Mel = grow(want.Mel, int(pl.dims.n_mel)*S, true)
	out.Energy = grow(want.Energy, S, true)
	out.MFCC = grow(want.MFCC, int(pl.dims.n_coefs)*S, want.MFCC != nil)
	out.Deltas = grow(want.Deltas, int(pl.dims.n_coefs)*S, want.Deltas != nil)
	out.DeltaDeltas = grow(want.DeltaDeltas, int(pl.dims.n_coefs)*S, want.DeltaDeltas != nil)
	out.Gabor = grow(want.Gabor, int(pl.dims.gabor_len), want.Gabor != nil)
	o := C.aud_outputs{mel: ptr(out.Mel), mfcc: ptr(out.MFCC), deltas: ptr(out.Deltas), delta_deltas: ptr(out.DeltaDeltas),
		energy: ptr(out.Energy), gabor: ptr(out.Gabor)}
	return out, o
}

// AllocPinned returns n float32 of page-locked host memory as a Go slice
// (C memory: the GC does not move or free it; release with FreePinned).
func AllocPinned(n int) []float32 {
	if n <= 0 {
		return nil
	}
	p := C.aud_host_alloc(C.uint64_t(n * 4))
	if p == nil {
		return nil
	}
	return unsafe.Slice((*float32)(p), n)
}

func FreePinned(s []float32) {
	if len(s) > 0 {
		C.aud_host_free(unsafe.Pointer(&s[0]))
	}
}
